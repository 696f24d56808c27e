"""Dev tool (GPU box): cycles of the epilogue building blocks in isolation (csrc/microbench.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddmi_b200 import _lib
dev = 'cuda:0'
seed = torch.randn(2048, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
sink = torch.zeros(256, device=dev)
L = _lib.lib()
names = ['drain 4x ld.x32, 8 warps', 'drain 4x ld.x32, 4 warps', 'f16f8 split of 128 values', 'publish 32x st.shared.v4',
         'full stage (drain+act+split+publish)', 'tmem store 128 values', 'bf16 hi/lo split of 128 values',
         'publish 32x st.shared.v4, no fences', 'drain, one ld.x32 + wait at a time']
iters = 2000
for mode, nm in enumerate(names):
    for _ in range(2):
        _lib.check(L.ddmi_debug_microbench(mode, iters, seed.data_ptr(), out.data_ptr(), sink.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"mode {mode} {nm:40s}: warp0 {o[0] / iters:8.1f} cyc/iter   span {o[1] / iters:8.1f} cyc/iter")
