"""Dev tool (GPU box): tcgen05.mma rate in the decode kernels' configuration (csrc/microbench.cu::mmabench_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddmi_b200 import _lib
dev = 'cuda:0'
seed = torch.randn(2048, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
sink = torch.zeros(256, device=dev)
L = _lib.lib()
iters = 4000
print("scope  N   A-operand kind    stores | cycles / MMA")
for base in (100, 200):
    for v in range(32):
        if (v & 8) and (v & 4):
            continue
        for _ in range(2):
            _lib.check(L.ddmi_debug_microbench(base + v, iters, seed.data_ptr(), out.data_ptr(), sink.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        kind = 'f16f8 mix' if v & 8 else ('fp8 K=32' if v & 4 else 'fp16 K=16')
        print(f"{'pair ' if base == 100 else 'chip '} {128 if v & 1 else 256:3d}  {'TMEM' if v & 2 else 'smem'}      {kind:9s}  {'yes' if v & 16 else 'no ':3s}  | {o[0] / o[1]:7.1f}")
