"""Dev tool (GPU box): scattered channels-last texel gathers per SM vs load form / loads in flight / L1 left (csrc/microbench.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddmi_b200 import _lib
dev = 'cuda:0'
ntexel = 3 * (16 * 16 + 32 * 32 + 64 * 64)          # one item's nine occupancy planes: 4.1 MB
table = torch.randn(ntexel * 64, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
sink = torch.zeros(148 * 256, device=dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
names = ['ld.global.nc', 'ld.global.cg', 'ld.global.nc.L1::no_allocate', 'ld.global.cv', 'ld.global.L1::evict_first']
print("smem_KB variant                       U | cycles/texel-round  B/clk/SM")
for smem_kb in (224, 160, 32):
    for var, nm in enumerate(names):
        for u in (4, 12):
            iters = 400
            for _ in range(2):
                _lib.check(L.ddmi_debug_gatherbench(var, u, table.data_ptr(), ntexel, iters, smem_kb, 148, out.data_ptr(),
                                                    sink.data_ptr(), st))
                torch.cuda.synchronize()
            cyc = out.cpu().tolist()[0]
            byts = 32 * u * iters * 256           # 32 groups of 8 threads x U texels x 256 B per round
            print(f"{smem_kb:7d} {nm:30s} {u:2d} | {cyc / iters:10.1f} {byts / cyc:10.1f}")
