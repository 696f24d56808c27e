"""Dev tool: in-kernel cycle breakdown of the tcgen05 image kernel (CTA 0)."""
import ctypes, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ddmi_b200 import _lib
torch.set_grad_enabled(False)
B, R = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = 'cuda:0'
m = bench.build_mlp().to(dev)
g = torch.Generator().manual_seed(1)
planes = [torch.randn(B, 64, s, s, generator=g).to(dev) for s in (64, 128, 256)]
from ddmi_b200 import convert_to_coord_format_2d, get_scale_injection
e = (R - 1) / R
c = convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(dev)
for _ in range(2): m(c, hdbf=planes, si=get_scale_injection(R))
torch.cuda.synchronize()
flags = int(os.environ.get('DBG', '0'))
_lib.check(_lib.lib().ddmi_debug_set(flags))
buf = (ctypes.c_uint64 * 8)()
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); m(c, hdbf=planes, si=get_scale_injection(R)); t1.record(); torch.cuda.synchronize()
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
v = list(buf); tiles = max(v[6], 1)
print(f"dbg {flags} ms {t0.elapsed_time(t1):.2f}  coords/s {B*R*R/t0.elapsed_time(t1)*1e3:.3e}  tiles(cta0) {v[6]}")
print(f"per tile cycles: E-wait(MMA busy) {v[0]/tiles:.0f}  E-epilogue {v[1]/tiles:.0f}  E-gather {v[2]/tiles:.0f} | "
      f"MMA wait-operands {v[3]/tiles:.0f}  MMA wait-weights {v[4]/tiles:.0f}  MMA total {v[5]/tiles:.0f}")
