"""Dev tool: MMA-side cycle counters of the tcgen05 occupancy kernel (CTA 0)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddmi_b200
from ddmi_b200 import _lib
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else 'grid'
dev = 'cuda:0'
torch.manual_seed(1)
m = ddmi_b200.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256).to(dev)
g = torch.Generator().manual_seed(1)
hdbf = tuple([torch.randn(B, 64, s, s, generator=g).to(dev) for s in (16, 32, 64)] for _ in range(3))
if mode in ('grid', 'lattice'):
    pts = (1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (128,) * 3)).to(dev)
else:
    pts = ((torch.rand(2097152, 3, generator=g) - 0.5) * 1.1).to(dev)
if mode == 'lattice':
    ax = (1.1 * torch.linspace(-0.5, 0.5, 128)).to(dev)
    f = lambda: m.decode_logits_lattice((ax, ax, ax), hdbf)
else:
    f = lambda: m(pts[None].expand(B, -1, -1), hdbf).logits
for _ in range(2): f()
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 8)()
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); f(); t1.record(); torch.cuda.synchronize()
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
v = list(buf); tiles = max(v[6], 1)
print(f"{mode}: ms {t0.elapsed_time(t1):.2f} coords/s {B*pts.shape[0]/t0.elapsed_time(t1)*1e3:.3e} pair-tiles(cta0) {v[6]}")
print(f"per tile cycles: MMA wait-operands {v[3]/tiles:.0f}  wait-weights {v[4]/tiles:.0f}  MMA total {v[5]/tiles:.0f}")
