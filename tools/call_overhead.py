"""Dev tool (GPU box): host-side cost of one decoder call (tiny query sets: the kernels take ~0.1 ms)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddmi_b200, bench
from oracle import cases
torch.set_grad_enabled(False)
dev = 'cuda:0'
m = cases.build_module('occupancy').to(dev)
pts, hdbf = cases.occupancy_inputs(batch=1, n=20000)
c = tuple([t.to(dev) for t in ax] for ax in hdbf)
p = pts.to(dev)
im = bench.build_mlp().to(dev)
coords, planes, si = cases.image_inputs(batch=1, sizes=(64, 128, 256), res=128)
coords, planes = coords.to(dev), [t.to(dev) for t in planes]
for name, fn in (('occupancy 20k points', lambda: m(p, c).logits), ('image 128x128', lambda: im(coords, hdbf=planes, si=si))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(200): fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t) / 200 * 1e3:.3f} ms per call (wall, back to back)")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(100): m(p, c).logits
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
