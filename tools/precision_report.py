"""Dev tool (GPU): max-abs error of every decoder family against the committed golden vectors, per precision."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
from ddmi_b200 import nerf_helpers as nh
torch.set_grad_enabled(False)
DEV = 'cuda:0'
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
gold = lambda n: torch.load(os.path.join(G, n + '.pt'))
cu = lambda x: x.to(DEV) if torch.is_tensor(x) else (type(x)(cu(v) for v in x) if isinstance(x, (list, tuple)) else {k: cu(v) for k, v in x.items()})
rows = []
for prec in ('fp32', 'bf16x3', 'f16f8'):
    m = cases.build_module('image').to(DEV); m.precision = prec
    c, p, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    e_img = float((m(c.to(DEV), hdbf=cu(p), si=si).cpu() - gold('image_96')['out']).abs().max())
    m = cases.build_module('occupancy').to(DEV); m.precision = prec
    pts, h = cases.occupancy_inputs()
    out = m(pts.to(DEV), cu(h)).logits.cpu(); g = gold('occupancy')['out']
    e_occ = float((out - g).abs().max()); sign = float(((out > 0) == (g > 0)).float().mean())
    m = cases.build_module('video').to(DEV); m.precision = prec
    c, h = cases.video_inputs()
    e_vid = float((m(cu(c), cu(h)).cpu() - gold('video')['out']).abs().max())
    m = cases.build_module('nerf').to(DEV); m.precision = prec
    res, K, fea, c2w = cases.nerf_inputs()
    e1, _ = nh.get_embedder(10, 0); e2, _ = nh.get_embedder(4, 0)
    kw = nh.get_render_kwargs(cases.NERF_CFG, m, e1, e2)
    rgb = nh.render(res, res, K, cu(fea), None, 0, DEV, c2w=c2w, **kw).cpu()
    e_nrf = float((rgb - gold('nerf_render')['out']).abs().max())
    rows.append((prec, e_img, e_occ, sign, e_vid, e_nrf))
print("| precision | image RGB | occupancy logit (sign agreement) | video RGB | NeRF rgb_map |")
print("|---|---|---|---|---|")
for r in rows:
    print("| %s | %.2e | %.2e (%.4f%%) | %.2e | %.2e |" % (r[0], r[1], r[2], 100 * r[3], r[4], r[5]))
