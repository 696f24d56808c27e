import torch, sys, os
sys.path.insert(0, '.')
from ddmi_b200 import nerf_helpers as nh
g = torch.load('tests/golden/sample_pdf.pt')
bins, w = g['bins'].cuda(), g['weights'].cuda()
for name, kw, ref in (('det', dict(det=True), g['out']), ('pytest', dict(det=False, pytest=True), g['out_pytest'])):
    out = nh.sample_pdf(bins, w, ref.shape[-1], **kw).cpu()
    d = (out - ref).abs()
    width = float((g['bins'][:, 1:] - g['bins'][:, :-1]).max())
    print(name, out.shape, ref.shape, 'frac<2e-5', float((d < 2e-5).float().mean()), 'max', float(d.max()), 'width', width, 'n>2e-5', int((d >= 2e-5).sum()), 'of', d.numel())
    idx = (d >= 2e-5).nonzero()[:5]
    for i in idx: print('  ', i.tolist(), float(out[tuple(i)]), float(ref[tuple(i)]))
