"""Generate tests/golden/*.pt by running the REFERENCE itself (build container only).

Run:  python oracle/make_golden.py [names]  (needs /root/reference, nvcc + ninja for the
                                             reference's JIT ops; ~2 min the first time;
                                             names = fixtures to (re)write, default all)
Imports models/d2c_vae/mlp.py and utils/nerf_helpers.py from /root/reference with the
two CPU shims of SURVEY.md Appendix A (imageio stub; CPU restatement of
fused_leaky_relu, semantics from op/fused_bias_act_kernel.cu:28-47), loads the
package's seeded weights into the reference modules (which also proves state-dict
compatibility), runs them on the seeded inputs of oracle/cases.py and stores the
outputs.  It also prints oracle-vs-reference differences.
"""
import os
import sys
import types

os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
os.environ.setdefault('TORCH_EXTENSIONS_DIR', '/tmp/refext')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.modules['imageio'] = types.ModuleType('imageio')

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import models.d2c_vae.blocks as rblocks  # noqa: E402  (triggers the reference's JIT build)

_cpu_flr = lambda x, b, ns=0.2, sc=2 ** 0.5: F.leaky_relu(x + b.view(1, -1, *[1] * (x.ndim - 2)), ns) * sc
rblocks.fused_leaky_relu = _cpu_flr
rblocks.FusedLeakyReLU.forward = lambda self, x: _cpu_flr(x, self.bias, self.negative_slope, self.scale)

from models.d2c_vae import mlp as rmlp  # noqa: E402
from utils import nerf_helpers as rnh  # noqa: E402

from oracle import cases, ddmi_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


ONLY = set(sys.argv[1:])     # optional: names of the fixtures to (re)write; default all


def save(name, out, inputs, extra=None):
    if ONLY and name not in ONLY:
        print(f'{name}: not selected, left untouched')
        return
    d = {'out': out.float().contiguous(), 'input_checksum': cases.checksum(inputs)}
    d.update(extra or {})
    torch.save(d, os.path.join(OUT, name + '.pt'))
    print(f'{name}: out {tuple(out.shape)} |out|max {float(out.abs().max()):.4f}')


def flat(x):
    return [t for a in x for t in (a if isinstance(a, (list, tuple)) else [a])]


# ---- image ---------------------------------------------------------------
m = cases.build_module('image')
sd = cases.state_dict32(m)
ref = rmlp.MLP(in_ch=2, latent_dim=64, out_ch=3, ch=256)
ref.load_state_dict(sd, strict=True)
for tag, kw in (('image_96', dict(batch=2, sizes=(16, 32, 64), res=96)),
                ('image_native', dict(batch=1, sizes=(8, 16, 32), res=32))):
    coords, planes, si = cases.image_inputs(**kw)
    out = ref(coords, hdbf=planes, si=si)
    o2 = orc.image_decode(sd, coords, planes, si)
    print(f'  oracle-vs-reference {tag}: {float((out - o2).abs().max()):.3e}')
    save(tag, out, planes + list(sd.values()), {'si': si})

# ---- image with noise injection: a checkpoint whose NoiseInjection weights are non-zero, EXPLICIT noise tensors ------------
# StyledResBlock.forward does not forward a `noise=` argument (blocks.py:624-627), so the explicit tensors are handed to
# NoiseInjection.forward (blocks.py:292-297, its own `noise` parameter) through a FIFO, in call order.
_noise_fifo = []
_orig_noise_forward = rblocks.NoiseInjection.forward


def _noise_forward(self, image, noise=None):
    if noise is None and _noise_fifo:
        noise = _noise_fifo.pop(0)
    return _orig_noise_forward(self, image, noise=noise)


rblocks.NoiseInjection.forward = _noise_forward
m = cases.build_module('image_noise')
sd = cases.state_dict32(m)
ref = rmlp.MLP(in_ch=2, latent_dim=64, out_ch=3, ch=256)
ref.load_state_dict(sd, strict=True)
coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
noise = cases.image_noise_tensors(2, 96)
_noise_fifo.extend(noise)
out = ref(coords, hdbf=planes, si=si)
assert not _noise_fifo
o2 = orc.image_decode(sd, coords, planes, si, noise=noise)
o0 = orc.image_decode(cases.state_dict32(cases.build_module('image')), coords, planes, si)
print(f'  oracle-vs-reference image_noise: {float((out - o2).abs().max()):.3e}; effect of the noise on the output '
      f'{float((out - o0).abs().max()):.3f}')
save('image_noise', out, planes + noise + list(sd.values()), {'si': si})
rblocks.NoiseInjection.forward = _orig_noise_forward

# ---- occupancy -------------------------------------------------------------
m = cases.build_module('occupancy')
sd = cases.state_dict32(m)
ref = rmlp.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256)
ref.load_state_dict(sd, strict=True)
pts, hdbf = cases.occupancy_inputs()
out = ref(pts, hdbf).logits
o2 = orc.occupancy_logits(sd, pts, hdbf)
print(f'  oracle-vs-reference occupancy: {float((out - o2).abs().max()):.3e}')
save('occupancy', out, flat(hdbf) + [pts] + list(sd.values()))

# ---- video -----------------------------------------------------------------
m = cases.build_module('video')
sd = cases.state_dict32(m)
ref = rmlp.MLPVideo(in_ch=2, latent_dim=64, out_ch=3, ch=256)
ref.load_state_dict(sd, strict=True)
coords, hdbf = cases.video_inputs()
out = ref(coords, hdbf)
o2 = orc.video_decode(sd, coords, hdbf)
print(f'  oracle-vs-reference video: {float((out - o2).abs().max()):.3e}')
save('video', out, flat(hdbf) + list(sd.values()))

# ---- nerf --------------------------------------------------------------------
m = cases.build_module('nerf')
sd = cases.state_dict32(m)
ref = rmlp.MLPNeRF(D=6, W=256, in_channels_xyz=159, skips=[2, 4], in_channels_dir=27)
ref.load_state_dict(sd, strict=True)
x = cases.nerf_mlp_inputs()
out = ref(x)
o2 = orc.nerf_mlp(sd, x)
print(f'  oracle-vs-reference nerf_mlp: {float((out - o2).abs().max()):.3e}')
save('nerf_mlp', out, [x] + list(sd.values()))

res, K, fea, c2w = cases.nerf_inputs()
embed_fn, _ = rnh.get_embedder(10, 0)
embeddirs_fn, _ = rnh.get_embedder(4, 0)
kw = rnh.get_render_kwargs(cases.NERF_CFG, ref, embed_fn, embeddirs_fn)
rgb = rnh.render(res, res, K, fea, None, 0, 'cpu', chunk=4096, c2w=c2w, verbose=True, retraw=True,
                 hw_idx=None, **kw)
# oracle on the same rays
ro, rd = rnh.get_rays(res, res, K, c2w, 'cpu')
vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
rays = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), 2. * torch.ones(res * res, 1), 6. * torch.ones(res * res, 1), vd], -1)
o2 = orc.nerf_render_rays(sd, rays, fea, 64, True)
print(f'  oracle-vs-reference nerf_render: {float((rgb - o2).abs().max()):.3e}; rgb range {float(rgb.min()):.3f}..{float(rgb.max()):.3f}')
save('nerf_render', rgb, list(fea.values()) + list(sd.values()), {'rays': rays})

# hierarchical sampling: the reference's sample_pdf on seeded bins / weights, deterministic u and its own pytest-mode u
gpdf = torch.Generator().manual_seed(cases.SEED + 60)
z = torch.sort(2. + 4. * torch.rand(300, 64, generator=gpdf), -1).values
bins_pdf = .5 * (z[:, 1:] + z[:, :-1])                       # (300, 63), as render_rays builds z_vals_mid (:402-404)
w_pdf = torch.rand(300, 62, generator=gpdf) ** 4             # peaky weights, some ~0
w_pdf[7] = 0.                                                # an all-zero row (the 1e-5 floor makes it uniform)
out_det = rnh.sample_pdf(bins_pdf, w_pdf, 128, det=True)
out_rnd = rnh.sample_pdf(bins_pdf, w_pdf, 96, det=False, pytest=True)
print(f'  oracle-vs-reference sample_pdf: {float((orc.sample_pdf(bins_pdf, w_pdf, torch.linspace(0., 1., 128).expand(300, 128)) - out_det).abs().max()):.3e}')
save('sample_pdf', out_det, [bins_pdf, w_pdf], {'bins': bins_pdf, 'weights': w_pdf, 'out_pytest': out_rnd})

# stratified sampling (perturb = 1: the training-time setting, one torch.rand draw on the CPU generator) and lindisp
for tag, extra in (('nerf_render_perturb', dict(perturb=1.0)), ('nerf_render_lindisp', dict(lindisp=True)),
                   ('nerf_render_perturb_lindisp', dict(perturb=1.0, lindisp=True))):
    kw2 = dict(kw)
    kw2.update(extra)
    torch.manual_seed(cases.NERF_PERTURB_SEED)
    rgb = rnh.render(res, res, K, fea, None, 0, 'cpu', chunk=4096, c2w=c2w, verbose=True, retraw=True, hw_idx=None, **kw2)
    torch.manual_seed(cases.NERF_PERTURB_SEED)
    o2 = orc.nerf_render_rays(sd, rays, fea, 64, True, perturb=extra.get('perturb', 0.), lindisp=extra.get('lindisp', False))
    print(f'  oracle-vs-reference {tag}: {float((rgb - o2).abs().max()):.3e}')
    save(tag, rgb, list(fea.values()) + list(sd.values()), {'rays': rays})
print('done')
