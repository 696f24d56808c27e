"""Seeded parity cases shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE.

Each case is (decoder module with deterministic weights, inputs).  Weights are
the package's own constructors under a fixed torch seed (same distributions as
the reference's), then perturbed where the default init would hide half the
network: ``fc_1.weight`` is zero-initialised in the reference (blocks.py:705,
SURVEY.md F4) so it is re-drawn; activation / output biases are made non-zero;
the NeRF density / colour heads are scaled so rgb_map spans a useful range.
``noise.weight`` stays 0 (parity is undefined otherwise).
"""
import math

import numpy as np
import torch

import ddmi_b200
from ddmi_b200 import nerf_helpers as nh

SEED = 777  # the reference's default seed (main.py:58)


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _randn(g, *shape):
    return torch.randn(*shape, generator=g)


def _perturb_common(m, g):
    for name, p in m.named_parameters():
        if name.endswith('fc_1.weight'):
            bound = 1.0 / math.sqrt(p.shape[1])
            p.data = (torch.rand(p.shape, generator=g) * 2 - 1) * bound   # kaiming_uniform_(a=sqrt(5))
        elif name.endswith('activate.bias') or name == 'torgb.bias':
            p.data = 0.1 * _randn(g, *p.shape)


def build_module(kind, seed=SEED):
    torch.manual_seed(seed)
    g = _gen(seed + 1)
    if kind in ('image', 'image_noise'):
        m = ddmi_b200.MLP(in_ch=2, latent_dim=64, out_ch=3, ch=256)
    elif kind == 'occupancy':
        m = ddmi_b200.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256)
    elif kind == 'video':
        m = ddmi_b200.MLPVideo(in_ch=2, latent_dim=64, out_ch=3, ch=256)
    elif kind == 'nerf':
        m = ddmi_b200.MLPNeRF(D=6, W=256, in_channels_xyz=159, skips=[2, 4], in_channels_dir=27)
        m.sigma.weight.data *= 6.0
        m.sigma.bias.data += 0.5
        m.rgb[0].weight.data *= 8.0
    else:
        raise KeyError(kind)
    _perturb_common(m, g)
    if kind == 'image_noise':          # a "trained" checkpoint: every NoiseInjection.weight non-zero (blocks.py:286-297)
        for name, p in m.named_parameters():
            if name.endswith('noise.weight'):
                p.data = 0.05 + 0.15 * torch.rand(1, generator=g)
    return m.eval()


def state_dict32(m):
    return {k: v.detach().clone().float().cpu() for k, v in m.state_dict().items()}


def image_inputs(batch=2, sizes=(16, 32, 64), res=96, seed=SEED):
    g = _gen(seed + 10)
    planes = [_randn(g, batch, 64, s, s) for s in sizes]
    e = (res - 1) / res
    coords = ddmi_b200.convert_to_coord_format_2d(1, res, res, hstart=-e, hend=e, wstart=-e, wend=e)
    si = ddmi_b200.get_scale_injection(res, anchor_res=sizes[-1])
    return coords, planes, si


def image_noise_tensors(batch, res, seed=SEED):
    """12 explicit (batch,1,res,res) N(0,1) tensors, one per NoiseInjection (net_res1.conv1 .. net_res4.conv3)."""
    g = _gen(seed + 11)
    return [_randn(g, batch, 1, res, res) for _ in range(12)]


def occupancy_inputs(batch=2, sizes=(16, 32, 64), n=20000, spread=0.65, seed=SEED):
    g = _gen(seed + 20)
    hdbf = tuple([_randn(g, batch, 64, s, s) for s in sizes] for _ in range(3))
    pts = (torch.rand(batch, n, 3, generator=g) * 2 - 1) * spread   # beyond +-0.55: exercises the clamp
    if n >= 2:
        pts[0, 0] = torch.tensor([0.55, -0.55, 0.0])
        pts[0, 1] = torch.tensor([0.7, -0.7, 0.56])
    return pts, hdbf


def video_inputs(batch=1, T=4, sizes=(8, 16, 32), seed=SEED):
    g = _gen(seed + 30)
    xy = [_randn(g, batch, 64, s, s) for s in sizes]
    yt = [_randn(g, batch, 64, T, s) for s in sizes]
    xt = [_randn(g, batch, 64, T, s) for s in sizes]
    R = sizes[-1]
    e, et = (R - 1) / R, (T - 1) / T
    coords = ddmi_b200.convert_to_coord_format_3d(1, R, R, T, hstart=-e, hend=e, wstart=-e, wend=e,
                                                  tstart=-et, tend=et)
    return coords, (xy, yt, xt)


NERF_PERTURB_SEED = 4242     # torch.manual_seed before a stratified (perturb > 0) render: fixes the CPU torch.rand draw
NERF_CFG = {'model': {'TN': {'netchunk': 40000, 'peturb': 0, 'N_importance': 0, 'N_samples': 64,
                             'use_viewdirs': True, 'white_bkgd': True, 'raw_noise_std': 0}}}


def nerf_inputs(res=32, theta=40.0, seed=SEED):
    g = _gen(seed + 40)
    fea = {k: _randn(g, 1, 32, 64, 64) for k in ('xy', 'yz', 'xz')}
    focal = .5 * res / np.tan(.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 0.5 * res], [0, focal, 0.5 * res], [0, 0, 1]])
    c2w = nh.pose_spherical(theta, -20, 5)[:3, :4]
    return res, K, fea, c2w


def nerf_mlp_inputs(n=1500, seed=SEED):
    g = _gen(seed + 50)
    return _randn(g, n, 186)


def checksum(tensors):
    return float(sum(t.double().abs().sum() for t in tensors))


# ---------------------------------------------------------------------------
# occupancy post-step (marching cubes): seeded logit-like volumes
# ---------------------------------------------------------------------------
def mcubes_volume(shape, seed=0, noise=0.8, scale=1.0):
    """float32 (nx, ny, nz): a noisy ball (positive inside) -- many surface cases, touches the volume faces for small shapes."""
    g = torch.Generator().manual_seed(SEED + 70 + seed)
    ax = [torch.arange(s, dtype=torch.float32) - (s - 1) / 2 for s in shape]
    gx, gy, gz = torch.meshgrid(*ax, indexing='ij')
    r = torch.sqrt(gx * gx + gy * gy + gz * gz)
    vol = 0.4 * min(shape) - r + noise * torch.randn(shape, generator=g)
    return (scale * vol).to(torch.float32).contiguous()
