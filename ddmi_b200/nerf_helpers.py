"""NeRF query + volume renderer of the D2C-VAE decode path.

Mirrors the call surface of the reference's ``utils/nerf_helpers.py`` that the
trainers use (``get_embedder``, ``get_render_kwargs``, ``pose_spherical``,
``get_rays``, ``render``); everything between "rays" and "rgb_map" --
sample generation, triplane gather, positional embedding, MLPNeRF and
``raw2outputs`` compositing (nerf_helpers.py:296-530) -- is one call into the
C ABI (``ddmi_nerf_render``).  Ray generation stays host-side tensor plumbing.

Stratified ``perturb`` and ``lindisp`` are supported by handing the kernel a per-ray
depth table computed here exactly as the reference does (nerf_helpers.py:356-380, same
torch ops and the same CPU ``torch.rand`` draw, so a seeded run reproduces the
reference's samples).  Rejected loudly: hierarchical ``N_importance`` (the reference's
branch, :402-431, reuses the coarse pass's plane coordinates for the enlarged sample set
and cannot run), ``raw_noise_std`` (random), ``ndc`` rays, ``c2w_staticcam``.
"""
import math

import numpy as np
import torch

from . import _lib
from .mlp import MLPNeRF, _stream_ptr

PLANE_EXTENT = 3.5   # pts / 3.5 before the triplane lookup (nerf_helpers.py:384)


class _Embedder:
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]
    (nerf_helpers.py:82-112).  The fused kernel evaluates this in-kernel; the
    callable exists for API parity and for feeding MLPNeRF.forward directly."""

    def __init__(self, multires, input_dims=3):
        self.multires = multires
        self.input_dims = input_dims
        self.out_dim = input_dims * (1 + 2 * multires)
        self.freq_bands = 2. ** torch.linspace(0., multires - 1, steps=multires)

    def __call__(self, x):
        parts = [x]
        for f in self.freq_bands:
            parts.append(torch.sin(x * f))
            parts.append(torch.cos(x * f))
        return torch.cat(parts, -1)


def get_embedder(multires, i=0):
    """(embed_fn, out_dim); ``i == -1`` -> identity (nerf_helpers.py:115-130)."""
    if i == -1:
        return torch.nn.Identity(), 3
    e = _Embedder(multires)
    return e, e.out_dim


class NetworkFn:
    """``network_fn`` of the render kwargs: callable like the reference's
    ``lambda x: nerf(x)`` (nerf_helpers.py:55) but keeps the module reachable so
    ``render`` can hand its packed weights to the fused kernel."""

    def __init__(self, module):
        self.module = module

    def __call__(self, x):
        return self.module(x)


def get_render_kwargs(config, nerf, embed_fn, embeddirs_fn):
    """Same keys as the reference's dict (nerf_helpers.py:40-64)."""
    tn = config['model']['TN']
    return {
        'embed_fn': embed_fn,
        'embeddirs_fn': embeddirs_fn,
        'netchunk': tn['netchunk'],
        'perturb': tn['peturb'],
        'N_importance': tn['N_importance'],
        'network_fine': None,
        'N_samples': tn['N_samples'],
        'network_fn': NetworkFn(nerf),
        'use_viewdirs': tn['use_viewdirs'],
        'white_bkgd': tn['white_bkgd'],
        'raw_noise_std': tn['raw_noise_std'],
        'near': 2.,
        'far': 6.,
        'ndc': False,
    }


def pose_spherical(theta, phi, radius):
    """Camera-to-world of a camera on a sphere (nerf_helpers.py:22-38,66-71)."""
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    trans = torch.Tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]]).float()
    rphi = torch.Tensor([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0],
                         [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]]).float()
    rth = torch.Tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0],
                        [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]]).float()
    flip = torch.Tensor(np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]))
    return flip @ (rth @ (rphi @ trans))


def get_rays(H, W, K, c2w, device):
    """Pinhole rays, row j / column i, d = dirs . R^T (nerf_helpers.py:134-143)."""
    cols = torch.linspace(0, W - 1, W)
    rows = torch.linspace(0, H - 1, H)
    i = cols.view(1, W).expand(H, W).to(device)
    j = rows.view(H, 1).expand(H, W).to(device)
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    c2w = c2w.to(device)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def _find_module(network_fn):
    if isinstance(network_fn, MLPNeRF):
        return network_fn
    mod = getattr(network_fn, 'module', None)
    if isinstance(mod, MLPNeRF):
        return mod
    for cell in getattr(network_fn, '__closure__', None) or ():
        try:
            if isinstance(cell.cell_contents, MLPNeRF):
                return cell.cell_contents
        except ValueError:
            pass
    raise RuntimeError(
        "render(): network_fn must wrap a ddmi_b200.MLPNeRF (use ddmi_b200.nerf_helpers.get_render_kwargs, "
        "or pass the module itself); an opaque callable cannot be fused")


def sample_depths(rays, N_samples, perturb=0., lindisp=False):
    """z_vals (N_rays, N_samples) of the reference's render_rays (nerf_helpers.py:353-380), op for op: linear in depth or
    (lindisp) in disparity, then optionally stratified: one uniform draw per interval from ``torch.rand`` on the CPU generator,
    as the reference does, moved to the rays' device."""
    n = rays.shape[0]
    bounds = torch.reshape(rays[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=N_samples).to(near.device)
    if not lindisp:
        z_vals = near * (1. - t_vals) + far * (t_vals)
    else:
        z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * (t_vals))
    z_vals = z_vals.expand([n, N_samples])
    if perturb > 0.:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        t_rand = torch.rand(z_vals.shape).to(near.device)
        z_vals = lower + (upper - lower) * t_rand
    return z_vals.contiguous()


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """Hierarchical sampling (utils/nerf_helpers.py:166-209), same signature: bins (N_rays, N_bins) CUDA tensor, weights
    (N_rays, N_bins - 1).  The uniform numbers are drawn exactly as the reference draws them (torch.linspace for `det`,
    torch.rand on the CPU generator otherwise, numpy's seeded generator for `pytest`), the inverse-CDF lookup runs in the
    library's kernel (ddmi_sample_pdf).  The reference's own N_importance > 0 branch of render_rays cannot run
    (it reuses the coarse pass's plane coordinates), so this is exposed as the standalone function."""
    import numpy as np
    if not bins.is_cuda:
        raise RuntimeError("sample_pdf: bins must be a CUDA tensor (ddmi_b200 has no CPU path)")
    lead = list(weights.shape[:-1])
    if det:
        u = torch.linspace(0., 1., steps=N_samples).expand(lead + [N_samples])
    else:
        u = torch.rand(lead + [N_samples])
    if pytest:
        np.random.seed(0)
        u = torch.Tensor(np.broadcast_to(np.linspace(0., 1., N_samples), lead + [N_samples]).copy() if det
                         else np.random.rand(*(lead + [N_samples])))
    dev = bins.device
    b = bins.detach().to(torch.float32).reshape(-1, bins.shape[-1]).contiguous()
    w = weights.detach().to(device=dev, dtype=torch.float32).reshape(-1, weights.shape[-1]).contiguous()
    if w.shape[-1] != b.shape[-1] - 1 or w.shape[0] != b.shape[0]:
        raise RuntimeError(f"sample_pdf: weights {tuple(weights.shape)} must be bins {tuple(bins.shape)} minus one along the last axis")
    uu = u.to(device=dev, dtype=torch.float32).reshape(-1, N_samples).contiguous()
    out = torch.empty_like(uu)
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ddmi_sample_pdf(b.data_ptr(), w.data_ptr(), uu.data_ptr(), b.shape[0], b.shape[1], N_samples,
                                                  out.data_ptr(), _stream_ptr(dev)))
    return out.reshape(lead + [N_samples])


def render_rays_fused(rays, fea, module, N_samples, white_bkgd, return_raw=False, precision=None, perturb=0.,
                      lindisp=False):
    """rays (N,11) [o d near far viewdir]; fea = dict of (B,32,R,R) planes.
    Returns rgb_map (B,N,3) (and raw (B,N,S,4)).  precision: 'f16f8' (tcgen05 kernel, fp16 + FP8-correction operands;
    the default) or 'bf16x3' (same kernel, bf16 hi/lo operands) -- compositing is fused in-kernel when N_samples == 128
    -- or 'fp32' (CUDA-core kernels); env DDMI_B200_PRECISION overrides the default."""
    import os
    import warnings
    from . import packing
    explicit = precision or module.precision or os.environ.get('DDMI_B200_PRECISION')
    precision = explicit or 'f16f8'
    try:
        packed = module.packed_weights(precision)
    except packing.F16F8RangeError:
        if explicit:
            raise
        warnings.warn("ddmi_b200: a weight exceeds the f16f8 operand range (|w| >= 16); using precision='bf16x3'")
        precision = 'bf16x3'
        packed = module.packed_weights(precision)
    planes, sources = [], []
    for k in ('xy', 'yz', 'xz'):
        t = fea[k]
        if not t.is_cuda:
            raise RuntimeError("fea planes must be CUDA tensors (ddmi_b200 has no CPU path)")
        if t.dim() != 4:
            raise RuntimeError(f"fea['{k}'] must be (B,32,R,R), got {tuple(t.shape)}")
        sources.append(t)
        planes.append(t.detach() if _lib.is_channels_last(t.detach()) else t.detach().to(torch.float32).contiguous())
    from .mlp import _check_plane_set
    _check_plane_set(planes, 32, [f"fea['{k}']" for k in ('xy', 'yz', 'xz')])
    module._check_device(planes[0])
    dev = planes[0].device
    b = planes[0].shape[0]
    rays = rays.detach().to(device=dev, dtype=torch.float32).contiguous()
    n = rays.shape[0]
    per_ray = bool(lindisp) or (perturb is not None and perturb > 0.)
    if per_ray:      # the kernel reads a (N_rays, N_samples) depth table instead of near * (1 - t) + far * t
        t_vals = sample_depths(rays, N_samples, perturb or 0., lindisp).to(torch.float32)
    else:
        t_vals = torch.linspace(0., 1., steps=N_samples).to(dev)
    rgb = torch.empty((b, n, 3), device=dev, dtype=torch.float32)
    umma = precision in ('bf16x3', 'f16f8')
    need_raw = return_raw or not (umma and N_samples == 128)
    raw = torch.empty((b, n, N_samples, 4), device=dev, dtype=torch.float32) if need_raw else None
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        if umma:       # one transposition per latent, not per pose (tools/ldm/nerf.py:270 renders a loop of poses)
            keep, arr = module._nhwc_cache.get(planes, sources, st)
        else:
            planes = [p.contiguous() for p in planes]
            keep, arr = planes, _lib.planes_array(planes)
        entry = _lib.lib().ddmi_nerf_render_z if per_ray else _lib.lib().ddmi_nerf_render
        _lib.check(entry(
            arr, b, planes[0].shape[1], 1 if umma else 0, rays.data_ptr(), n, rays.shape[1],
            t_vals.data_ptr(), N_samples, PLANE_EXTENT, module.negative_slope, 1 if white_bkgd else 0,
            _lib.weights_struct(packed), rgb.data_ptr(), raw.data_ptr() if raw is not None else None, st))
        del keep
    return (rgb, raw) if return_raw else rgb


def render(H, W, K, fea, pose, idx, device, chunk=1024 * 32, rays=None, c2w=None, ndc=False,
           near=0., far=1., use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """Render rays -> rgb_map (N_rays, 3).  Signature of nerf_helpers.py:211-279.

    ``chunk`` / ``netchunk`` only bound the reference's memory use and do not
    change results; the fused kernel streams 128-row tiles, so they are accepted
    and ignored.  With planes of batch B > 1 (an extension: B objects seen from
    the same camera) the result is (B, N_rays, 3).
    """
    if ndc or c2w_staticcam is not None:
        raise NotImplementedError("ndc rays / c2w_staticcam are outside the fused decode path")
    if not use_viewdirs:
        raise NotImplementedError("the fused NeRF kernel is built for use_viewdirs=True (in_channels_dir=27)")
    for key, bad in (('N_importance', lambda v: v and v > 0), ('raw_noise_std', lambda v: v and v > 0)):
        if bad(kwargs.get(key, 0)):
            raise NotImplementedError(f"render(): {key}={kwargs[key]!r} is outside the fused decode path "
                                      "(SURVEY.md §8f row 4)")
    module = _find_module(kwargs['network_fn'])
    N_samples = int(kwargs['N_samples'])
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w, device)
    else:
        rays_o, rays_d = rays
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    ray_batch = torch.cat([rays_o, rays_d, near_t, far_t, viewdirs], -1)
    hw_idx = kwargs.get('hw_idx')
    if hw_idx is not None:
        ray_batch = ray_batch[hw_idx]
    rgb = render_rays_fused(ray_batch, fea, module, N_samples, bool(kwargs.get('white_bkgd', False)),
                            perturb=float(kwargs.get('perturb', 0.) or 0.), lindisp=bool(kwargs.get('lindisp', False)))
    return rgb[0] if rgb.shape[0] == 1 else rgb


def render_poses(H, W, K, fea, poses, device, near=0., far=1., **kwargs):
    """Batching across views: the reference renders a Python loop of poses, one `render(...)` call each
    (tools/ldm/nerf.py:266-272).  Rays are independent, so the rays of all V poses go down as ONE launch.
    poses: sequence / tensor of V camera-to-world matrices (3x4 or 4x4).  -> (V, H*W, 3), or (B, V, H*W, 3) for B objects.
    Identical to stacking `render(H, W, K, fea, None, 0, device, c2w=pose, near=near, far=far, use_viewdirs=True, **kwargs)`
    over the poses when perturb == 0 (with perturb the reference draws its stratified offsets per call, i.e. in a different
    order of the same CPU generator)."""
    ro, rd = [], []
    for c2w in poses:
        o, d = get_rays(H, W, K, c2w[:3, :4], device)
        ro.append(o.reshape(-1, 3))
        rd.append(d.reshape(-1, 3))
    V = len(ro)
    rgb = render(H, W, K, fea, None, 0, device, rays=(torch.cat(ro), torch.cat(rd)), near=near, far=far, use_viewdirs=True, **kwargs)
    return rgb.reshape(V, H * W, 3) if rgb.dim() == 2 else rgb.reshape(rgb.shape[0], V, H * W, 3)
