"""GPU parity: the CUDA path (through the drop-in modules -> C ABI) against
(a) the reference's own outputs committed under tests/golden/ and (b) the CPU
oracle on further seeded shapes / edge cases.  Tolerances are the north star's:
max-abs 1e-3 on RGB / density / logits, same occupancy sign on >= 99.99 %."""
import os

import pytest
import torch

import ddmi_b200
from ddmi_b200 import nerf_helpers as nh
from oracle import cases, ddmi_oracle as orc

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'
TOL = 1e-3


def _golden(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'))


def _cuda(x):
    if isinstance(x, (list, tuple)):
        return type(x)(_cuda(t) for t in x)
    if isinstance(x, dict):
        return {k: _cuda(v) for k, v in x.items()}
    return x.to(DEV)


def _sign_agreement(a, b):
    return float(((a > 0) == (b > 0)).float().mean())


def _signs_ok(out, ref):
    """North star: same occupancy sign on >= 99.99 % of the query points.  A flip needs |logit| below the decode error
    (~3e-5 of N(0,0.5)-like logits: about 5 points in 1e5), so on a set of fewer than 1e4 points the RATE is quantised above
    the bar; there one flip is allowed.  In every case a flipped point must sit inside the max-abs tolerance of zero."""
    flips = (out > 0) != (ref > 0)
    n = out.numel()
    allowed = max(1, int(1e-4 * n))
    return int(flips.sum()) <= allowed and (not bool(flips.any()) or float(ref[flips].abs().max()) < TOL)


def test_device_is_blackwell():
    from ddmi_b200 import _lib
    import ctypes
    sm, maj, mnr = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.lib().ddmi_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)))
    assert maj.value == 10 and sm.value >= 100


# ---------------------------------------------------------------- image
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
@pytest.mark.parametrize("tag,kw", [("image_96", dict(batch=2, sizes=(16, 32, 64), res=96)),
                                    ("image_native", dict(batch=1, sizes=(8, 16, 32), res=32))])
def test_image_golden(golden_dir, tag, kw, precision):
    g = _golden(golden_dir, tag)
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    coords, planes, si = cases.image_inputs(**kw)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu()
    assert out.shape == g['out'].shape
    err = float((out - g['out']).abs().max())
    assert err < (2e-5 if precision == 'fp32' else TOL), err


@pytest.mark.parametrize("nsplit", ["off", "full"])
def test_image_golden_shared_memory_operand_kernel(golden_dir, nsplit, monkeypatch):
    """f16f8 defaults to the TMEM-resident-activation kernel (image_umma_kernel<.., TS = 1>); the kernel that keeps the
    activation in shared memory (DDMI_B200_IMAGE_TS=0, the interpreter-driven engine, with and without the N split) stays
    built and must hold the same bar, noise included."""
    monkeypatch.setenv('DDMI_B200_IMAGE_TS', '0')
    monkeypatch.setenv('DDMI_B200_NSPLIT', nsplit)
    g = _golden(golden_dir, 'image_96')
    m = cases.build_module('image').to(DEV)
    m.precision = 'f16f8'
    coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu()
    assert float((out - g['out']).abs().max()) < TOL
    g = _golden(golden_dir, 'image_noise')
    m = cases.build_module('image_noise').to(DEV)
    m.precision = 'f16f8'
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=si, noise=_cuda(cases.image_noise_tensors(2, 96))).cpu()
    assert float((out - g['out']).abs().max()) < TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_image_noise_injection_golden(golden_dir, precision):
    """A checkpoint with non-zero NoiseInjection weights (any trained one): explicit noise tensors, against the REFERENCE
    run with the same tensors handed to NoiseInjection.forward (oracle/make_golden.py)."""
    g = _golden(golden_dir, 'image_noise')
    m = cases.build_module('image_noise').to(DEV)
    m.precision = precision
    coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    noise = cases.image_noise_tensors(2, 96)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=si, noise=_cuda(noise)).cpu()
    err = float((out - g['out']).abs().max())
    assert err < (3e-5 if precision == 'fp32' else TOL), err
    stacked = m(coords.to(DEV), hdbf=_cuda(planes), si=si, noise=torch.stack(noise).to(DEV)).cpu()
    assert torch.equal(stacked, out)
    plain = cases.build_module('image').to(DEV)
    plain.precision = precision
    assert float((plain(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu() - out).abs().max()) > 0.05   # the noise matters
    with pytest.raises(RuntimeError, match="12 tensors"):
        m(coords.to(DEV), hdbf=_cuda(planes), si=si, noise=_cuda(noise[:5]))


@pytest.mark.parametrize("precision", ["fp32", "f16f8"])
def test_image_noise_injection_seeded_stream(precision):
    """noise=<seed>: the in-kernel Philox stream equals its numpy restatement fed as explicit tensors; ragged tile tail,
    batch > 1; noise=None draws a fresh seed per call (reproducible under torch.manual_seed)."""
    m = cases.build_module('image_noise').to(DEV)
    m.precision = precision
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(7)
    planes = [torch.randn(3, 64, s, s, generator=g) for s in (8, 16, 24)]
    coords = torch.rand(1, 2, 9, 31, generator=g) * 2 - 1
    seed = 0x1234567855AA
    z = orc.philox_noise(seed, 3, 9 * 31)
    ref = orc.image_decode(sd, coords, planes, 0.6, noise=z)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=0.6, noise=seed).cpu()
    assert float((out - ref).abs().max()) < (5e-5 if precision == 'fp32' else TOL)
    torch.manual_seed(99)
    a = m(coords.to(DEV), hdbf=_cuda(planes), si=0.6)
    b = m(coords.to(DEV), hdbf=_cuda(planes), si=0.6)
    torch.manual_seed(99)
    c = m(coords.to(DEV), hdbf=_cuda(planes), si=0.6)
    assert float((a - b).abs().max()) > 0.01 and torch.equal(a, c)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_image_ragged_and_scattered_coords(precision):
    """n_coords not a multiple of the tile; out-of-range coords hit the border clamp."""
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(5)
    planes = [torch.randn(3, 64, s, s, generator=g) for s in (8, 16, 24)]
    coords = (torch.rand(1, 2, 7, 19, generator=g) * 2.4 - 1.2)
    ref = orc.image_decode(sd, coords, planes, 0.7)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=0.7).cpu()
    assert float((out - ref).abs().max()) < (2e-5 if precision == 'fp32' else TOL)


def test_image_style_cache_tracks_si_and_weights():
    m = cases.build_module('image').to(DEV)
    m.precision = 'fp32'
    sd = cases.state_dict32(m)
    coords, planes, _ = cases.image_inputs(batch=1, sizes=(8, 16, 32), res=16)
    a = m(coords.to(DEV), hdbf=_cuda(planes), si=1.0).cpu()
    b = m(coords.to(DEV), hdbf=_cuda(planes), si=0.5).cpu()
    assert float((b - orc.image_decode(sd, coords, planes, 0.5)).abs().max()) < 2e-5
    assert float((a - b).abs().max()) > 1e-3
    m.torgb.bias += 1.0   # in-place on the parameter bumps its version (as optimizers / load_state_dict do)
    c = m(coords.to(DEV), hdbf=_cuda(planes), si=0.5).cpu()
    assert float((c - b - 1.0).abs().max()) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "f16f8"])
def test_image_store_modes(precision):
    """Caller epilogues fused into the output stage: clamp is bit-exact against clamping the kernel's own fp32 output, the uint8
    channels-last mode bit-exact against the reference's expression applied to it, and within 1 LSB of the reference's result."""
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    planes = [p * 3.0 for p in planes]                                  # push part of the signal outside [-1, 1]
    ref = orc.image_decode(cases.state_dict32(m), coords, planes, si)
    f32 = m(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu()
    assert float((f32.abs() > 1).float().mean()) > 0.01                  # the clamp matters on this input
    clamped = m(coords.to(DEV), hdbf=_cuda(planes), si=si, store='clamp').cpu()
    assert torch.equal(clamped, orc.store_clamp(f32))
    u8 = m(coords.to(DEV), hdbf=_cuda(planes), si=si, store='u8').cpu()
    assert u8.dtype == torch.uint8 and u8.shape == (2, 96, 96, 3)
    assert torch.equal(u8, orc.store_u8_channels_last(f32))
    d = (u8.int() - orc.store_u8_channels_last(ref).int()).abs()
    assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.99
    with pytest.raises(ValueError):
        m(coords.to(DEV), hdbf=_cuda(planes), si=si, store='u16')


# ---------------------------------------------------------------- occupancy
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_occupancy_golden(golden_dir, precision):
    g = _golden(golden_dir, 'occupancy')
    m = cases.build_module('occupancy').to(DEV)
    m.precision = precision
    pts, hdbf = cases.occupancy_inputs()
    d = m(pts.to(DEV), _cuda(hdbf))
    assert isinstance(d, torch.distributions.Bernoulli)
    out = d.logits.cpu()
    assert float((out - g['out']).abs().max()) < (2e-5 if precision == 'fp32' else TOL)
    assert _sign_agreement(out, g['out']) >= 0.9999


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_occupancy_shared_points_and_single_point(precision):
    m = cases.build_module('occupancy').to(DEV)
    m.precision = precision
    tol = 2e-5 if precision == 'fp32' else TOL
    sd = cases.state_dict32(m)
    pts, hdbf = cases.occupancy_inputs(batch=3, n=1)
    out = m(pts.to(DEV), _cuda(hdbf)).logits.cpu()
    assert float((out - orc.occupancy_logits(sd, pts, hdbf)).abs().max()) < tol
    pts, hdbf = cases.occupancy_inputs(batch=2, n=777)
    shared = pts[:1]
    out = m(shared.to(DEV).expand(2, -1, -1), _cuda(hdbf)).logits.cpu()
    ref = orc.occupancy_logits(sd, shared.expand(2, -1, -1).contiguous(), hdbf)
    assert float((out - ref).abs().max()) < tol


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_occupancy_dense_grid_chunks_like_eval_points(precision):
    """Generator3D.eval_points call shape: 1.1*make_3d_grid, split into chunks, mlp(pi[None], c).logits."""
    m = cases.build_module('occupancy').to(DEV)
    m.precision = precision
    sd = cases.state_dict32(m)
    _, hdbf = cases.occupancy_inputs(batch=1, n=1)
    p = 1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (20,) * 3)
    c = _cuda(hdbf)
    got = torch.cat([m(pi[None].to(DEV), c).logits.squeeze(0).cpu() for pi in torch.split(p, 3000)])
    ref = orc.occupancy_logits(sd, p[None], hdbf)[0]
    assert float((got - ref).abs().max()) < TOL
    assert _signs_ok(got, ref)


# ---------------------------------------------------------------- video
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_video_golden(golden_dir, precision):
    g = _golden(golden_dir, 'video')
    m = cases.build_module('video').to(DEV)
    m.precision = precision
    coords, hdbf = cases.video_inputs()
    out = m(_cuda(coords), _cuda(hdbf)).cpu()
    assert out.shape == g['out'].shape
    assert float((out - g['out']).abs().max()) < (2e-5 if precision == 'fp32' else TOL)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_video_batch_and_anisotropic(precision):
    m = cases.build_module('video').to(DEV)
    m.precision = precision
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(9)
    T, H, W = 3, 10, 14
    xy = [torch.randn(2, 64, H // s, W // s, generator=g) for s in (2, 1, 1)]
    yt = [torch.randn(2, 64, T, H // s, generator=g) for s in (2, 1, 1)]
    xt = [torch.randn(2, 64, T, W // s, generator=g) for s in (2, 1, 1)]
    coords = ddmi_b200.convert_to_coord_format_3d(1, H, W, T, hstart=-.9, hend=.9, wstart=-.8, wend=.8, tstart=-.5, tend=.5)
    ref = orc.video_decode(sd, coords, (xy, yt, xt))
    out = m(_cuda(coords), _cuda((xy, yt, xt))).cpu()
    assert float((out - ref).abs().max()) < (2e-5 if precision == 'fp32' else TOL)


@pytest.mark.parametrize("precision", ["fp32", "f16f8"])
def test_video_store_modes(precision):
    m = cases.build_module('video').to(DEV)
    m.precision = precision
    coords, hdbf = cases.video_inputs()
    hdbf = tuple([p * 4.0 for p in axis] for axis in hdbf)
    ref = orc.video_decode(cases.state_dict32(m), coords, hdbf)
    f32 = m(_cuda(coords), _cuda(hdbf)).cpu()
    clamped = m(_cuda(coords), _cuda(hdbf), store='clamp').cpu()
    assert torch.equal(clamped, orc.store_clamp(f32))
    u8 = m(_cuda(coords), _cuda(hdbf), store='u8').cpu()
    b, c, t, h, w = f32.shape
    assert u8.dtype == torch.uint8 and u8.shape == (b, t, h, w, c)
    assert torch.equal(u8, orc.store_u8_channels_last(f32))
    d = (u8.int() - orc.store_u8_channels_last(ref).int()).abs()
    assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.99


# ---------------------------------------------------------------- nerf
def test_nerf_mlp_golden(golden_dir):
    g = _golden(golden_dir, 'nerf_mlp')
    m = cases.build_module('nerf').to(DEV)
    x = cases.nerf_mlp_inputs()
    out = m(x.to(DEV)).cpu()
    assert float((out - g['out']).abs().max()) < 2e-5
    sig = m(x[:, :159].to(DEV), sigma_only=True).cpu()
    assert float((sig[:, 0] - g['out'][:, 3]).abs().max()) < 2e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_nerf_render_golden(golden_dir, precision):
    g = _golden(golden_dir, 'nerf_render')
    m = cases.build_module('nerf').to(DEV)
    m.precision = precision
    res, K, fea, c2w = cases.nerf_inputs()
    e1, _ = nh.get_embedder(10, 0)
    e2, _ = nh.get_embedder(4, 0)
    kw = nh.get_render_kwargs(cases.NERF_CFG, m, e1, e2)
    rgb = nh.render(res, res, K, _cuda(fea), None, 0, DEV, chunk=4096, c2w=c2w, verbose=True, retraw=True,
                    hw_idx=None, **kw).cpu()
    assert rgb.shape == g['out'].shape == (res * res, 3)
    assert float((rgb - g['out']).abs().max()) < (1e-4 if precision == 'fp32' else TOL)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_nerf_render_batch_and_ray_subset(precision):
    m = cases.build_module('nerf').to(DEV)
    sd = cases.state_dict32(m)
    tol = 1e-4 if precision == 'fp32' else TOL
    g = torch.Generator().manual_seed(11)
    fea = {k: torch.randn(2, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')}
    gold_rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][100:165]
    rgb, raw = nh.render_rays_fused(gold_rays.to(DEV), _cuda(fea), m, 48, True, return_raw=True, precision=precision)
    for b in range(2):
        fb = {k: v[b:b + 1] for k, v in fea.items()}
        ref, ref_raw = orc.nerf_render_rays(sd, gold_rays, fb, 48, True, return_raw=True)
        assert float((raw[b].cpu() - ref_raw).abs().max()) < tol
        assert float((rgb[b].cpu() - ref).abs().max()) < tol


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
@pytest.mark.parametrize("n_obj,n_rays", [(1, 37), (3, 64)])
def test_nerf_render_fused_compositing_128_samples(n_obj, n_rays, precision):
    """N_samples == 128: one tile == one ray, compositing runs inside the tcgen05 kernel (no raw round trip)."""
    m = cases.build_module('nerf').to(DEV)
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(13)
    fea = {k: torch.randn(n_obj, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')}
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][300:300 + n_rays]
    rgb = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, True, precision=precision)
    rgb2, raw = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, True, return_raw=True, precision=precision)
    assert float((rgb - rgb2).abs().max()) == 0.0
    for b in range(n_obj):
        fb = {k: v[b:b + 1] for k, v in fea.items()}
        ref, ref_raw = orc.nerf_render_rays(sd, rays, fb, 128, True, return_raw=True)
        assert float((raw[b].cpu() - ref_raw).abs().max()) < TOL
        assert float((rgb[b].cpu() - ref).abs().max()) < TOL
    black = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, False, precision=precision)
    assert float((rgb - black).min()) >= -1e-6      # the white background only adds (1 - acc) >= 0


def test_unsupported_render_options_raise():
    m = cases.build_module('nerf').to(DEV)
    res, K, fea, c2w = cases.nerf_inputs()
    e1, _ = nh.get_embedder(10, 0)
    e2, _ = nh.get_embedder(4, 0)
    kw = nh.get_render_kwargs(cases.NERF_CFG, m, e1, e2)
    for key, val in (('N_importance', 64), ('raw_noise_std', 1.0)):
        kw2 = dict(kw)
        kw2[key] = val
        with pytest.raises(NotImplementedError):
            nh.render(res, res, K, _cuda(fea), None, 0, DEV, c2w=c2w, **kw2)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
@pytest.mark.parametrize("tag,perturb,lindisp", [("nerf_render_perturb", 1.0, False), ("nerf_render_lindisp", 0., True),
                                                 ("nerf_render_perturb_lindisp", 1.0, True)])
def test_nerf_render_stratified_and_lindisp_golden(golden_dir, tag, perturb, lindisp, precision):
    """render(..., perturb=1 / lindisp=True): the host side builds the reference's per-ray depth table (same CPU torch.rand
    draw under the same seed) and the kernels read it through ddmi_nerf_render_z."""
    g = _golden(golden_dir, tag)
    m = cases.build_module('nerf').to(DEV)
    m.precision = precision
    res, K, fea, c2w = cases.nerf_inputs()
    e1, _ = nh.get_embedder(10, 0)
    e2, _ = nh.get_embedder(4, 0)
    kw = nh.get_render_kwargs(cases.NERF_CFG, m, e1, e2)
    kw.update(perturb=perturb, lindisp=lindisp)
    torch.manual_seed(cases.NERF_PERTURB_SEED)
    rgb = nh.render(res, res, K, _cuda(fea), None, 0, DEV, chunk=4096, c2w=c2w, **kw).cpu()
    assert float((rgb - g['out']).abs().max()) < (1e-4 if precision == 'fp32' else TOL)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
def test_nerf_render_with_a_real_leaky_slope(precision):
    """The reference's nn.LeakyReLU(True) has slope 1.0 (identity), which the tcgen05 kernel special-cases; a module whose
    activations really leak (slope 0.2) goes through the general path."""
    m = cases.build_module('nerf').to(DEV)
    for mod in m.modules():
        if isinstance(mod, torch.nn.LeakyReLU):
            mod.negative_slope = 0.2
    assert m.negative_slope == 0.2
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(23)
    fea = {k: torch.randn(1, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')}
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][40:81]
    rgb, raw = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, True, return_raw=True, precision=precision)
    ref, ref_raw = orc.nerf_render_rays(sd, rays, fea, 128, True, slope=0.2, return_raw=True)
    tol = 1e-4 if precision == 'fp32' else TOL
    assert float((raw[0].cpu() - ref_raw).abs().max()) < tol
    assert float((rgb[0].cpu() - ref).abs().max()) < tol


def test_nerf_stratified_fused_compositing_128_samples():
    """per-ray depth table + in-kernel compositing (N_samples == 128) against the oracle on the same draw."""
    m = cases.build_module('nerf').to(DEV)
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(17)
    fea = {k: torch.randn(1, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')}
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][500:541]
    torch.manual_seed(5)
    rgb = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, True, perturb=1.0)
    torch.manual_seed(5)
    ref = orc.nerf_render_rays(sd, rays, fea, 128, True, perturb=1.0)
    assert float((rgb[0].cpu() - ref).abs().max()) < TOL


def test_sample_pdf_golden(golden_dir):
    """nerf_helpers.sample_pdf (same signature as the reference's) against the reference's outputs: det=True and pytest=True."""
    g = _golden(golden_dir, 'sample_pdf')
    bins, w = g['bins'].to(DEV), g['weights'].to(DEV)
    # The reference's lookup is DIScontinuous at bins whose pdf is below its 1e-5 guard (t is then taken against a unit
    # denominator), so a last-bit difference in the normalising sum (tree vs vectorised order) may move a sample that sits on
    # such an edge to the neighbouring bin -- and with det=True the last sample of EVERY ray (u = 1.0) sits exactly on the end of
    # the cdf, whose last bit decides the bin.  So: at most about one moved sample per ray (measured: 41 of 38400), everything
    # else within 2e-5, and a moved sample stays within one bin width.
    def close(a, ref):
        d = (a - ref).abs()
        width = float((g['bins'][:, 1:] - g['bins'][:, :-1]).max())
        moved = int((d >= 2e-5).sum())
        return moved <= a.shape[0] // 2 and float(d.max()) <= width
    out = nh.sample_pdf(bins, w, 128, det=True).cpu()
    assert out.shape == g['out'].shape and close(out, g['out'])
    rnd = nh.sample_pdf(bins, w, 96, det=False, pytest=True).cpu()
    assert close(rnd, g['out_pytest'])
    # samples fall inside their ray's bin range and follow the weights: a ray with one dominant bin puts its samples there
    assert bool((out >= g['bins'][:, :1] - 1e-6).all()) and bool((out <= g['bins'][:, -1:] + 1e-6).all())
    w1 = torch.zeros(1, 62)
    w1[0, 30] = 1.0
    one = nh.sample_pdf(bins[:1], w1.to(DEV), 64, det=True).cpu()
    inside = (one >= g['bins'][0, 30]) & (one <= g['bins'][0, 31])
    assert float(inside.float().mean()) > 0.95


# ---------------------------------------------------------------- tcgen05 bring-up
@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (64, 64), (16, 256), (128, 32)])
def test_umma_selftest(N, K):
    from ddmi_b200 import _lib
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    d = torch.full((128, N), float('nan'), device=DEV)
    _lib.check(_lib.lib().ddmi_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K,
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    assert float((d.double() - ref).abs().max()) < 2e-3 * float(ref.abs().max())
    assert float((d.double() - ref).abs().max()) < 1e-3


@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (64, 64), (16, 256), (128, 32)])
def test_umma2_selftest(N, K):
    """CTA-pair MMA (cta_group::2): M = 256 across a 2-CTA cluster, B split N/2 per CTA."""
    from ddmi_b200 import _lib
    g = torch.Generator().manual_seed(N * 1000 + K + 7)
    a = torch.randn(256, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    d = torch.full((256, N), float('nan'), device=DEV)
    _lib.check(_lib.lib().ddmi_selftest_umma2(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K,
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    assert float((d.double() - ref).abs().max()) < 1e-3


@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (64, 64), (16, 256), (128, 32)])
def test_f16f8_selftest(N, K):
    """fp16 main term + two FP8 correction terms (kind::f8f6f4, e5m2 activations x e4m3 weights) into one accumulator:
    checked against the exact product and, tightly, against a torch emulation of the same operand rounding."""
    from ddmi_b200 import _lib
    g = torch.Generator().manual_seed(N * 1000 + K + 13)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(N, K, generator=g) * 0.1
    d = torch.full((128, N), float('nan'), device=DEV)
    ad, bd = a.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ddmi_selftest_f16f8(ad.data_ptr(), bd.data_ptr(), d.data_ptr(), N, K,
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    S = 4096.0
    q8 = lambda x: x.clamp(-448, 448).to(torch.float8_e4m3fn).double()           # weight side
    q5 = lambda x: x.clamp(-57344, 57344).to(torch.float8_e5m2).double()         # activation side
    a16 = a.to(torch.float16)
    w16 = (b * S).to(torch.float16)
    a8 = (a16.view(torch.int16) & -256).view(torch.float16).double()             # e5m2(a) = the high byte of fp16(a) (truncation)
    emu = (a16.double() @ w16.double().t() + q5((a - a16.float()) * S) @ q8(b).t()
           + a8 @ q8(b * S - w16.float()).t()) / S
    ref = a.double() @ b.double().t()
    got = d.double().cpu()
    assert float((got - emu).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("x,y,c", [(0, 0, 0), (36, 5, 64), (100, 127, 64), (124, 126, 0)])
def test_tma_plane_window_selftest(x, y, c, variant):
    """cp.async.bulk.tensor.3d over an NCHW fp32 plane batch: a 64 x 2 x 64-channel box at (x, y, channel c); texels outside the
    plane read as 0 (windows at the right / bottom edge)."""
    from ddmi_b200 import _lib
    B, C, H, W = 2, 64, 128, 128
    plane = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(3)).to(DEV)
    out = torch.full((64, 2, 64), float('nan'), device=DEV)
    mapd = torch.zeros(64, dtype=torch.int32, device=DEV)
    _lib.check(_lib.lib().ddmi_selftest_tma(plane.data_ptr(), B, C, H, W, x, y, c, variant, mapd.data_ptr(), out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = torch.zeros(64, 2, 64, device=DEV)
    xs, ys = min(64, W - x), min(2, H - y)
    ref[:, :ys, :xs] = plane.reshape(B * C, H, W)[c:c + 64, y:y + ys, x:x + xs]
    assert torch.equal(out, ref)


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
@pytest.mark.parametrize("R,sizes,span", [(512, (64, 128, 256), 1.0), (1024, (64, 128, 256), 1.0), (320, (32, 64, 128), 1.25),
                                          (256, (18, 36, 72), 1.0), (200, (64, 128, 256), 1.0)])
def test_image_tma_staged_windows_match_the_direct_gather(R, sizes, span, precision, monkeypatch):
    """Regular grids take their plane windows through TMA (decode_umma.cu: patch_plan / gather_staged); the blend order is the
    direct gather's, so both paths must agree BIT FOR BIT -- on grids that reach past [-1, 1] (border clamp), on plane widths
    TMA cannot map (18: not a multiple of 4 -> direct), and on row lengths that make tiles straddle image rows (200)."""
    g = torch.Generator().manual_seed(R)
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    planes = _cuda([torch.randn(2, 64, s, s, generator=g) for s in sizes])
    e = span * (R - 1) / R
    coords = ddmi_b200.convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(DEV)
    monkeypatch.setenv("DDMI_B200_NO_TMA_PATCH", "0")
    staged = m(coords, hdbf=planes, si=0.5)
    monkeypatch.setenv("DDMI_B200_NO_TMA_PATCH", "1")
    direct = m(coords, hdbf=planes, si=0.5)
    assert torch.equal(staged, direct)
    m.precision = 'fp32'
    assert float((staged - m(coords, hdbf=planes, si=0.5)).abs().max()) < TOL


# ---------------------------------------------------------------- size-independent properties at larger shapes
@pytest.mark.parametrize("scheme", ["bf16x3", "f16f8"])
def test_image_full_size_cross_check_and_crop_independence(scheme):
    """BASELINE configs[0]-size grid (256^2) and an up-sampled 512^2 grid: the tcgen05 kernel agrees with the fp32
    CUDA-core kernel to 1e-3, and decoding a crop of the coordinates reproduces the crop of the full decode
    (results do not depend on how rows fall into tiles / CTA pairs)."""
    m = cases.build_module('image').to(DEV)
    g = torch.Generator().manual_seed(21)
    planes = _cuda([torch.randn(3, 64, s, s, generator=g) for s in (64, 128, 256)])
    for R in (256, 512):
        e = (R - 1) / R
        coords = ddmi_b200.convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(DEV)
        si = ddmi_b200.get_scale_injection(R)
        m.precision = scheme
        full = m(coords, hdbf=planes, si=si)
        m.precision = 'fp32'
        exact = m(coords, hdbf=planes, si=si)
        err = float((full - exact).abs().max())
        print(f"image {R}x{R} {scheme} vs fp32 kernel: max abs {err:.3e}")
        assert err < TOL
        m.precision = scheme
        crop = m(coords[:, :, 37:101, 5:R - 3], hdbf=planes, si=si)
        assert torch.equal(crop, full[:, :, 37:101, 5:R - 3])
        one = m(coords, hdbf=[p[1:2] for p in planes], si=si)
        assert torch.equal(one[0], full[1])


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_occupancy_full_grid_sign_agreement(precision):
    """One item on the full 128^3 grid + 100k random points (config C4 per item): tcgen05 vs fp32 kernel."""
    m = cases.build_module('occupancy').to(DEV)
    _, hdbf = cases.occupancy_inputs(batch=1, n=1)
    g = torch.Generator().manual_seed(22)
    p = torch.cat([1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (128,) * 3),
                   (torch.rand(100000, 3, generator=g) - 0.5) * 1.1]).to(DEV)
    c = _cuda(hdbf)
    m.precision = precision
    a = m(p[None], c).logits
    m.precision = 'fp32'
    b = m(p[None], c).logits
    assert float((a - b).abs().max()) < TOL
    assert _sign_agreement(a, b) >= 0.9999


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_video_batch_independence_and_cross_check(precision):
    m = cases.build_module('video').to(DEV)
    g = torch.Generator().manual_seed(23)
    T, R = 8, 64
    xy = _cuda([torch.randn(3, 64, s, s, generator=g) for s in (16, 32, 64)])
    yt = _cuda([torch.randn(3, 64, T, s, generator=g) for s in (16, 32, 64)])
    xt = _cuda([torch.randn(3, 64, T, s, generator=g) for s in (16, 32, 64)])
    e, et = (R - 1) / R, (T - 1) / T
    coords = _cuda(ddmi_b200.convert_to_coord_format_3d(1, R, R, T, hstart=-e, hend=e, wstart=-e, wend=e, tstart=-et, tend=et))
    m.precision = precision
    full = m(coords, (xy, yt, xt))
    one = m(coords, ([p[2:3] for p in xy], [p[2:3] for p in yt], [p[2:3] for p in xt]))
    assert torch.equal(one[0], full[2])
    m.precision = 'fp32'
    exact = m(coords, (xy, yt, xt))
    assert float((full - exact).abs().max()) < TOL


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_nerf_full_view_cross_check(precision):
    """A full 64x64-ray view at 128 samples (fused compositing) vs the fp32 kernels (separate compositing kernel)."""
    m = cases.build_module('nerf').to(DEV)
    res, K, fea, c2w = cases.nerf_inputs(res=64, theta=110.0)
    ro, rd = nh.get_rays(res, res, K, c2w, 'cpu')
    vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
    rays = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), 2. * torch.ones(res * res, 1), 6. * torch.ones(res * res, 1), vd], -1).to(DEV)
    a = nh.render_rays_fused(rays, _cuda(fea), m, 128, True, precision=precision)
    b = nh.render_rays_fused(rays, _cuda(fea), m, 128, True, precision='fp32')
    assert float((a - b).abs().max()) < TOL
    assert float(a.max() - a.min()) > 0.2


# ---------------------------------------------------------------- the BASELINE configs themselves against the oracle
@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
@pytest.mark.parametrize("R,rows", [(1024, [(0, 8), (500, 532), (1016, 1024)]), (2048, [(0, 4), (1337, 1353), (2044, 2048)])])
def test_image_bench_config_vs_oracle(R, rows, precision):
    """configs[1] (the bench line): AFHQ-shape planes 64^2/128^2/256^2, the 1024^2 (si = 0.25) and 2048^2 (si = 0.125)
    query grids.  The oracle decodes row bands of the full grid (top border, interior, bottom border); the kernel
    decodes the FULL grid in one launch (so rows fall into tiles / CTA pairs as in the bench) and must match on the bands."""
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(41)
    planes = [torch.randn(1, 64, s, s, generator=g) for s in (64, 128, 256)]
    e = (R - 1) / R
    coords = ddmi_b200.convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e)
    si = ddmi_b200.get_scale_injection(R)
    assert si == 256 / R
    full = m(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu()
    for r0, r1 in rows:
        ref = orc.image_decode(sd, coords[:, :, r0:r1], planes, si)
        err = float((full[:, :, r0:r1] - ref).abs().max())
        print(f"image {R}x{R} rows {r0}:{r1} {precision}: max abs vs oracle {err:.3e}")
        assert err < TOL


def test_video_config_shape_vs_oracle():
    """configs[2], one item: 256x256x16 volume on 64/128/256 planes, default precision, oracle chunked over rows."""
    m = cases.build_module('video').to(DEV)
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(42)
    xy = [torch.randn(1, 64, s, s, generator=g) for s in (64, 128, 256)]
    yt = [torch.randn(1, 64, 16, s, generator=g) for s in (64, 128, 256)]
    xt = [torch.randn(1, 64, 16, s, generator=g) for s in (64, 128, 256)]
    coords = ddmi_b200.convert_to_coord_format_3d(1, 256, 256, 16, hstart=-255 / 256, hend=255 / 256, wstart=-255 / 256,
                                                  wend=255 / 256, tstart=-15 / 16, tend=15 / 16)
    out = m(_cuda(coords), _cuda((xy, yt, xt))).cpu()
    assert out.shape == (1, 3, 16, 256, 256)
    worst = 0.0
    for h0 in range(0, 256, 64):
        sub = {'xy': coords['xy'][:, :, h0:h0 + 64], 'yt': coords['yt'][:, :, :, h0:h0 + 64], 'xt': coords['xt']}
        ref = orc.video_decode(sd, sub, (xy, yt, xt), thw=(16, 64, 256))
        worst = max(worst, float((out[:, :, :, h0:h0 + 64] - ref).abs().max()))
    print(f"video 256x256x16 vs oracle: max abs {worst:.3e}")
    assert worst < TOL


def test_occupancy_config_shape_vs_oracle():
    """configs[3], one item: every 4th z-slab of the 128^3 grid + 100k random points against the oracle."""
    m = cases.build_module('occupancy').to(DEV)
    sd = cases.state_dict32(m)
    _, hdbf = cases.occupancy_inputs(batch=1, n=1)
    g = torch.Generator().manual_seed(43)
    grid = 1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (128,) * 3)
    p = torch.cat([grid.reshape(128, 128 * 128, 3)[::4].reshape(-1, 3), (torch.rand(100000, 3, generator=g) - 0.5) * 1.1])
    out = m(p[None].to(DEV), _cuda(hdbf)).logits.cpu()[0]
    ref = torch.cat([orc.occupancy_logits(sd, pi[None], hdbf)[0] for pi in torch.split(p, 100000)])
    err = float((out - ref).abs().max())
    print(f"occupancy 128^3 slabs + 100k random vs oracle: max abs {err:.3e}, sign agreement {_sign_agreement(out, ref):.6f}")
    assert err < TOL and _sign_agreement(out, ref) >= 0.9999


def test_nerf_config_shape_vs_oracle():
    """configs[4], one object: a 128x128-ray view x 128 samples, compositing fused, against the oracle (chunked over rays)."""
    m = cases.build_module('nerf').to(DEV)
    sd = cases.state_dict32(m)
    res, K, fea, c2w = cases.nerf_inputs(res=128, theta=75.0)
    ro, rd = nh.get_rays(res, res, K, c2w, 'cpu')
    vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
    rays = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), 2. * torch.ones(res * res, 1), 6. * torch.ones(res * res, 1), vd], -1)
    rgb = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, 128, True)[0].cpu()
    ref = torch.cat([orc.nerf_render_rays(sd, r, fea, 128, True) for r in torch.split(rays, 2048)])
    err = float((rgb - ref).abs().max())
    print(f"nerf 128x128x128 vs oracle: max abs {err:.3e}")
    assert err < TOL and float(ref.max() - ref.min()) > 0.2


# ---------------------------------------------------------------- operand range of the f16f8 scheme
@pytest.mark.parametrize("scale", [10.0, 50.0, 300.0])
def test_f16f8_holds_its_accuracy_on_large_signals(scale):
    """Planes far from N(0,1) (x10, x50, x300: hidden activations up to ~2000 / ~1e4): the f16f8 scheme's correction operands
    must not saturate (the activation side is e5m2, range 57344; include/ddmi_b200.h states the limit), so the error stays at
    the same fraction of |out|max as on unit-scale planes."""
    m = cases.build_module('image').to(DEV)
    m.precision = 'f16f8'
    coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    planes = [p * scale for p in planes]
    ref = orc.image_decode(cases.state_dict32(m), coords, planes, si)
    out = m(coords.to(DEV), hdbf=_cuda(planes), si=si).cpu()
    rel = float((out - ref).abs().max()) / float(ref.abs().max())
    print(f"image planes x{scale:g}: |out|max {float(ref.abs().max()):.1f}, max err / |out|max {rel:.3e}")
    assert rel < 3e-4

    mo = cases.build_module('occupancy').to(DEV)
    mo.precision = 'f16f8'
    pts, hdbf = cases.occupancy_inputs(batch=1, n=20000)
    hdbf = tuple([p * scale for p in axis] for axis in hdbf)
    ref = orc.occupancy_logits(cases.state_dict32(mo), pts, hdbf)
    out = mo(pts.to(DEV), _cuda(hdbf)).logits.cpu()
    rel = float((out - ref).abs().max()) / float(ref.abs().max())
    print(f"occupancy planes x{scale:g}: |logit|max {float(ref.abs().max()):.1f}, max err / |logit|max {rel:.3e}")
    assert rel < 3e-4

    mv = cases.build_module('video').to(DEV)
    mv.precision = 'f16f8'
    cv, hv = cases.video_inputs()
    hv = tuple([p * scale for p in axis] for axis in hv)
    ref = orc.video_decode(cases.state_dict32(mv), cv, hv)
    out = mv(_cuda(cv), _cuda(hv)).cpu()
    rel = float((out - ref).abs().max()) / float(ref.abs().max())
    print(f"video planes x{scale:g}: |out|max {float(ref.abs().max()):.1f}, max err / |out|max {rel:.3e}")
    assert rel < 3e-4


# ---------------------------------------------------------------- host-side caches and input validation
def test_data_writes_and_ema_style_swaps_reach_the_kernel():
    m = cases.build_module('image').to(DEV)
    coords, planes, si = cases.image_inputs(batch=1, sizes=(8, 16, 32), res=16)
    a = m(coords.to(DEV), hdbf=_cuda(planes), si=si)
    m.torgb.bias.data.copy_(m.torgb.bias.data + 1.0)              # LitEma.copy_to idiom: no version bump
    b = m(coords.to(DEV), hdbf=_cuda(planes), si=si)
    assert float((b - a - 1.0).abs().max()) < 1e-5


def test_mismatched_planes_raise():
    m = cases.build_module('image').to(DEV)
    coords, planes, si = cases.image_inputs(batch=2, sizes=(8, 16, 32), res=16)
    bad = _cuda(planes)
    bad[1] = bad[1][:1]
    with pytest.raises(RuntimeError, match="batch"):
        m(coords.to(DEV), hdbf=bad, si=si)
    mo = cases.build_module('occupancy').to(DEV)
    pts, hdbf = cases.occupancy_inputs(batch=2, n=64)
    hb = _cuda(hdbf)
    hb[2][1] = hb[2][1][:, :32]
    with pytest.raises(RuntimeError, match="channels"):
        mo(pts.to(DEV), hb)
    mn = cases.build_module('nerf').to(DEV)
    res, K, fea, c2w = cases.nerf_inputs(res=8)
    fb = _cuda(fea)
    fb['yz'] = fb['yz'].repeat(2, 1, 1, 1)
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][:16]
    with pytest.raises(RuntimeError, match="batch"):
        nh.render_rays_fused(rays.to(DEV), fb, mn, 128, True)


def test_nerf_mlp_forward_ignores_the_tensor_core_precision_override(monkeypatch):
    """DDMI_B200_PRECISION selects the default of the fused decoders; MLPNeRF.forward (rows in, rows out) has an fp32 kernel only."""
    monkeypatch.setenv('DDMI_B200_PRECISION', 'f16f8')
    m = cases.build_module('nerf').to(DEV)
    x = cases.nerf_mlp_inputs(n=200)
    ref = orc.nerf_mlp(cases.state_dict32(m), x)
    assert float((m(x.to(DEV)).cpu() - ref).abs().max()) < 2e-5


def test_channels_last_plane_cache_tracks_the_planes():
    """eval_points-style chunk loop: the channels-last copies are made once per latent; an in-place update of a plane (version
    bump) or a new tensor invalidates them."""
    m = cases.build_module('occupancy').to(DEV)
    sd = cases.state_dict32(m)
    pts, hdbf = cases.occupancy_inputs(batch=1, n=3000)
    c = _cuda(hdbf)
    a = m(pts.to(DEV), c).logits.cpu()
    key = m._nhwc_cache.key
    b = torch.cat([m(pi.to(DEV), c).logits.cpu() for pi in torch.split(pts, 1000, dim=1)], dim=1)
    assert m._nhwc_cache.key == key and torch.equal(a, b)
    c[0][2].mul_(0.5)                                              # in place: same storage, new version
    hdbf[0][2].mul_(0.5)
    d = m(pts.to(DEV), c).logits.cpu()
    assert m._nhwc_cache.key != key
    assert float((d - orc.occupancy_logits(sd, pts, hdbf)).abs().max()) < TOL
    with torch.inference_mode():                                   # inference tensors have no version counter: never cached
        ci = _cuda(hdbf)
        e = m(pts.to(DEV), ci).logits.cpu()
    assert float((e - d).abs().max()) == 0.0


# ---------------------------------------------------------------- ragged / degenerate shapes
def test_occupancy_empty_and_per_item_points():
    m = cases.build_module('occupancy').to(DEV)
    sd = cases.state_dict32(m)
    pts, hdbf = cases.occupancy_inputs(batch=3, n=130)          # 130 = one full tile + 2 rows; distinct points per item
    out = m(pts.to(DEV), _cuda(hdbf)).logits.cpu()
    assert float((out - orc.occupancy_logits(sd, pts, hdbf)).abs().max()) < TOL
    empty = m(pts[:, :0].to(DEV), _cuda(hdbf)).logits
    assert tuple(empty.shape) == (3, 0)


def test_video_ragged_volume():
    """W not a multiple of the tile, odd tile count (the pair's second CTA decodes a duplicate and stores nothing)."""
    m = cases.build_module('video').to(DEV)
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(31)
    T, H, W = 3, 7, 19                                              # 399 voxels -> 4 tiles; x3 items -> 12 tiles... x1 item -> odd
    xy = [torch.randn(1, 64, 5, 6, generator=g), torch.randn(1, 64, 7, 9, generator=g), torch.randn(1, 64, H, W, generator=g)]
    yt = [torch.randn(1, 64, 2, 5, generator=g), torch.randn(1, 64, 3, 7, generator=g), torch.randn(1, 64, T, H, generator=g)]
    xt = [torch.randn(1, 64, 2, 6, generator=g), torch.randn(1, 64, 3, 9, generator=g), torch.randn(1, 64, T, W, generator=g)]
    coords = ddmi_b200.convert_to_coord_format_3d(1, H, W, T, hstart=-1.2, hend=1.1, wstart=-.9, wend=1.3, tstart=-1, tend=1)
    ref = orc.video_decode(sd, coords, (xy, yt, xt))
    out = m(_cuda(coords), _cuda((xy, yt, xt))).cpu()
    assert out.shape == ref.shape == (1, 3, T, H, W)
    assert float((out - ref).abs().max()) < TOL


def test_single_cta_kernels_of_the_resnet_decoders(monkeypatch):
    """DDMI_B200_CTA_PAIR=0 (bf16x3 only): the single-CTA instantiations of the occupancy / video kernels run the same programs
    (PE operands on their own barrier, R1's video pieces parked in H, net_p on the tensor core) -- against the oracle."""
    monkeypatch.setenv('DDMI_B200_CTA_PAIR', '0')
    m = cases.build_module('occupancy').to(DEV)
    m.precision = 'bf16x3'
    pts, hdbf = cases.occupancy_inputs(batch=2, n=777)
    out = m(pts.to(DEV), _cuda(hdbf)).logits.cpu()
    ref = orc.occupancy_logits(cases.state_dict32(m), pts, hdbf)
    assert float((out - ref).abs().max()) < TOL
    mv = cases.build_module('video').to(DEV)
    mv.precision = 'bf16x3'
    coords, hv = cases.video_inputs()
    outv = mv(_cuda(coords), _cuda(hv)).cpu()
    assert float((outv - orc.video_decode(cases.state_dict32(mv), coords, hv)).abs().max()) < TOL


def test_full_size_table_paths_are_bit_identical_to_the_direct_paths(monkeypatch):
    """BASELINE configs[2] / [3] at full size, two items each: the video feature tables against per-voxel gathers and the 128^3
    lattice query against the expanded point list -- equal bit for bit (size-independent property; the oracle comparisons at
    these shapes are test_video_config_shape_vs_oracle / test_occupancy_config_shape_vs_oracle)."""
    g = torch.Generator().manual_seed(71)
    mv = cases.build_module('video').to(DEV)
    hd = _cuda(([torch.randn(2, 64, s, s, generator=g) for s in (64, 128, 256)],
                [torch.randn(2, 64, 16, s, generator=g) for s in (64, 128, 256)],
                [torch.randn(2, 64, 16, s, generator=g) for s in (64, 128, 256)]))
    cv = _cuda(ddmi_b200.convert_to_coord_format_3d(1, 256, 256, 16, hstart=-255 / 256, hend=255 / 256, wstart=-255 / 256,
                                                    wend=255 / 256, tstart=-15 / 16, tend=15 / 16))
    tab = mv(cv, hd)
    monkeypatch.setenv('DDMI_B200_VIDEO_TABLE_MAX', '0')
    assert torch.equal(tab, mv(cv, hd))
    mo = cases.build_module('occupancy').to(DEV)
    ho = _cuda(tuple([torch.randn(2, 64, s, s, generator=g) for s in (16, 32, 64)] for _ in range(3)))
    ax = (1.1 * torch.linspace(-0.5, 0.5, 128)).to(DEV)
    lat = mo.decode_logits_lattice((ax, ax, ax), ho)
    pts = (1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (128,) * 3)).to(DEV)
    assert torch.equal(lat.reshape(2, -1), mo.decode_logits(pts[None], ho))
    assert float(lat.std()) > 0 and float(tab.std()) > 0


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
@pytest.mark.parametrize("shape", [(5, 7, 50), (33, 17, 129), (16, 16, 256)])
def test_occupancy_lattice_matches_the_point_list(shape, precision):
    """decode_logits_lattice (feature tables per (i,j) / (j,k) / (i,k), ddmi_decode_occupancy_lattice) == the expanded point list,
    bit for bit: lattice lines that straddle tiles, unequal axes, axis values beyond the padded box (clamped), batch 2."""
    m = cases.build_module('occupancy').to(DEV)
    m.precision = precision
    g = torch.Generator().manual_seed(61)
    hdbf = _cuda(tuple([torch.randn(2, 64, s, s, generator=g) for s in (16, 32, 64)] for _ in range(3)))
    nx, ny, nz = shape
    axes = [torch.linspace(-0.62, 0.58, nx), torch.linspace(-0.5, 0.5, ny) * 1.1, (torch.rand(nz, generator=g) - 0.5) * 1.3]
    lat = m.decode_logits_lattice([a.to(DEV) for a in axes], hdbf)
    assert lat.shape == (2, nx, ny, nz)
    pts = torch.cartesian_prod(*axes)[None].to(DEV)
    ref = m.decode_logits(pts, hdbf).reshape(2, nx, ny, nz)
    assert torch.equal(lat, ref)
    assert float(lat.std()) > 0
    m.precision = 'fp32'                                   # the exact-arithmetic path expands the lattice itself
    exact = m.decode_logits_lattice([a.to(DEV) for a in axes], hdbf)
    assert float((lat - exact).abs().max()) < TOL


@pytest.mark.parametrize("shape", [(3, 7, 19), (4, 16, 128), (2, 24, 320)])
def test_video_feature_tables_match_the_direct_gather(shape, monkeypatch):
    """f16f8: the operand-format feature tables (one record per distinct (h,w) / (t,h) / (t,w) grid entry, ddmi_decode_video_ws)
    reproduce the per-voxel gather bit for bit -- ragged volumes, tiles that straddle rows, batch > 1, out-of-range coords,
    negative-zero and tiny features (the relu copy is the raw record with negative channels masked)."""
    m = cases.build_module('video').to(DEV)
    m.precision = 'f16f8'
    g = torch.Generator().manual_seed(57)
    T, H, W = shape
    mk = lambda *sz: torch.randn(*sz, generator=g)
    xy = [mk(2, 64, max(H // 4, 2), max(W // 4, 2)), mk(2, 64, max(H // 2, 2), max(W // 2, 2)), mk(2, 64, H, W)]
    yt = [mk(2, 64, max(T // 2, 1), max(H // 4, 2)), mk(2, 64, T, max(H // 2, 2)), mk(2, 64, T, H)]
    xt = [mk(2, 64, max(T // 2, 1), max(W // 4, 2)), mk(2, 64, T, max(W // 2, 2)), mk(2, 64, T, W)]
    xy[2][:, :8] = 0.0
    xy[2][:, 8:16] = -0.0
    xy[2][:, 16:24] *= 1e-9
    coords = _cuda(ddmi_b200.convert_to_coord_format_3d(1, H, W, T, hstart=-1.2, hend=1.1, wstart=-.9, wend=1.3, tstart=-1, tend=1))
    hdbf = _cuda((xy, yt, xt))
    assert ddmi_b200._lib.lib().ddmi_video_workspace_bytes(2, T, H, W, ddmi_b200._lib.PREC_F16F8) == 2 * 3 * (H * W + T * H + T * W) * 256
    tab = m(coords, hdbf)
    monkeypatch.setenv('DDMI_B200_VIDEO_TABLE_MAX', '0')
    direct = m(coords, hdbf)
    assert torch.equal(tab, direct)
    assert float(tab.abs().max()) > 0


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
@pytest.mark.parametrize("n_samples", [1, 100, 192])
def test_nerf_sample_counts_that_straddle_tiles(n_samples, precision):
    m = cases.build_module('nerf').to(DEV)
    sd = cases.state_dict32(m)
    g = torch.Generator().manual_seed(33)
    fea = {k: torch.randn(1, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')}
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'][500:511]
    rgb, raw = nh.render_rays_fused(rays.to(DEV), _cuda(fea), m, n_samples, False, return_raw=True, precision=precision)
    ref, ref_raw = orc.nerf_render_rays(sd, rays, fea, n_samples, False, return_raw=True)
    assert float((raw[0].cpu() - ref_raw).abs().max()) < TOL
    assert float((rgb[0].cpu() - ref).abs().max()) < TOL


# ---------------------------------------------------------------- run-to-run determinism (protocol races would show here)
@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_repeated_decodes_are_bit_identical(precision):
    """Many tiles per CTA pair on every SM, decoded three times: the producer / issuer / epilogue hand-shakes (mbarriers, ring
    re-use, TMEM re-use) leave no room for timing-dependent results, so the outputs must be bit-identical."""
    g = torch.Generator().manual_seed(31)
    m = cases.build_module('image').to(DEV)
    m.precision = precision
    planes = _cuda([torch.randn(6, 64, s, s, generator=g) for s in (32, 64, 128)])
    R = 384
    e = (R - 1) / R
    coords = ddmi_b200.convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(DEV)
    outs = [m(coords, hdbf=planes, si=0.5) for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])

    mo = cases.build_module('occupancy').to(DEV)
    mo.precision = precision
    pts, hdbf = cases.occupancy_inputs(batch=3, n=150001)
    lo = [mo(pts.to(DEV), _cuda(hdbf)).logits for _ in range(3)]
    assert torch.equal(lo[0], lo[1]) and torch.equal(lo[0], lo[2])

    mv = cases.build_module('video').to(DEV)
    mv.precision = precision
    T, H, W = 5, 96, 160
    xy = [torch.randn(2, 64, H // s, W // s, generator=g) for s in (4, 2, 1)]
    yt = [torch.randn(2, 64, T, H // s, generator=g) for s in (4, 2, 1)]
    xt = [torch.randn(2, 64, T, W // s, generator=g) for s in (4, 2, 1)]
    cv = ddmi_b200.convert_to_coord_format_3d(1, H, W, T, hstart=-.9, hend=.9, wstart=-.9, wend=.9, tstart=-.8, tend=.8)
    vo = [mv(_cuda(cv), _cuda((xy, yt, xt))) for _ in range(3)]
    assert torch.equal(vo[0], vo[1]) and torch.equal(vo[0], vo[2])

    mn = cases.build_module('nerf').to(DEV)
    fea = _cuda({k: torch.randn(2, 32, 64, 64, generator=g) for k in ('xy', 'yz', 'xz')})
    rays = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'nerf_render.pt'))['rays'].to(DEV).repeat(3, 1)
    ro = [nh.render_rays_fused(rays, fea, mn, 128, True, precision=precision) for _ in range(3)]
    assert torch.equal(ro[0], ro[1]) and torch.equal(ro[0], ro[2])


# ---------------------------------------------------------------- occupancy post-step
def test_marching_cubes_golden(golden_dir):
    """ddmi_mcubes_* against the reference's libmcubes outputs (tests/golden/mcubes.pt): bit-exact float64 vertices and int64
    triangles in the reference's order -- open volumes (no padding), a degenerate one-layer volume, and extract_mesh."""
    from ddmi_b200 import generation as gen
    g = torch.load(os.path.join(golden_dir, 'mcubes.pt'))
    for i, name in enumerate(('a', 'b', 'c')):
        c = g[name]
        vol = cases.mcubes_volume(c['shape'], seed=i)
        v, t = gen.marching_cubes(vol.to(DEV), c['iso'])
        assert v.dtype == torch.float64 and t.dtype == torch.int64
        assert torch.equal(v.cpu(), c['vertices'].reshape(-1, 3)) and torch.equal(t.cpu(), c['triangles'].reshape(-1, 3))
    c = g['mesh']
    vol = cases.mcubes_volume(c['shape'], seed=9, noise=0.5, scale=1.5)
    v, t = gen.extract_mesh(vol.to(DEV), 0.2, 0.1)
    assert torch.equal(v.cpu(), c['vertices']) and torch.equal(t.cpu(), c['triangles'])
    with pytest.raises(RuntimeError):
        gen.marching_cubes(vol, 0.0)                         # CPU tensor
    with pytest.raises(RuntimeError):
        gen.marching_cubes(vol.to(DEV)[0], 0.0)              # not 3-D


def test_marching_cubes_full_grid_vs_reference_and_oracle():
    """128^3 (the reference's generation resolution): against the reference's own code (oracle/_ref) when it travelled with
    the snapshot, and always through size-independent properties: closed surface (every edge in exactly two triangles),
    every vertex on a grid edge between a sample above and one below the threshold, index range, empty volume."""
    import numpy as np
    from ddmi_b200 import generation as gen
    from oracle import mcubes_oracle as mo
    vol = cases.mcubes_volume((128, 128, 128), seed=5, noise=0.3)
    v, t = gen.extract_mesh(vol.to(DEV), 0.2, 0.1)
    assert v.shape[0] > 50000 and int(t.min()) == 0 and int(t.max()) == v.shape[0] - 1
    ref = mo.ref_marching_cubes(np.pad(vol.numpy().astype(np.float64), 1, 'constant', constant_values=-1e6),
                                np.log(0.2) - np.log(0.8))
    if ref is not None:
        rv = 1.1 * ((ref[0] - 0.5 - 1) / 127.0 - 0.5)
        assert np.array_equal(rv, v.cpu().numpy()) and np.array_equal(ref[1], t.cpu().numpy())
    tt = t.cpu().numpy()
    e = np.sort(np.concatenate([tt[:, [0, 1]], tt[:, [1, 2]], tt[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    # raw grid-unit vertices: exactly one coordinate is off the half-integer lattice (or the vertex sits on a sample)
    rawv, _ = gen.marching_cubes(torch.nn.functional.pad(vol, (1,) * 6, value=-1e6).to(DEV), np.log(0.2) - np.log(0.8))
    frac = (rawv - 0.5) - torch.floor(rawv - 0.5)
    assert int(((frac != 0).sum(dim=1) <= 1).all())
    ev, et = gen.extract_mesh(torch.full((16, 16, 16), -5.0, device=DEV), 0.2, 0.1)
    assert ev.shape == (0, 3) and et.shape == (0, 3)


def test_generate_mesh_pipeline_matches_oracle_on_decoded_logits():
    """eval_points -> value grid -> extract_mesh, all on the device, against the oracle's mesh of the SAME decoded logits
    (the decode itself is covered by the occupancy parity tests), and eval_points' chunking against one big query."""
    import numpy as np
    from ddmi_b200 import generation as gen
    from oracle import mcubes_oracle as mo
    m = cases.build_module('occupancy').to(DEV)
    _, hdbf = cases.occupancy_inputs(n=10)
    c = tuple([t[:1].to(DEV) * 3.0 for t in axis] for axis in hdbf)
    v, t, grid = gen.generate_mesh(c, m, resolution0=24, points_batch_size=5000)
    assert grid.shape == (24, 24, 24)
    pts = (1.1 * ddmi_b200.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (24,) * 3)).to(DEV)
    assert torch.equal(grid.reshape(-1), m(pts[None], c).logits[0])
    ov, ot = mo.extract_mesh(grid.cpu().numpy(), 0.2, 0.1)
    assert np.array_equal(ov.reshape(-1, 3), v.cpu().numpy()) and np.array_equal(ot.reshape(-1, 3), t.cpu().numpy())


def test_nerf_render_poses_equals_the_loop_of_renders():
    """SURVEY 8f row 4, batching across views: one launch over the rays of all poses == the reference's loop of render() calls."""
    m = cases.build_module('nerf').to(DEV)
    res, K, fea, _ = cases.nerf_inputs(res=12)
    embed_fn, _ = nh.get_embedder(10, 0)
    embeddirs_fn, _ = nh.get_embedder(4, 0)
    cfg = {'model': {'TN': {'netchunk': 40000, 'peturb': 0, 'N_importance': 0, 'N_samples': 128, 'use_viewdirs': True,
                            'white_bkgd': True, 'raw_noise_std': 0}}}
    kw = nh.get_render_kwargs(cfg, m, embed_fn, embeddirs_fn)
    near, far = kw.pop('near'), kw.pop('far')
    kw.pop('use_viewdirs', None)
    kw.pop('ndc', None)
    poses = [nh.pose_spherical(th, -20, 5) for th in (0.0, 95.0, 250.0)]
    feac = _cuda(fea)
    batched = nh.render_poses(res, res, K, feac, poses, DEV, near=near, far=far, **kw)
    assert batched.shape == (3, res * res, 3)
    for v, pose in enumerate(poses):
        one = nh.render(res, res, K, feac, None, 0, DEV, chunk=4096, c2w=pose[:3, :4], near=near, far=far, use_viewdirs=True,
                        verbose=True, retraw=True, hw_idx=None, **kw)
        assert torch.equal(one, batched[v])


# ---------------------------------------------------------------- plane-producer tail
@pytest.mark.parametrize("channels_last", [False, True])
def test_plane_tail_golden(golden_dir, channels_last):
    """ddmi_plane_head / ddmi_plane_tail against the layers' inputs and outputs captured inside the reference Decoder's own
    forward (tests/golden/plane_tail.pt), both output layouts; fp32 kernels: 2e-5 of |out|max."""
    g = torch.load(os.path.join(golden_dir, 'plane_tail.pt'))
    for tag in ('plain', 'tanh'):
        c = g[tag]
        sd = c['sd']
        ins = [sd[f'up.{i}.hdbf.0.weight'].shape[1] if f'up.{i}.hdbf.0.weight' in sd else None for i in range(3)]
        t = ddmi_b200.PlaneTail(sd['conv_out.weight'].shape[1], 64, ins, tanh_out=c['tanh_out']).to(DEV)
        t.load_state_dict(sd, strict=True)
        out = t.tail(c['tail_in'].to(DEV), channels_last=channels_last)
        assert out.shape == c['tail_out'].shape
        assert out.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        assert float((out.cpu() - c['tail_out']).abs().max()) < 2e-5 * max(1.0, float(c['tail_out'].abs().max()))
        for k, (hin, hout) in c['heads'].items():
            o = t.head(int(k[4:]), hin.to(DEV), channels_last=channels_last)
            assert float((o.cpu() - hout).abs().max()) < 2e-5 * max(1.0, float(hout.abs().max()))
    with pytest.raises(RuntimeError):
        t.tail(c['tail_in'])                                 # CPU tensor
    with pytest.raises(RuntimeError):
        t.tail(torch.zeros(1, 7, 8, 8, device=DEV))          # wrong channel count


def test_plane_tail_odd_shapes_and_srn_cars_width():
    """Ragged tiles (H, W not multiples of the 8 x 32 output tile), C_in not a multiple of the 8-channel chunk, 32 output
    channels (srn-cars planes), against the oracle."""
    from oracle import plane_tail_oracle as po
    g = torch.Generator().manual_seed(11)
    t = ddmi_b200.PlaneTail(96, 32, (None, 44), tanh_out=False).to(DEV)
    for p in t.parameters():
        p.data = torch.randn(p.shape, generator=g).to(DEV) * 0.2
    sd = {k: v.cpu() for k, v in t.state_dict().items()}
    h = torch.randn(3, 96, 37, 45, generator=g) * 2 + 0.5
    out = t.tail(h.to(DEV))
    ref = po.tail(sd, h, 32, False)
    assert float((out.cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    h1 = torch.randn(2, 44, 5, 70, generator=g)
    o1 = t.head(1, h1.to(DEV), channels_last=True)
    assert float((o1.cpu() - po.head(sd, 1, h1)).abs().max()) < 2e-5 * 10


def test_channels_last_planes_from_the_tail_feed_the_occupancy_decoder_without_a_copy():
    """A plane emitted channels-last by PlaneTail (torch.channels_last strides) goes into MLP3D as is: same logits as the NCHW
    planes, and the library's transposition kernel does not run (the cache holds views of the caller's storage)."""
    m = cases.build_module('occupancy').to(DEV)
    pts, hdbf = cases.occupancy_inputs(batch=2, n=3000)
    nchw = _cuda(hdbf)
    cl = tuple([t.contiguous(memory_format=torch.channels_last) for t in axis] for axis in nchw)
    a = m(pts.to(DEV), nchw).logits
    b = m(pts.to(DEV), cl).logits
    assert torch.equal(a, b)
    held = m._nhwc_cache.value[0]
    assert all(x.data_ptr() == t.data_ptr() for x, t in zip(held, [t for axis in cl for t in axis]))
