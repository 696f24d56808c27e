"""Drop-in INR decoders of the D2C-VAE (reference: models/d2c_vae/mlp.py).

Same constructors, parameter names / shapes (state-dict compatible) and
``forward`` signatures as the reference classes; ``forward`` is host-side weight
folding + ONE fused CUDA launch through the C ABI (include/ddmi_b200.h).
Inference only: if autograd would need a gradient through the decode, we raise
instead of silently detaching (the reference back-props through this path during
training; that is out of scope).
"""
import contextlib
import os
import warnings

import torch
from torch import distributions as dist
from torch import nn

from . import _lib, packing
from .blocks import ResnetBlockFC, SinusoidalPosEmb, StyledResBlock, ToRGB

_PREC = {'fp32': _lib.PREC_FP32, 'bf16x3': _lib.PREC_BF16X3, 'f16f8': _lib.PREC_F16F8}


def _resolve_precision(name, supported, default):
    name = name or os.environ.get('DDMI_B200_PRECISION') or default
    if name not in _PREC:
        raise ValueError(f"precision must be one of {sorted(_PREC)} (got {name!r})")
    if name not in supported:
        raise NotImplementedError(f"precision {name!r} has no kernel for this decoder (supported: {supported})")
    return _PREC[name]


def _store_mode(store):
    if store not in _lib.STORE_MODES:
        raise ValueError(f"store must be one of None, 'f32', 'clamp', 'u8' (got {store!r})")
    return _lib.STORE_MODES[store]


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _as_plane(t, name, keep_channels_last=False):
    """keep_channels_last: the caller gathers from channels-last planes anyway (occupancy / NeRF tensor-core kernels), so a
    plane that already IS channels-last (torch.channels_last strides, as ddmi_b200.plane_tail emits them) is passed through."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (ddmi_b200 has no CPU path)")
    if t.dim() != 4:
        raise RuntimeError(f"{name} must be (B,C,H,W), got {tuple(t.shape)}")
    t = t.detach()
    if keep_channels_last and _lib.is_channels_last(t):
        return t
    return t.to(torch.float32).contiguous()


def _check_plane_set(planes, channels, names):
    """Every plane of a decode shares ONE batch / channel count / device: the C ABI takes a single (batch, channels) pair for
    all of them, so a smaller plane would be indexed past its end.  (The reference fails in grid_sample / cat on the same
    inputs.)"""
    b, dev = planes[0].shape[0], planes[0].device
    for t, name in zip(planes, names):
        if t.shape[0] != b:
            raise RuntimeError(f"{name} has batch {t.shape[0]}, expected {b} (all planes of a decode share the batch)")
        if t.shape[1] != channels:
            raise RuntimeError(f"{name} has {t.shape[1]} channels, expected latent_dim = {channels}")
        if t.device != dev:
            raise RuntimeError(f"{name} is on {t.device}, expected {dev} (all planes of a decode live on one GPU)")
        if t.shape[2] < 1 or t.shape[3] < 1:
            raise RuntimeError(f"{name} has an empty extent {tuple(t.shape)}")


class _ChannelsLastCache:
    """Channels-last copies of the planes of the most recent decode, keyed by the identity of the source tensors
    (data_ptr, version counter, shape).  The reference's callers query the SAME planes in a loop -- 100k-point chunks of one
    latent in Generator3D.eval_points (convocc/src/conv_onet/generation.py:130-144), one render per pose in
    tools/ldm/nerf.py:270 -- so the transposition runs once per latent instead of once per call.  As with the packed
    weights, writes through `.data` that keep the version counter are not seen: re-create the tensor or call clear()."""

    def __init__(self):
        self.key, self.value, self.sources = None, None, None

    def get(self, planes, sources, stream):
        vers = [packing.tensor_version(t) for t in sources]
        if any(v is None for v in vers):        # inference-mode tensors carry no version counter: never cached
            return _lib.planes_channels_last(planes, stream)
        key = tuple((t.data_ptr(), v, tuple(t.shape), str(t.device)) for t, v in zip(sources, vers))
        if key != self.key:
            self.value = _lib.planes_channels_last(planes, stream)
            self.key = key
            self.sources = list(sources)        # keep the sources alive: a freed tensor's address can be re-used
        return self.value

    def clear(self):
        self.key, self.value, self.sources = None, None, None


class _FusedDecoder(nn.Module):
    """Common plumbing: grad guard + packed-weight cache."""

    _supported = ('fp32',)
    _default_precision = 'fp32'

    def _init_fused(self, precision=None):
        self.precision = precision
        self._pack_cache = {}
        self._nhwc_cache = _ChannelsLastCache()

    def invalidate_packed(self):
        """Drop every cached packed-weight blob and channels-last plane copy (they are rebuilt on the next call)."""
        self._pack_cache = {}
        self._nhwc_cache.clear()

    def _guard_grad(self, *tensors):
        if torch.is_grad_enabled() and (
                any(p.requires_grad for p in self.parameters())
                or any(torch.is_tensor(t) and t.requires_grad for t in tensors)):
            raise RuntimeError(
                "ddmi_b200 decoders are forward/inference only; call them under "
                "torch.no_grad() / torch.inference_mode() (no autograd graph is built)")

    @contextlib.contextmanager
    def weights_unchanged(self):
        """Inside this block the parameters are not modified (the caller promises): the packed-weight cache is validated
        on the first decode only, later decodes skip the per-call parameter digest (two norm kernels + one host sync,
        ~0.25 ms -- as much as a 50k-point decode).  Used by the library's own loops (generation.eval_points, sharding)."""
        self._frozen = getattr(self, '_frozen', 0) + 1
        self._frozen_checked = False
        try:
            yield self
        finally:
            self._frozen -= 1

    def _packed(self, key, builder):
        if getattr(self, '_frozen', 0) and self._frozen_checked and key in self._pack_cache.get('entries', {}):
            return self._pack_cache['entries'][key]
        fp = packing.param_fingerprint(self)
        self._frozen_checked = True
        if not packing.same_fingerprint(self._pack_cache.get('fingerprint'), fp):   # parameters changed: drop every packed blob
            self._pack_cache = {'fingerprint': fp, 'entries': {}}
        entries = self._pack_cache['entries']
        if key not in entries:
            if len(entries) >= 16:                         # bounded (one entry per (precision, si))
                entries.pop(next(iter(entries)))
            entries[key] = builder()
        return entries[key]

    def _packed_auto(self, prec, key_of, build_of):
        """Packed weights for `prec`.  When f16f8 is only the DEFAULT (neither the module nor DDMI_B200_PRECISION asked
        for it) and a weight does not fit its fp16 range, fall back to bf16x3 with a warning; an explicit request raises."""
        try:
            return prec, self._packed(key_of(prec), lambda: build_of(prec))
        except packing.F16F8RangeError:
            if self.precision or os.environ.get('DDMI_B200_PRECISION'):
                raise
            warnings.warn("ddmi_b200: a weight exceeds the f16f8 operand range (|w| >= 16); using precision='bf16x3'")
            prec = _lib.PREC_BF16X3
            return prec, self._packed(key_of(prec), lambda: build_of(prec))

    def _check_device(self, t):
        dev = next(self.parameters()).device
        if not t.is_cuda or dev != t.device:
            raise RuntimeError(
                f"decoder parameters are on {dev} but inputs on {t.device}; both must be the same CUDA device")


class MLP(_FusedDecoder):
    """Image decoder.  Reference: models/d2c_vae/mlp.py:12-66."""

    _supported = ('fp32', 'bf16x3', 'f16f8')
    _default_precision = 'f16f8'

    def __init__(self, *, in_ch=2, latent_dim=64, out_ch=3, ch=256, precision=None):
        super().__init__()
        if (in_ch, latent_dim, out_ch, ch) != (2, 64, 3, 256):
            raise NotImplementedError(
                "the fused image kernel is specialised for in_ch=2, latent_dim=64, out_ch=3, ch=256 "
                "(configs/d2c-vae/afhq.yaml:44-48)")
        self.latent_dim = latent_dim
        dim = ch // 4
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, ch), nn.GELU(), nn.Linear(ch, ch))
        self.net_res1 = StyledResBlock(in_ch + latent_dim, ch, 1, ch, demodulate=True)
        self.net_res2 = StyledResBlock(ch + in_ch + latent_dim, ch, 1, ch, demodulate=True)
        self.net_res3 = StyledResBlock(ch + in_ch + latent_dim, ch, 1, ch, demodulate=True)
        self.net_res4 = StyledResBlock(ch, ch, 1, ch, demodulate=True)
        self.torgb = ToRGB(ch, out_ch, ch)
        self._init_fused(precision)

    def _noise_source(self, noise, b, h, w, dev):
        """-> (mode, ctypes pointer array or None, seed, tensors to keep alive) for a decode whose NoiseInjection weights are
        non-zero.  noise = None: a fresh seed from torch's default CPU generator per call (like the reference, which draws new
        noise every forward; reproducible under torch.manual_seed) for the library's Philox stream; an int: that seed; 12
        tensors (b,1,h,w) or one (12,b,1,h,w): explicit noise, order net_res1.conv1, conv2, conv3, net_res2.conv1, ..."""
        import ctypes
        if noise is None:
            return _lib.NOISE_PHILOX, None, int(torch.randint(0, 2 ** 62, (1,)).item()), None
        if isinstance(noise, int):
            return _lib.NOISE_PHILOX, None, noise & (2 ** 64 - 1), None
        ts = list(noise.unbind(0)) if torch.is_tensor(noise) else list(noise)
        if len(ts) != 12:
            raise RuntimeError(f"noise must hold 12 tensors (one per StyledConv), got {len(ts)}")
        keep = []
        for i, t in enumerate(ts):
            if t.numel() != b * h * w or t.shape[0] != b:
                raise RuntimeError(f"noise[{i}] must be ({b},1,{h},{w}), got {tuple(t.shape)}")
            keep.append(t.detach().to(device=dev, dtype=torch.float32).contiguous())
        arr = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in keep])
        return _lib.NOISE_TENSORS, arr, 0, keep

    def forward(self, coords, hdbf, si=1, store=None, noise=None):
        """coords (1,2,h,w) in [-1,1]; hdbf = 3 planes (b,64,S,S) coarse->fine;
        returns (b,3,h,w).  mlp.py:34-66.
        noise (extension; only read when some NoiseInjection.weight is non-zero, i.e. for trained checkpoints): None, an int
        seed, or the 12 explicit (b,1,h,w) tensors -- see _noise_source.
        store (extension, the epilogue the reference's callers apply -- fused into the kernel's output stage):
        None / 'f32' -> the reference's value; 'clamp' -> .clamp(-1, 1) (evals/eval.py:162,226);
        'u8' -> ((x.clamp(-1,1) + 1) * 127.5).type(torch.uint8) channels-last, (b,h,w,3) (evals/eval.py:289,336-337)."""
        assert hdbf is not None and len(hdbf) == 3
        self._guard_grad(*hdbf)
        planes = [_as_plane(t, f'hdbf[{i}]') for i, t in enumerate(hdbf)]
        self._check_device(planes[0])
        _check_plane_set(planes, self.latent_dim, [f'hdbf[{i}]' for i in range(3)])
        if coords.dim() != 4 or coords.shape[0] != 1 or coords.shape[1] != 2:
            raise RuntimeError(f"coords must be (1,2,h,w), got {tuple(coords.shape)}")
        _, _, h, w = coords.shape
        b = planes[0].shape[0]
        c = coords.detach().to(device=planes[0].device, dtype=torch.float32).contiguous()
        prec = _resolve_precision(self.precision, self._supported, self._default_precision)
        si = float(si)
        env_pair = os.environ.get('DDMI_B200_CTA_PAIR', '1') != '0'     # tcgen05 kernels: CTA pairs (cta_group::2) by default
        pair_of = lambda pr: env_pair or pr == _lib.PREC_F16F8          # the f16f8 kernels exist for pairs only
        prec, packed = self._packed_auto(prec, lambda pr: ('image', pr, si, pair_of(pr)),
                                         lambda pr: packing.pack_image(self, si, pr, pair_of(pr)))
        mode = _store_mode(store)
        if mode == _lib.STORE_U8_CHANNELS_LAST:
            out = torch.empty((b, h, w, 3), device=c.device, dtype=torch.uint8)
        else:
            out = torch.empty((b, 3, h, w), device=c.device, dtype=torch.float32)
        n = h * w
        cx, cy = c[0, 0], c[0, 1]
        nmode, narr, seed, keep = (_lib.NOISE_NONE, None, 0, None)
        if packed.noise_active:
            nmode, narr, seed, keep = self._noise_source(noise, b, h, w, c.device)
        with torch.cuda.device(c.device):
            wst = _lib.weights_struct(packed)
            _lib.check(_lib.lib().ddmi_decode_image_noise(
                _lib.planes_array(planes), b, planes[0].shape[1], cx.data_ptr(), cy.data_ptr(), n,
                wst, mode, nmode, narr, seed, out.data_ptr(), _stream_ptr(c.device)))
        del keep
        return out


class MLP3D(_FusedDecoder):
    """Occupancy decoder.  Reference: models/d2c_vae/mlp.py:69-111."""

    _supported = ('fp32', 'bf16x3', 'f16f8')
    _default_precision = 'f16f8'

    def __init__(self, *, in_ch, latent_dim, out_ch, ch=256, precision=None):
        super().__init__()
        if (in_ch, latent_dim, out_ch, ch) != (3, 64, 1, 256):
            raise NotImplementedError(
                "the fused occupancy kernel is specialised for in_ch=3, latent_dim=64, out_ch=1, ch=256 "
                "(configs/d2c-vae/shapenet.yaml:44-48)")
        self.latent_dim = latent_dim
        self.net_p = nn.Linear(in_ch, ch)
        self.net_res1 = ResnetBlockFC(latent_dim, ch)
        self.net_res2 = ResnetBlockFC(ch + latent_dim, ch)
        self.net_res3 = ResnetBlockFC(ch + latent_dim, ch)
        self.net_res4 = ResnetBlockFC(ch, ch)
        self.net_out = nn.Linear(ch, out_ch)
        self._init_fused(precision)

    def decode_logits(self, coords, hdbf):
        assert len(hdbf) == 3
        for axis in hdbf:
            assert len(axis) == 3
        self._guard_grad(coords, *[t for axis in hdbf for t in axis])
        sources = [hdbf[a][s] for a in range(3) for s in range(3)]
        names = [f'hdbf[{a}][{s}]' for a in range(3) for s in range(3)]
        planes = [_as_plane(t, nm, keep_channels_last=True) for t, nm in zip(sources, names)]
        self._check_device(planes[0])
        _check_plane_set(planes, self.latent_dim, names)
        b = planes[0].shape[0]
        if coords.dim() != 3 or coords.shape[-1] != 3 or coords.shape[0] not in (1, b):
            raise RuntimeError(f"coords must be ({b},N,3), got {tuple(coords.shape)}")
        pts = coords.detach().to(device=planes[0].device, dtype=torch.float32)
        if pts.shape[0] == 1 and b > 1:
            pts = pts.expand(b, -1, -1)
        n = pts.shape[1]
        if n == 0:                                   # empty query set: nothing to launch (the reference returns (B, 0) too)
            return torch.empty((b, 0), device=planes[0].device, dtype=torch.float32)
        if pts.stride(0) == 0 or b == 1:
            base, bstride = pts[0].contiguous(), 0
        else:
            base = pts.contiguous()
            bstride = n * 3
        prec = _resolve_precision(self.precision, self._supported, self._default_precision)
        env_pair = os.environ.get('DDMI_B200_CTA_PAIR', '1') != '0'
        pair_of = lambda pr: env_pair or pr == _lib.PREC_F16F8
        prec, packed = self._packed_auto(prec, lambda pr: ('occ', pr, pair_of(pr)),
                                         lambda pr: packing.pack_occupancy(self, pr, pair_of(pr)))
        logits = torch.empty((b, n), device=base.device, dtype=torch.float32)
        with torch.cuda.device(base.device):
            st = _stream_ptr(base.device)
            if prec != _lib.PREC_FP32:       # scattered queries: channels-last planes, float4 gathers
                keep, arr = self._nhwc_cache.get(planes, sources, st)
                layout = 1
            else:
                planes = [p.contiguous() for p in planes]
                keep, arr, layout = planes, _lib.planes_array(planes), 0
            _lib.check(_lib.lib().ddmi_decode_occupancy(
                arr, b, planes[0].shape[1], layout, base.data_ptr(), n, bstride, 0.1,
                _lib.weights_struct(packed), logits.data_ptr(), st))
            del keep
        return logits

    def decode_logits_lattice(self, axes, hdbf):
        """Logits on the query lattice axes = (xs, ys, zs) (1-D tensors): (B, nx, ny, nz), bit-identical to
        decode_logits(torch.cartesian_prod(xs, ys, zs)[None], hdbf).reshape(B, nx, ny, nz) -- the reference's dense-grid query
        `box_size * make_3d_grid(...)` (convocc/src/conv_onet/generation.py:90-97) without the point list: on a lattice each
        plane's sample depends on two of the three indices only, so the library samples nx ny + ny nz + nx nz feature vectors
        per scale once and every point reads three of them (include/ddmi_b200.h, ddmi_decode_occupancy_lattice)."""
        assert len(hdbf) == 3 and all(len(axis) == 3 for axis in hdbf) and len(axes) == 3
        self._guard_grad(*axes, *[t for axis in hdbf for t in axis])
        sources = [hdbf[a][s] for a in range(3) for s in range(3)]
        names = [f'hdbf[{a}][{s}]' for a in range(3) for s in range(3)]
        planes = [_as_plane(t, nm, keep_channels_last=True) for t, nm in zip(sources, names)]
        self._check_device(planes[0])
        _check_plane_set(planes, self.latent_dim, names)
        dev = planes[0].device
        b = planes[0].shape[0]
        ax = [a.detach().to(device=dev, dtype=torch.float32).reshape(-1) for a in axes]
        nx, ny, nz = (int(a.numel()) for a in ax)
        prec = _resolve_precision(self.precision, self._supported, self._default_precision)
        if prec != _lib.PREC_FP32:
            prec, packed = self._packed_auto(prec, lambda pr: ('occ', pr, True), lambda pr: packing.pack_occupancy(self, pr, True))
        ws_bytes = int(_lib.lib().ddmi_occupancy_lattice_workspace_bytes(b, nx, ny, nz)) if min(nx, ny, nz) > 0 else 0
        if prec == _lib.PREC_FP32 or ws_bytes == 0 or ws_bytes > int(os.environ.get('DDMI_B200_LATTICE_TABLE_MAX', 8 << 30)):
            pts = torch.cartesian_prod(*ax).reshape(1, nx * ny * nz, 3)          # exact-arithmetic path / oversized lattices
            return self.decode_logits(pts, hdbf).reshape(b, nx, ny, nz)
        logits = torch.empty((b, nx, ny, nz), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            keep, arr = self._nhwc_cache.get(planes, sources, st)
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            _lib.check(_lib.lib().ddmi_decode_occupancy_lattice(
                arr, b, planes[0].shape[1], 1, torch.cat(ax).data_ptr(), nx, ny, nz, 0.1, _lib.weights_struct(packed),
                logits.data_ptr(), ws.data_ptr(), ws_bytes, st))
            del keep, ws
        return logits

    def forward(self, coords, hdbf):
        """coords (B,N,3); hdbf = (xy, yz, xz), each a 3-list of (B,64,R,R);
        returns Bernoulli(logits (B,N)).  mlp.py:82-111."""
        return dist.Bernoulli(logits=self.decode_logits(coords, hdbf), validate_args=False)   # (validation = 0.15 ms of host time)


class MLPVideo(_FusedDecoder):
    """Video decoder.  Reference: models/d2c_vae/mlp.py:114-157."""

    _supported = ('fp32', 'bf16x3', 'f16f8')
    _default_precision = 'f16f8'

    def __init__(self, *, in_ch, latent_dim, out_ch, ch=256, precision=None, **ignore_kwargs):
        super().__init__()
        if (latent_dim, out_ch, ch) != (64, 3, 256):
            raise NotImplementedError(
                "the fused video kernel is specialised for latent_dim=64, out_ch=3, ch=256 "
                "(configs/d2c-vae/skytimelapse.yaml:49-53)")
        self.latent_dim = latent_dim
        self.out_ch = out_ch
        self.net_res1 = ResnetBlockFC(latent_dim * 3, ch)
        self.net_res2 = ResnetBlockFC(ch + latent_dim * 3, ch)
        self.net_res3 = ResnetBlockFC(ch + latent_dim * 3, ch)
        self.net_res4 = ResnetBlockFC(ch)
        self.net_out = nn.Linear(ch, out_ch)
        self._init_fused(precision)

    def forward(self, coords, hdbf, store=None):
        """coords = {'xy':(1,2,H,W),'xt':(1,2,T,W),'yt':(1,2,T,H)}; hdbf =
        (xy, yt, xt) 3-lists; returns (b,3,t,h,w) with t,h,w taken from the
        finest planes exactly as the reference does (mlp.py:135-136,155-156).
        store (extension, see MLP.forward): 'clamp' -> clamp(-1,1); 'u8' -> uint8 frames (b,t,h,w,3), the
        `rearrange((fake.clamp(-1,1) + 1) * 127.5, 'b c t h w -> b t h w c').type(torch.uint8)` of evals/eval.py:336-337."""
        assert len(hdbf) == 3
        xy_hdbf, yt_hdbf, xt_hdbf = hdbf
        assert len(xy_hdbf) == 3 and len(yt_hdbf) == 3 and len(xt_hdbf) == 3
        self._guard_grad(*xy_hdbf, *yt_hdbf, *xt_hdbf)
        names = [f'hdbf[{a}][{s}]' for a in range(3) for s in range(3)]
        planes = [_as_plane(hdbf[a][s], names[3 * a + s]) for a in range(3) for s in range(3)]
        self._check_device(planes[0])
        _check_plane_set(planes, self.latent_dim, names)
        dev = planes[0].device
        b, _, h, w = xy_hdbf[-1].shape
        t = yt_hdbf[-1].shape[2]
        cxy = coords['xy'].detach().to(device=dev, dtype=torch.float32).contiguous()
        cyt = coords['yt'].detach().to(device=dev, dtype=torch.float32).contiguous()
        cxt = coords['xt'].detach().to(device=dev, dtype=torch.float32).contiguous()
        if cxy.shape[0] != 1 or cyt.shape[0] != 1 or cxt.shape[0] != 1:
            raise RuntimeError("coords grids must have batch 1 (they are shared by all items)")
        _, _, H, W = cxy.shape
        T = cyt.shape[2]
        if tuple(cyt.shape[1:]) != (2, T, H) or tuple(cxt.shape[1:]) != (2, T, W):
            raise RuntimeError("inconsistent coords grids: expected xy (1,2,H,W), yt (1,2,T,H), xt (1,2,T,W)")
        prec = _resolve_precision(self.precision, self._supported, self._default_precision)
        env_pair = os.environ.get('DDMI_B200_CTA_PAIR', '1') != '0'
        pair_of = lambda pr: env_pair or pr == _lib.PREC_F16F8
        prec, packed = self._packed_auto(prec, lambda pr: ('video', pr, pair_of(pr)),
                                         lambda pr: packing.pack_video(self, pr, pair_of(pr)))
        mode = _store_mode(store)
        u8 = mode == _lib.STORE_U8_CHANNELS_LAST
        out = torch.empty((b, T * H * W, self.out_ch) if u8 else (b, self.out_ch, T * H * W), device=dev,
                          dtype=torch.uint8 if u8 else torch.float32)
        with torch.cuda.device(dev):
            # feature tables (include/ddmi_b200.h, ddmi_decode_video_ws): scratch for this call, sized by the library; very large
            # volumes / batches decode with direct gathers instead (DDMI_B200_VIDEO_TABLE_MAX bytes, default 8 GB; 0 disables)
            ws, ws_bytes = None, int(_lib.lib().ddmi_video_workspace_bytes(b, T, H, W, prec))
            if 0 < ws_bytes <= int(os.environ.get('DDMI_B200_VIDEO_TABLE_MAX', 8 << 30)):
                ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            _lib.check(_lib.lib().ddmi_decode_video_ws(
                _lib.planes_array(planes), b, planes[0].shape[1], cxy.data_ptr(), cyt.data_ptr(),
                cxt.data_ptr(), T, H, W, _lib.weights_struct(packed), mode, out.data_ptr(),
                ws.data_ptr() if ws is not None else None, ws_bytes if ws is not None else 0, _stream_ptr(dev)))
            del ws       # stream-ordered caching allocator: the block is not handed out again before this stream's work is done
        return out.reshape(b, t, h, w, self.out_ch) if u8 else out.reshape(b, self.out_ch, t, h, w)


class MLPNeRF(_FusedDecoder):
    """NeRF density/colour MLP.  Reference: models/d2c_vae/mlp.py:199-281."""

    def __init__(self, D=8, W=256, in_channels_xyz=96, in_channels_dir=27, skips=[2, 4, 6], precision=None):
        super().__init__()
        self.D, self.W = D, W
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.skips = skips
        # nn.LeakyReLU(True): the positional argument is negative_slope, so the
        # slope is 1.0 and the activation is the identity.  Kept
        # as a module so state dicts / repr line up; the kernel takes the slope.
        for i in range(D):
            if i == 0:
                layer = nn.Linear(in_channels_xyz, W)
            elif i in skips:
                layer = nn.Linear(W + in_channels_xyz, W)
            else:
                layer = nn.Linear(W, W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(layer, nn.LeakyReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2), nn.LeakyReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        self._init_fused(precision)

    @property
    def negative_slope(self):
        return float(self.dir_encoding[1].negative_slope)

    def packed_weights(self, precision=None):
        """precision: 'fp32' (MLPNeRF.forward on given rows) or 'bf16x3' / 'f16f8' (fused tcgen05 ray render)."""
        prec = _resolve_precision(precision or self.precision, ('fp32', 'bf16x3', 'f16f8'), self._default_precision)
        return self._packed(('nerf', prec), lambda: packing.pack_nerf(self, prec))

    def forward(self, x, sigma_only=False):
        """x (B,186) [or (B,159) if sigma_only] -> (B,4) [rgb, sigma] or (B,1).  mlp.py:241-281."""
        self._guard_grad(x)
        self._check_device(x)
        need = self.in_channels_xyz if sigma_only else self.in_channels_xyz + self.in_channels_dir
        if x.dim() != 2 or x.shape[1] != need:
            raise RuntimeError(f"x must be (B,{need}), got {tuple(x.shape)}")
        xx = x.detach().to(torch.float32).contiguous()
        n = xx.shape[0]
        out = torch.empty((n, 1 if sigma_only else 4), device=xx.device, dtype=torch.float32)
        if n == 0:
            return out
        packed = self.packed_weights('fp32')     # rows in, rows out: the fp32 kernel (the tcgen05 kernel is the fused render)
        with torch.cuda.device(xx.device):
            _lib.check(_lib.lib().ddmi_nerf_mlp(
                xx.data_ptr(), n, xx.shape[1], 1 if sigma_only else 0, self.negative_slope,
                _lib.weights_struct(packed), out.data_ptr(), _stream_ptr(xx.device)))
        return out
