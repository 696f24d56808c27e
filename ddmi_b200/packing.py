"""Host-side weight folding and packing for the fused decode kernels.

Folding = the per-call weight algebra the reference performs inside its layers
(style modulation / demodulation of the image MLP, equalised-lr scales, the two
constant scale-injection input channels), done once per call on tiny matrices in
float64 and rounded to fp32.  Packing = laying the folded matrices out in the
order and format a kernel family streams them:

* ``PREC_FP32``   : per GEMM segment a row-major ``[K_pad][N]`` fp32 block
  (``W[:, seg].T``), K padded to a multiple of 16 with zero rows.
* ``PREC_BF16X3`` : per 16-wide K step a ``[hi | lo]`` pair of bf16 blocks in the
  tcgen05 shared-memory "core matrix" order, in consumption order.

Both formats share one ``vec`` blob of fp32 vectors (biases, folded constants,
the narrow output heads that run in the epilogue).
"""
import math
from dataclasses import dataclass

import torch

from ._lib import PREC_BF16X3, PREC_FP32


@dataclass
class Packed:
    precision: int
    gemm: torch.Tensor   # flat device tensor (fp32 or int16 bit patterns of bf16)
    vec: torch.Tensor    # flat fp32 device tensor


# ---------------------------------------------------------------------------
# generic helpers
# ---------------------------------------------------------------------------
def _seg_fp32(W, lo, hi, kpad=None):
    """[K_pad][N] block of W[:, lo:hi].T (W is (N, K_total) float64/32)."""
    blk = W[:, lo:hi].t().contiguous()
    k = hi - lo
    kp = kpad if kpad is not None else (k + 15) // 16 * 16
    if kp != k:
        blk = torch.cat([blk, blk.new_zeros(kp - k, blk.shape[1])], dim=0)
    return blk.reshape(-1)


def _split_bf16(x32):
    """fp32 -> (hi, lo) bf16 with hi + lo == x to ~2^-17 relative."""
    hi = x32.to(torch.bfloat16)
    lo = (x32 - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def umma_kstep_blocks(W32, lo, hi, n_pad=None):
    """Pack W[:, lo:hi] (N, K) for the tcgen05 kernels.

    Returns a flat int16 tensor: for every 16-wide K step ``[hi block | lo block]``,
    each block = 2 core-matrix columns x N rows x 8 bf16 (canonical K-major,
    no-swizzle layout: element (n, k) at ((k//8)*N + n)*8 + k%8 within the step).
    K is zero-padded to a multiple of 16, N to ``n_pad``.
    """
    W = W32[:, lo:hi].to(torch.float32)
    n, k = W.shape
    kp = (k + 15) // 16 * 16
    npad = n if n_pad is None else n_pad
    if kp != k or npad != n:
        Wp = W.new_zeros(npad, kp)
        Wp[:n, :k] = W
        W = Wp
    h, l = _split_bf16(W)
    out = []
    for part in (h, l):
        # (N, K) -> (K/16 steps, 2 kgroups, N, 8)
        p = part.reshape(npad, kp // 16, 2, 8).permute(1, 2, 0, 3).contiguous()
        out.append(p.view(torch.int16).reshape(kp // 16, -1))
    return torch.cat(out, dim=1).reshape(-1)  # per step: hi block then lo block


def _params64(module):
    return {k: v.detach().to(torch.float64) for k, v in module.state_dict().items()}


def param_fingerprint(module):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in module.parameters())


# ---------------------------------------------------------------------------
# image MLP (models/d2c_vae/mlp.py:12-66)
# ---------------------------------------------------------------------------
def image_style(p, si, ch=256):
    """style = time_mlp(si): SinusoidalPosEmb (blocks.py:11-23) -> Linear ->
    exact-erf GELU -> Linear.  Identical for every batch item (mlp.py:46-47)."""
    dim = ch // 4
    half = dim // 2
    dev = p['time_mlp.1.weight'].device
    freqs = torch.exp(torch.arange(half, device=dev, dtype=torch.float64) * -(math.log(10000) / (half - 1)))
    e = float(si) * freqs
    emb = torch.cat((e.sin(), e.cos()))
    h = p['time_mlp.1.weight'] @ emb + p['time_mlp.1.bias']
    h = torch.nn.functional.gelu(h)
    return p['time_mlp.3.weight'] @ h + p['time_mlp.3.bias']


def _modulated(p, prefix, style, demodulate):
    """Folded ModulatedConv2d 1x1 weight (blocks.py:242-251)."""
    W = p[prefix + '.weight'][0, :, :, 0, 0]                    # (out, in)
    mod_w, mod_b = p[prefix + '.modulation.weight'], p[prefix + '.modulation.bias']
    m = (mod_w * (1.0 / math.sqrt(mod_w.shape[1]))) @ style + mod_b   # EqualLinear, lr_mul = 1
    Wf = (1.0 / math.sqrt(W.shape[1])) * W * m[None, :]
    if demodulate:
        Wf = Wf * torch.rsqrt((Wf * Wf).sum(dim=1, keepdim=True) + 1e-8)
    return Wf


def fold_image(module, si):
    """Returns dict: blocks[i] = {W1,W2,W3,Ws|None,b1,b2,b3,cs}, Wrgb, brgb (float64).

    * W1 / Ws keep only the real input columns ([h | PE]); the two constant
      scale-injection channels (mlp.py:44,51) become the constants folded into
      b1 / cs.
    * Ws, cs carry the block's 1/sqrt(2) (blocks.py:636); conv3's sqrt(2)
      activation gain cancels against it inside the kernel.
    """
    p = _params64(module)
    for k, v in p.items():
        if k.endswith('noise.weight') and float(v.abs().max()) != 0.0:
            raise NotImplementedError(
                f"{k} != 0: the reference draws fresh N(0,1) noise inside forward "
                "(blocks.py:293-297), so its output is not reproducible; noise injection is unsupported")
    style = image_style(p, si)
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    blocks = []
    for i in range(1, 5):
        pre = f'net_res{i}'
        d = {}
        W1 = _modulated(p, f'{pre}.conv1.conv', style, True)
        kin = W1.shape[1]
        nreal = kin - 2 if i < 4 else kin
        c1 = (W1[:, nreal:].sum(dim=1) * float(si)) if i < 4 else torch.zeros_like(W1[:, 0])
        d['W1'] = W1[:, :nreal]
        d['b1'] = p[f'{pre}.conv1.activate.bias'] + c1
        d['W2'] = _modulated(p, f'{pre}.conv2.conv', style, True)
        d['b2'] = p[f'{pre}.conv2.activate.bias']
        d['W3'] = _modulated(p, f'{pre}.conv3.conv', style, True)
        d['b3'] = p[f'{pre}.conv3.activate.bias']
        if i < 4:
            Ws = p[f'{pre}.skip.0.weight'][:, :, 0, 0] * (1.0 / math.sqrt(kin)) * inv_sqrt2
            d['Ws'] = Ws[:, :nreal]
            d['cs'] = Ws[:, nreal:].sum(dim=1) * float(si)
        else:
            d['Ws'] = None
            d['cs'] = torch.zeros_like(d['b1'])
        blocks.append(d)
    Wrgb = _modulated(p, 'torgb.conv', style, False)
    return {'blocks': blocks, 'Wrgb': Wrgb, 'brgb': p['torgb.bias'].reshape(-1)}


def _image_vec(f, gain=1.0):
    """gain: sqrt(2) when the activation gain of conv1 / conv2 is folded into weights + biases."""
    v = []
    for d in f['blocks']:
        v += [d['b1'] * gain, d['b2'] * gain, d['b3'], d['cs']]
    v += [f['Wrgb'].reshape(-1), f['brgb']]
    return torch.cat(v).to(torch.float32).contiguous()


def pack_image(module, si, precision):
    f = fold_image(module, si)
    segs = []
    if precision == PREC_FP32:
        for i, d in enumerate(f['blocks']):
            hk = 256 if i > 0 else 0          # columns fed by the running activation
            for W in (d['W1'],):
                if hk: segs.append(_seg_fp32(W, 0, hk))
                if i < 3: segs.append(_seg_fp32(W, hk, hk + 64))
            segs.append(_seg_fp32(d['W2'], 0, 256))
            segs.append(_seg_fp32(d['W3'], 0, 256))
            if d['Ws'] is not None:
                if hk: segs.append(_seg_fp32(d['Ws'], 0, hk))
                segs.append(_seg_fp32(d['Ws'], hk, hk + 64))
        gemm = torch.cat(segs).to(torch.float32).contiguous()
    elif precision == PREC_BF16X3:
        # consumption order of csrc/decode_umma.cu: per block skip first, then conv1..3.
        # lrelu(x)*sqrt2 == lrelu(x*sqrt2): conv1 / conv2 carry their activation gain in W and b.
        gain = math.sqrt(2.0)
        for i, d in enumerate(f['blocks']):
            hk = 256 if i > 0 else 0
            if d['Ws'] is not None:
                if hk: segs.append(umma_kstep_blocks(d['Ws'], 0, hk))
                segs.append(umma_kstep_blocks(d['Ws'], hk, hk + 64))
            if hk: segs.append(umma_kstep_blocks(d['W1'] * gain, 0, hk))
            if i < 3: segs.append(umma_kstep_blocks(d['W1'] * gain, hk, hk + 64))
            segs.append(umma_kstep_blocks(d['W2'] * gain, 0, 256))
            segs.append(umma_kstep_blocks(d['W3'], 0, 256))
        segs.append(umma_kstep_blocks(f['Wrgb'], 0, 256, n_pad=16))
        return Packed(precision, torch.cat(segs).contiguous(), _image_vec(f, gain))
    else:
        raise ValueError(f"unknown precision {precision}")
    return Packed(precision, gemm, _image_vec(f))


# ---------------------------------------------------------------------------
# ResnetBlockFC decoders: occupancy MLP3D (mlp.py:69-111), video MLPVideo (:114-157)
# ---------------------------------------------------------------------------
def _pack_resnet_chain(p, kx, precision):
    if precision != PREC_FP32:
        raise ValueError("ResnetBlockFC decoders are packed for the fp32 kernels only in this build")
    segs, vec = [], []
    # R1
    W0, Ws, W1 = p['net_res1.fc_0.weight'], p['net_res1.shortcut.weight'], p['net_res1.fc_1.weight']
    nh1 = W0.shape[0]
    segs += [_seg_fp32(W0, 0, kx), _seg_fp32(Ws, 0, kx), _seg_fp32(W1, 0, nh1)]
    vec += [p['net_res1.fc_0.bias'], p['net_res1.fc_1.bias']]
    for i in (2, 3):
        W0, Ws, W1 = (p[f'net_res{i}.fc_0.weight'], p[f'net_res{i}.shortcut.weight'], p[f'net_res{i}.fc_1.weight'])
        segs += [_seg_fp32(W0, 0, 256), _seg_fp32(W0, 256, 256 + kx),
                 _seg_fp32(Ws, 0, 256), _seg_fp32(Ws, 256, 256 + kx), _seg_fp32(W1, 0, 256)]
        vec += [p[f'net_res{i}.fc_0.bias'], p[f'net_res{i}.fc_1.bias']]
    segs += [_seg_fp32(p['net_res4.fc_0.weight'], 0, 256), _seg_fp32(p['net_res4.fc_1.weight'], 0, 256)]
    vec += [p['net_res4.fc_0.bias'], p['net_res4.fc_1.bias']]
    return segs, vec


def pack_occupancy(module, precision=PREC_FP32):
    p = _params64(module)
    segs, vec = _pack_resnet_chain(p, 64, precision)
    vec[1] = vec[1] + p['net_p.bias']                       # net_p bias rides on R1.fc_1's
    vec += [p['net_p.weight'].t().contiguous().reshape(-1),  # [3][256]
            p['net_out.weight'].reshape(-1), p['net_out.bias'].reshape(-1)]
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous(),
                  torch.cat(vec).to(torch.float32).contiguous())


def pack_video(module, precision=PREC_FP32):
    p = _params64(module)
    segs, vec = _pack_resnet_chain(p, 192, precision)
    vec += [p['net_out.weight'].reshape(-1), p['net_out.bias'].reshape(-1)]
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous(),
                  torch.cat(vec).to(torch.float32).contiguous())


# ---------------------------------------------------------------------------
# NeRF MLP (mlp.py:199-281), D=6, W=256, skips=[2,4], xyz 159, dir 27
# ---------------------------------------------------------------------------
def pack_nerf(module, precision=PREC_FP32):
    if precision != PREC_FP32:
        raise ValueError("MLPNeRF is packed for the fp32 kernels only in this build")
    if (module.D, module.W, module.in_channels_xyz, module.in_channels_dir, list(module.skips)) != (6, 256, 159, 27, [2, 4]):
        raise NotImplementedError(
            "the fused NeRF kernel is specialised for D=6, W=256, in_channels_xyz=159, "
            "in_channels_dir=27, skips=[2,4] (configs/d2c-vae/srn_cars.yaml:45-49)")
    p = _params64(module)
    segs, vec = [], []
    for i in range(6):
        W = p[f'xyz_encoding_{i + 1}.0.weight']
        if i == 0:
            segs.append(_seg_fp32(W, 0, 159, 160))
        elif i in (2, 4):
            segs += [_seg_fp32(W, 0, 159, 160), _seg_fp32(W, 159, 415)]
        else:
            segs.append(_seg_fp32(W, 0, 256))
        vec.append(p[f'xyz_encoding_{i + 1}.0.bias'])
    segs.append(_seg_fp32(p['xyz_encoding_final.weight'], 0, 256))
    Wd = p['dir_encoding.0.weight']
    segs += [_seg_fp32(Wd, 0, 256), _seg_fp32(Wd, 256, 283, 32)]
    vec += [p['xyz_encoding_final.bias'], p['dir_encoding.0.bias'],
            p['sigma.weight'].reshape(-1), p['sigma.bias'].reshape(-1),
            p['rgb.0.weight'].reshape(-1), p['rgb.0.bias'].reshape(-1)]
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous(),
                  torch.cat(vec).to(torch.float32).contiguous())
