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
* ``PREC_F16F8``  : per 16-wide K step ``[fp16(S W) | FP8 block of the step's 32-wide pair]``
  (S = 4096; even step: ``e4m3(W)``, odd step: ``e4m3(S W - fp16(S W))``) in the same
  core-matrix order and slot size (csrc/umma.cuh, "f16f8" scheme).

Both formats share one ``vec`` blob of fp32 vectors (biases, folded constants,
the narrow output heads that run in the epilogue).
"""
import math
import os
from dataclasses import dataclass

import torch

from ._lib import PREC_BF16X3, PREC_F16F8, PREC_FP32


@dataclass
class Packed:
    precision: int
    gemm: torch.Tensor   # flat device tensor (fp32 or int16 bit patterns of bf16)
    vec: torch.Tensor    # flat fp32 device tensor
    program: torch.Tensor = None        # bf16x3: int32 MMA program (device)
    program_host: torch.Tensor = None   # same, host copy (validated by the launcher)
    pair: bool = False                  # bf16x3: stream packed for CTA pairs ([half 0 | half 1] per K step)
    noise_active: bool = False          # image: some NoiseInjection.weight is non-zero (the decode needs a noise source)
    ts: bool = False                    # image, f16f8: program of the TMEM-resident-activation kernel (image_umma_kernel<.., TS>)
    vec_host: torch.Tensor = None       # host copy of vec (ts: the ToRGB weights travel as a kernel parameter)


class UmmaProgram:
    """Builds the MMA program and the weight stream of a tcgen05 kernel TOGETHER, so the
    producer warp's ring is consumed in exactly the order it is filled (csrc/decode_umma.cu).

    Op encoding (uint32): bits[1:0] kind (0 UNIT, 1 WAIT, 2 COMMIT, 3 END).
    UNIT = a run of consecutive 16-wide K steps of one 128-row x N block (3 MMAs per step):
    [3:2] N code (0:128, 1:256, 2:16, 3:64), [4] accumulate flag of the FIRST step (later steps always
    accumulate), [7:5] accumulator column / 64, [15:8] first A-hi K group, [23:16] first A-lo K group
    (K groups = 8 columns x 128 rows = 2 KB, counted from the A region base in shared memory; each
    step advances both by 2), [28:24] number of steps - 1.
    WAIT: bits [4:2] select one of 8 operand barriers (e.g. quarter q of the previous epilogue's output is ready).
    COMMIT: bits [3:2] select one of 4 completion barriers the epilogue threads wait on.
    """
    NCODE = {128: 0, 256: 1, 16: 2, 64: 3}

    def __init__(self, pair=False, scheme='bf16x3'):
        self.ops = []
        self.segs = []
        self.pair = pair      # CTA pairs: each K step is stored as [rows 0..N/2-1 | rows N/2..N-1]
        # 'f16f8': same 16-wide steps, always in 32-wide pairs (fp16 MMA + one K = 32 e4m3 MMA each: r8 x w8 on the even
        # step, a8 x s8 on the odd one); [15:8] = first fp16 K group, [23:16] = first K group of the FP8 operands
        # [r8 r8 a8 a8] per pair -- both advance 2 per step like the bf16 hi / lo groups; pairs only.
        self.scheme = scheme
        assert scheme in ('bf16x3', 'f16f8') and (scheme == 'bf16x3' or pair)

    def block(self, W, a_hi_kg, a_lo_kg, acc_col, first, n_pad=None, a_in_tmem=False, half=False):
        """acc[:, acc_col:acc_col+N] (+)= A[:, K] @ W.T ; W is (N, K), K zero-padded to 16.  a_in_tmem: the A operand
        lives in tensor memory (K group g = TMEM columns 4g..4g+3), CTA-pair kernels only (op bit 29).
        half (f16f8, N = 128, K a multiple of 64; op bit 30): a half-width unit of an N-split layer.  The stream then holds,
        per 32-wide step pair, [CTA 0: step 0 | step 1][CTA 1: step 0 | step 1] (8 KB each), so one 8 KB ring slot = one
        pair and the two-slot hand-shake unit covers 64 K columns (the engine copies whole 8 KB slots either way)."""
        n = W.shape[0] if n_pad is None else n_pad
        assert n in self.NCODE and acc_col % 64 == 0 and acc_col + n <= 512
        if self.scheme == 'f16f8':
            k32 = (W.shape[1] + 31) // 32
            assert 1 <= k32 <= 16 and a_hi_kg + 4 * k32 <= 256 and a_lo_kg + 4 * k32 <= 256
            assert not a_in_tmem or 4 * (a_lo_kg + 4 * k32) <= 512
            assert not half or (n == 128 and W.shape[1] % 64 == 0)
            self.ops.append(0 | (self.NCODE[n] << 2) | ((0 if first else 1) << 4) | ((acc_col // 64) << 5)
                            | (a_hi_kg << 8) | (a_lo_kg << 16) | ((2 * k32 - 1) << 24) | ((1 if a_in_tmem else 0) << 29)
                            | ((1 if half else 0) << 30))
            Wp = W.new_zeros(n, 32 * k32)
            Wp[:W.shape[0], :W.shape[1]] = W
            # per CTA half: (steps, bytes per step); half-width: regrouped per step pair
            halves = [f16f8_kstep_blocks(Wp[h * (n // 2):(h + 1) * (n // 2)]).reshape(k32 if half else 2 * k32, -1)
                      for h in (0, 1)]
            self.segs.append(torch.cat(halves, dim=1).reshape(-1))
            return
        k16 = (W.shape[1] + 15) // 16
        assert 1 <= k16 <= 32 and a_hi_kg + 2 * k16 <= 256 and a_lo_kg + 2 * k16 <= 256
        self.ops.append(0 | (self.NCODE[n] << 2) | ((0 if first else 1) << 4) | ((acc_col // 64) << 5)
                        | (a_hi_kg << 8) | (a_lo_kg << 16) | ((k16 - 1) << 24) | ((1 if a_in_tmem else 0) << 29))
        assert not a_in_tmem or (self.pair and 4 * (a_lo_kg + 2 * k16) <= 512)
        if not self.pair:
            self.segs.append(umma_kstep_blocks(W, 0, W.shape[1], n_pad=n_pad))
        else:
            Wp = W
            if n != W.shape[0]:
                Wp = W.new_zeros(n, W.shape[1])
                Wp[:W.shape[0]] = W
            halves = [umma_kstep_blocks(Wp[h * (n // 2):(h + 1) * (n // 2)], 0, W.shape[1]).reshape(k16, -1)
                      for h in (0, 1)]
            self.segs.append(torch.cat(halves, dim=1).reshape(-1))

    def wait(self, which):
        assert 0 <= which < 8
        self.ops.append(1 | (which << 2))

    def commit(self, which=0):
        assert 0 <= which < 4
        self.ops.append(2 | (which << 2))

    def finish(self, device):
        """-> (weight stream on `device`, program on `device`, program on the host)."""
        ops = torch.tensor(self.ops + [3, 3, 3, 3], dtype=torch.int64).to(torch.int32)
        return torch.cat(self.segs).contiguous().to(device), ops.to(device), ops.contiguous()


F8_SCALE = 4096.0


class F16F8RangeError(ValueError):
    """A weight does not fit fp16 after the 2^12 scaling of the f16f8 scheme (|w| >= 16)."""


def f16f8_kstep_blocks(W):
    """Pack W (N, K), K a multiple of 32, for the f16f8 tcgen05 kernels: a flat uint8 tensor, per 16-wide K step
    [fp16(S W[:, step]): 2 K groups x N rows x 8 halves | FP8: 2 K groups x N rows x 16 bytes], where the FP8 block
    covers the 32-wide PAIR the step belongs to: e4m3(W) on the even step, e4m3(S W - fp16(S W)) on the odd step.
    Raises if a weight does not fit fp16 after scaling (|W| >= 16)."""
    W = W.to(torch.float32)
    n, k = W.shape
    assert k % 32 == 0
    ws = W * F8_SCALE
    if float(ws.abs().max()) > 65504.0:
        raise F16F8RangeError("f16f8 packing: |weight| * 4096 exceeds the fp16 range; use precision='bf16x3'")
    w16 = ws.to(torch.float16)
    res = ws - w16.to(torch.float32)
    w8 = W.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    s8 = res.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    p16 = w16.reshape(n, k // 16, 2, 8).permute(1, 2, 0, 3).contiguous().view(torch.uint8).reshape(k // 32, 2, -1)
    p8 = torch.stack([t.reshape(n, k // 32, 2, 16).permute(1, 2, 0, 3).contiguous().view(torch.uint8).reshape(k // 32, -1)
                      for t in (w8, s8)], dim=1)                      # (pairs, even / odd step, bytes)
    return torch.cat([p16, p8], dim=2).reshape(-1)


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
    """The module's state dict as float64 HOST tensors.  Folding and packing run on the CPU (a few ms of tiny matrix
    algebra) and the packed blobs are shipped with three H2D copies: done on the device, the same algebra is ~700 tiny
    kernel launches per pack.  One device-side cat + one D2H copy fetches all parameters."""
    sd = module.state_dict()
    keys = list(sd)
    if not keys:
        return {}
    flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys]).cpu().to(torch.float64)
    out, o = {}, 0
    for k in keys:
        n = sd[k].numel()
        out[k] = flat[o:o + n].reshape(sd[k].shape)
        o += n
    return out


def module_device(module):
    return next(module.parameters()).device


def tensor_version(t):
    """The autograd version counter of `t`, or None for inference-mode tensors (they do not track one)."""
    try:
        return t._version
    except RuntimeError:
        return None


def param_fingerprint(module):
    """Cache key of the packed weights: identity (data_ptr, shape, version counter) of every parameter PLUS a content
    digest (L1 and L2 norm of every parameter, two fused multi-tensor kernels).  The digest catches writes that bypass the
    version counter -- `p.data.copy_(...)`, the idiom of the reference's EMA swap (models/ema.py LitEma.copy_to / restore) --
    which identity alone would miss.  Returns (identity tuple, digest tensor on the parameters' device)."""
    params = [p.detach() for p in module.parameters()]
    ident = tuple((p.data_ptr(), tensor_version(p), tuple(p.shape)) for p in params)
    if not params:
        return ident, torch.zeros(0)
    digest = torch.stack(torch._foreach_norm(params, 1) + torch._foreach_norm(params, 2))
    return ident, digest


def same_fingerprint(a, b):
    if a is None or b is None or a[0] != b[0]:
        return False
    return a[1].shape == b[1].shape and bool(torch.equal(a[1], b[1]))


# ---------------------------------------------------------------------------
# image MLP (models/d2c_vae/mlp.py:12-66)
# ---------------------------------------------------------------------------
def image_style(p, si, ch=256):
    """style = time_mlp(si): SinusoidalPosEmb (blocks.py:11-23) -> Linear ->
    exact-erf GELU -> Linear.  Identical for every batch item (mlp.py:46-47)."""
    dim = ch // 4
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float64) * -(math.log(10000) / (half - 1)))
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
        # NoiseInjection.weight of the three StyledConvs (blocks.py:286-297): out = conv + weight * noise, before the bias
        d['nw'] = torch.cat([p[f'{pre}.conv{j}.noise.weight'].reshape(1) for j in (1, 2, 3)])
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
    """gain: sqrt(2) when the activation gain of conv1 / conv2 is folded into weights + biases (the noise term sits inside
    the activation, so its weight takes the same gain).  Layout: 4 x [b1 b2 b3 cs] (256 each) | Wrgb (3 x 256) | brgb (3) |
    noise weights (12: conv1, conv2, conv3 of each block)."""
    v = []
    for d in f['blocks']:
        v += [d['b1'] * gain, d['b2'] * gain, d['b3'], d['cs']]
    v += [f['Wrgb'].reshape(-1), f['brgb']]
    for d in f['blocks']:
        v += [d['nw'] * torch.tensor([gain, gain, 1.0], dtype=d['nw'].dtype)]
    return torch.cat(v).to(torch.float32).contiguous()


def _noise_active(f):
    return any(float(d['nw'].abs().max()) != 0.0 for d in f['blocks'])


def _pack_image_ts(f, dev):
    """Program + stream of csrc/decode_umma.cu::image_ts_kernel (f16f8, CTA pairs): the 256-wide running activation H lives
    in TENSOR memory (A-from-TMEM MMAs run at the tensor pipe's full rate, 136 cycles per 256 x 256 x 16 step; with the A
    operand in shared memory the same step takes 162-173) next to ONE accumulator:
      TMEM columns [0, 256) acc | [256, 384) H fp16 (K group g = columns 256 + 4g) | [384, 512) H FP8, per 32 columns [r8 x8 | a8 x8]
      shared-memory K groups 64..79: the PE features X (SS units);  K groups 0..63: the parked skip values (epilogue threads only).
    Every GEMM group is N-split, [a: K 0..127][b: K 0..127][X part, full width][a: K 128..255] COMMIT 0 [b: K 128..255] COMMIT 1
    (a / b = output columns 0..127 / 128..255 as half-width units), so the epilogue threads convert and publish a's columns
    while the tensor core still works on b's -- in place: operand columns 0..127 are dead once a's second K-half has run.
    The skip GEMM of a block is its own group in front of conv1 (both read the block's input); its result is parked in shared
    memory by the thread that will add it in conv3's epilogue.  ToRGB is evaluated by the epilogue threads in fp32.
    Operand barriers: 0..3 = H quarters published (groups that follow a publishing stage), 5 / 4 = accumulator columns
    0..127 / 128..255 drained (every group)."""
    gain = math.sqrt(2.0)
    HT16, HT8, XH, XL = 64, 96, 64, 72
    P = UmmaProgram(pair=True, scheme='f16f8')
    nsplit = 'full'            # csrc/image_ts_issuer.cuh writes this sequence out as straight-line code: keep them in step

    def group(W, k_h, k_x, published):
        """W: (256, k_h + k_x); k_h in (0, 256) columns from H (TMEM), k_x in (0, 64) from X (shared memory)."""
        waited = set()

        def wait(*which):
            for w in which:
                if w not in waited:
                    P.wait(w)
                    waited.add(w)

        def over_h(h, quarters, first):
            rows = slice(128 * h, 128 * h + 128) if h is not None else slice(0, 256)
            # maximal runs of quarters that need no new WAIT in between become one unit
            runs = []
            for q in quarters:
                if (published and q not in waited) or not runs:
                    runs.append([q])
                else:
                    runs[-1].append(q)
            for r in runs:
                if published:
                    wait(r[0])
                P.block(W[rows, 64 * r[0]:64 * r[-1] + 64], HT16 + 8 * r[0], HT8 + 8 * r[0], 0 if h is None else 128 * h,
                        first and r[0] == quarters[0], a_in_tmem=True, half=h is not None)

        if k_h:
            if nsplit == 'full':
                wait(5)
                over_h(0, (0, 1), True)
                wait(4)
                over_h(1, (0, 1), True)
            else:
                wait(5, 4)
                over_h(None, (0, 1), True)
            if k_x:
                P.block(W[:, k_h:k_h + k_x], XH, XL, 0, False)
            over_h(0, (2, 3), False)
            P.commit(0)
            over_h(1, (2, 3), False)
            P.commit(1)
        else:                                  # block 0: X only
            if published:
                wait(0, 1, 2, 3)
            wait(5)
            P.block(W[0:128, 0:k_x], XH, XL, 0, True, half=True)
            P.commit(0)
            wait(4)
            P.block(W[128:256, 0:k_x], XH, XL, 128, True, half=True)
            P.commit(1)
        assert not published or waited >= {0, 1, 2, 3}

    for i, d in enumerate(f['blocks']):
        k_h, k_x = (256 if i > 0 else 0), (64 if i < 3 else 0)
        if d['Ws'] is not None:
            group(d['Ws'], k_h, k_x, True)                    # skip -> parked
        group(d['W1'] * gain, k_h, k_x, d['Ws'] is None)      # conv1 (follows the skip group: nothing new was published)
        group(d['W2'] * gain, 256, 0, True)
        group(d['W3'], 256, 0, True)
    gemm, prog_dev, prog_host = P.finish(dev)
    vec_host = _image_vec(f, gain)
    return Packed(PREC_F16F8, gemm, vec_host.to(dev), prog_dev, prog_host, True, _noise_active(f), ts=True, vec_host=vec_host)


def pack_image(module, si, precision, pair=True):
    f = fold_image(module, si)
    dev = module_device(module)
    segs = []
    if precision == PREC_F16F8 and pair and os.environ.get('DDMI_B200_IMAGE_TS', '1') != '0':
        return _pack_image_ts(f, dev)
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
    elif precision in (PREC_BF16X3, PREC_F16F8):
        # Program of csrc/decode_umma.cu::image_umma_kernel.  (PREC_F16F8: same program and K-group numbers -- a
        # 64-column quarter is two 32-wide steps, its fp16 groups are 8q.., its FP8 groups 32 + 8q..)  K groups of the A region:
        # [H hi 0..31 | H lo 32..63 | X hi 64..71 | X lo 72..79]; acc1 = TMEM columns 0..255, acc2 = 256..511.
        # lrelu(x)*sqrt2 == lrelu(x*sqrt2): conv1 / conv2 carry their activation gain in W and b.
        # Every GEMM group is [WAIT q, K steps over H columns 64q..64q+63] for q = 0..3, then COMMIT: the
        # previous epilogue drains the whole accumulator, then publishes its output quarter by quarter
        # (operand barrier q).  acc2 (skip) is only written after WAIT 3 because the previous conv3
        # epilogue still reads it until then.
        gain = math.sqrt(2.0)
        HH, HL, XH, XL = 0, 32, 64, 72
        P = UmmaProgram(pair=pair, scheme='f16f8' if precision == PREC_F16F8 else 'bf16x3')
        # schedule: how many operand quarters must be published before a K run starts
        #   'quarters' : run q needs barriers 0..q      (finest overlap)
        #   'halves'   : runs 0,1 need 0..1; runs 2,3 need 0..3
        #   'serial'   : every run needs all four       (no overlap; reference schedule)
        sched = os.environ.get('DDMI_B200_SCHEDULE', 'quarters')
        need = {'quarters': (1, 2, 3, 4), 'halves': (2, 2, 4, 4), 'serial': (4, 4, 4, 4)}[sched]
        # N split (f16f8): the second K-half of a layer runs as two half-width (N = 128) passes, output columns 0..127 first
        # (COMMIT 0), then 128..255 (COMMIT 1).  The epilogue threads convert and publish the first half of the layer's
        # output while the tensor core still works on the second, so the next layer's first K-half can start as soon as that
        # finishes: the exposed drain -> convert -> publish chain of every layer shrinks to the accumulator hand-over.
        # Publishing in place is safe because operand columns 0..127 are only read by the first K-half ('full': every K-half
        # is N-split, [a: K 0..127][b: K 0..127][a: K 128..255][b: K 128..255]; 'off': one COMMIT pair at the end).
        # Operand barrier 4 = "accumulator columns 128..255 drained": waited before the first unit that overwrites them.
        nsplit = os.environ.get('DDMI_B200_NSPLIT', 'off') if precision == PREC_F16F8 else 'off'
        assert nsplit in ('mixed', 'full', 'off')
        state = {'waited': set()}

        def wait(*which):              # consume operand barriers (each exactly once per GEMM group)
            for w in which:
                if w not in state['waited']:
                    P.wait(w)
                    state['waited'].add(w)

        def wait_upto(k):
            wait(*range(k))

        def end_group(split_done=False):
            wait(0, 1, 2, 3, 4)        # every group consumes each barrier exactly once
            if not split_done:
                P.commit(0)
            P.commit(1)
            state['waited'] = set()

        def dense256(W, acc, n_pad=None, first=True, quarters=(0, 1, 2, 3)):          # K = 256 from H, full width
            for q in quarters:
                wait_upto(need[q])
                wait(4)
                P.block(W[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, acc, first and q == quarters[0], n_pad=n_pad)

        def half_pass(W, h, quarters, first):      # output columns 128h..128h+127 over the given K quarters
            Wh = W[128 * h:128 * h + 128]
            if h == 1:
                wait(4)
            for q in quarters:
                wait_upto(need[q])
                P.block(Wh[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 128 * h, first and q == quarters[0], half=True)

        def layer256(W, first=True, mid=None):
            """One 256 -> 256 layer into acc1 and its two commits; `mid()` emits extra units (the skip GEMM) that must read
            the operand before either half of it is overwritten."""
            if nsplit == 'off':
                dense256(W, 0, first=first)
                if mid: mid()
                end_group()
                return
            if nsplit == 'mixed':
                dense256(W, 0, first=first, quarters=(0, 1))
            else:
                half_pass(W, 0, (0, 1), first)
                half_pass(W, 1, (0, 1), first)
            if mid: mid()
            half_pass(W, 0, (2, 3), False)
            wait(0, 1, 2, 3, 4)
            P.commit(0)
            half_pass(W, 1, (2, 3), False)
            end_group(split_done=True)

        for i, d in enumerate(f['blocks']):
            W1 = d['W1'] * gain
            if i == 0:                 # x = PE only (K = 64)
                wait(0, 1, 2, 3, 4)
                P.block(d['Ws'][:, 0:64], XH, XL, 256, True)
                P.block(W1[:, 0:64], XH, XL, 0, True)
                end_group()
            elif i < 3:                # x = [h (256) | PE (64)], conv1 -> acc1, skip -> acc2
                wait_upto(need[0])     # the accumulator is drained once barriers 0 and 4 have completed
                wait(4)
                P.block(W1[:, 256:320], XH, XL, 0, True)

                def skip(d=d):
                    wait(0, 1, 2, 3)   # acc2 is read by the previous conv3 epilogue until its last quarter
                    P.block(d['Ws'][:, 0:256], HH, HL, 256, True)
                    P.block(d['Ws'][:, 256:320], XH, XL, 256, False)
                layer256(W1, first=False, mid=skip)
            else:
                layer256(W1)
            layer256(d['W2'] * gain)
            layer256(d['W3'])
        dense256(f['Wrgb'], 0, n_pad=16)   # ToRGB: N = 16 block (3 real rows)
        end_group()
        gemm, prog_dev, prog_host = P.finish(dev)
        return Packed(precision, gemm, _image_vec(f, gain).to(dev), prog_dev, prog_host, pair, _noise_active(f))
    else:
        raise ValueError(f"unknown precision {precision}")
    return Packed(precision, gemm.to(dev), _image_vec(f).to(dev), noise_active=_noise_active(f))


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


def _pack_occupancy_umma(p, pair, precision=PREC_BF16X3, dev='cpu'):
    """Program + stream + vec of csrc/decode_umma_occ.cuh.  A-region K groups: H hi 0..31, H lo 32..63,
    Xa (raw PE) hi 64..71 / lo 80..87, Xb (relu PE) hi 72..79 / lo 88..95; acc1 = TMEM cols 0.., acc2 = 256..
    Each ResnetBlockFC with a shortcut is three GEMM groups: shortcut on RAW h -> acc2, fc_0 on relu(h) -> acc1,
    fc_1 on relu(net) accumulated ONTO acc2.  K runs over h follow the epilogue's quarter-by-quarter publication (operand
    barriers A0..A3); the PE operands of R2 / R3 have their own barrier (A4) and run between the two h phases."""
    HH, HL, XAH, XBH, XAL, XBL = 0, 32, 64, 72, 80, 88    # f16f8: "hi" = fp16 K groups, "lo" = FP8 K groups (same numbers)
    P = UmmaProgram(pair=pair, scheme='f16f8' if precision == PREC_F16F8 else 'bf16x3')

    def wait_all():
        for q in range(4):
            P.wait(q)

    def over_h(W, acc, first):
        for q in range(4):
            P.wait(q)
            P.block(W[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, acc, first and q == 0)

    # R1: x = PE(64).  fc_0 and the shortcut share one group so both PE buffers are released together
    wait_all()
    P.block(p['net_res1.fc_0.weight'], XBH, XBL, 0, True)                       # N = 64
    P.block(p['net_res1.shortcut.weight'], XAH, XAL, 256, True)
    P.commit()
    wait_all()
    P.block(p['net_res1.fc_1.weight'], HH, HL, 256, False)                      # K = 64 hidden, onto the shortcut
    # net_p (Linear(3, 256) on the query point, mlp.py:103) rides on the same accumulator: the epilogue threads publish
    # [p | 0] as columns 64..95 of H next to R1's hidden layer (three FMAs per output in the epilogue cost ~4 K cycles of the
    # stage every other layer waits for)
    Wp = p['net_p.weight']
    P.block(torch.cat([Wp, Wp.new_zeros(Wp.shape[0], 32 - Wp.shape[1])], dim=1), HH + 8, HL + 8, 256, False)
    P.commit()
    for i in (2, 3):                                                            # x = [h(256) | PE(64)]
        Ws, W0, W1 = (p[f'net_res{i}.shortcut.weight'], p[f'net_res{i}.fc_0.weight'], p[f'net_res{i}.fc_1.weight'])
        over_h(Ws, 256, True)
        P.commit()                                # raw h consumed: the epilogue threads rewrite H with relu(h) ...
        P.wait(4)                                 # ... while the PE parts (gathered behind the raw-h publication, barrier A4)
        P.block(Ws[:, 256:320], XAH, XAL, 256, False)                           # keep the tensor core busy;
        P.block(W0[:, 256:320], XBH, XBL, 0, True)                              # fc_0's accumulator starts here
        P.commit(1)                               # PE buffers free: the next gather runs under fc_0's MMAs over h
        over_h(W0, 0, False)
        P.commit()
        over_h(W1, 256, False)                                                  # x_s + dx accumulate in TMEM
        P.commit()
    over_h(p['net_res4.fc_0.weight'], 0, True)                                  # R4: identity shortcut stays in acc2
    P.commit()
    over_h(p['net_res4.fc_1.weight'], 256, False)
    P.commit()
    vec = torch.cat([p['net_res1.fc_0.bias'], p['net_res1.fc_1.bias'] + p['net_p.bias'],
                     p['net_p.weight'].t().contiguous().reshape(-1),
                     p['net_res2.fc_0.bias'], p['net_res2.fc_1.bias'], p['net_res3.fc_0.bias'], p['net_res3.fc_1.bias'],
                     p['net_res4.fc_0.bias'], p['net_res3.fc_1.bias'] + p['net_res4.fc_1.bias'],
                     p['net_out.weight'].reshape(-1), p['net_out.bias'].reshape(-1)]).to(torch.float32).contiguous()
    gemm, prog_dev, prog_host = P.finish(dev)
    return Packed(precision, gemm, vec.to(dev), prog_dev, prog_host, pair)


def pack_occupancy(module, precision=PREC_FP32, pair=True):
    p = _params64(module)
    dev = module_device(module)
    if precision in (PREC_BF16X3, PREC_F16F8):
        return _pack_occupancy_umma(p, pair, precision, dev)
    segs, vec = _pack_resnet_chain(p, 64, precision)
    vec[1] = vec[1] + p['net_p.bias']                       # net_p bias rides on R1.fc_1's
    vec += [p['net_p.weight'].t().contiguous().reshape(-1),  # [3][256]
            p['net_out.weight'].reshape(-1), p['net_out.bias'].reshape(-1)]
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous().to(dev),
                  torch.cat(vec).to(torch.float32).contiguous().to(dev))


def _pack_video_umma(p, pair, precision=PREC_BF16X3, dev='cpu'):
    """Program + stream + vec of csrc/decode_umma_video.cuh (the protocol is spelled out there).  A-region K groups
    as for occupancy: H hi 0..31 / lo 32..63, Xa (raw piece) hi 64..71 / lo 80..87, Xb (relu piece) hi 72..79 / lo 88..95.
    Operand barriers: 0..3 raw-h quarters / R1 pieces / later pieces, 4..7 relu-h and net quarters.
    Completion barriers: D0 = "accumulators final for this phase", D1 = "piece consumed, Xa / Xb free" (R2 / R3 only)."""
    HH, HL, XAH, XBH, XAL, XBL = 0, 32, 64, 72, 80, 88    # f16f8: "hi" = fp16 K groups, "lo" = FP8 K groups (same numbers)
    P = UmmaProgram(pair=pair, scheme='f16f8' if precision == PREC_F16F8 else 'bf16x3')

    def piece(Ws, W0, col0, first, kg=None, fc0_first=None):
        """one 64-wide piece: fc_0 (relu piece, acc1) and shortcut (raw piece, acc2); kg = (raw hi, raw lo, relu hi, relu lo)
        K groups of the piece (default: the X region); first / fc0_first: the shortcut / fc_0 accumulator starts here"""
        ah, al, bh, bl = kg or (XAH, XAL, XBH, XBL)
        P.block(W0[:, col0:col0 + 64], bh, bl, 0, first if fc0_first is None else fc0_first)
        P.block(Ws[:, col0:col0 + 64], ah, al, 256, first)

    # ---- R1: x = [xy | yt | xt] of scale 0.  The activation buffer H is idle until R1's hidden layer is published, so pieces
    # 1 and 2 sit in H (piece j: raw K groups 16(j-1).., relu 16(j-1)+8..) next to piece 0 in the X region: the three pieces
    # run back to back without a hand-shake per piece.  fc_0 (192 outputs) is zero-padded to one N = 256 block: one unit
    # per piece instead of an N = 128 and an N = 64 one (the issuer's cost per unit, not the MMAs, is what counted).
    Ws, W0, W1 = p['net_res1.shortcut.weight'], p['net_res1.fc_0.weight'], p['net_res1.fc_1.weight']
    W0 = torch.cat([W0, W0.new_zeros(64, W0.shape[1])])
    for j in range(3):
        P.wait(j)
        piece(Ws, W0, 64 * j, j == 0, None if j == 0 else (HH + 16 * (j - 1), HL + 16 * (j - 1), HH + 16 * (j - 1) + 8, HL + 16 * (j - 1) + 8))
    P.commit(0)
    for q in range(3):                                 # fc_1: K = 192 hidden, onto the shortcut
        P.wait(4 + q)
        P.block(W1[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 256, False)
    P.commit(0)
    # ---- R2, R3: x = [h (256) | xy | yt | xt]
    for i in (2, 3):
        Ws, W0, W1 = (p[f'net_res{i}.shortcut.weight'], p[f'net_res{i}.fc_0.weight'], p[f'net_res{i}.fc_1.weight'])
        for q in range(4):                             # shortcut over RAW h (piece 0 was stored before h: same barriers)
            P.wait(q)
            P.block(Ws[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 256, q == 0)
        P.commit(0)                                    # raw h consumed: the epilogue threads rewrite H with relu(h) ...
        piece(Ws, W0, 256, False, fc0_first=True)       # ... while piece 0 (fc_0 starts its accumulator here) keeps the pipe busy
        P.commit(1)
        # relu(h) quarters with the two remaining pieces in between: each piece's store (after the previous D1) hides
        # behind a quarter's MMAs
        P.wait(4)
        P.block(W0[:, 0:64], HH, HL, 0, False)
        P.wait(5)
        P.block(W0[:, 64:128], HH + 8, HL + 8, 0, False)
        P.wait(0)
        piece(Ws, W0, 320, False)
        P.commit(1)
        for q in (2, 3):
            P.wait(4 + q)
            P.block(W0[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 0, False)
        P.wait(1)
        piece(Ws, W0, 384, False)
        P.commit(0)
        for q in range(4):                             # fc_1 onto the shortcut
            P.wait(4 + q)
            P.block(W1[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 256, False)
        P.commit(0)
    # ---- R4: identity shortcut (acc2 keeps accumulating)
    for q in range(4):
        P.wait(q)
        P.block(p['net_res4.fc_0.weight'][:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 0, q == 0)
    P.commit(0)
    for q in range(4):
        P.wait(4 + q)
        P.block(p['net_res4.fc_1.weight'][:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 256, False)
    P.commit(0)
    vec = torch.cat([p['net_res1.fc_0.bias'], p['net_res1.fc_1.bias'], p['net_res2.fc_0.bias'], p['net_res2.fc_1.bias'],
                     p['net_res3.fc_0.bias'], p['net_res3.fc_1.bias'], p['net_res4.fc_0.bias'],
                     p['net_res3.fc_1.bias'] + p['net_res4.fc_1.bias'], p['net_out.weight'].reshape(-1),
                     p['net_out.bias'].reshape(-1)]).to(torch.float32).contiguous()
    gemm, prog_dev, prog_host = P.finish(dev)
    return Packed(precision, gemm, vec.to(dev), prog_dev, prog_host, pair)


def pack_video(module, precision=PREC_FP32, pair=True):
    p = _params64(module)
    dev = module_device(module)
    if precision in (PREC_BF16X3, PREC_F16F8):
        return _pack_video_umma(p, pair, precision, dev)
    segs, vec = _pack_resnet_chain(p, 192, precision)
    vec += [p['net_out.weight'].reshape(-1), p['net_out.bias'].reshape(-1)]
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous().to(dev),
                  torch.cat(vec).to(torch.float32).contiguous().to(dev))


# ---------------------------------------------------------------------------
# NeRF MLP (mlp.py:199-281), D=6, W=256, skips=[2,4], xyz 159, dir 27
# ---------------------------------------------------------------------------
def _pack_nerf_umma(p, precision=PREC_BF16X3, dev='cpu'):
    """Program + stream + vec of csrc/decode_umma_nerf.cuh (CTA pairs).  A-region K groups: H hi 0..31, H lo 32..63,
    X ([latent 96 | gamma(pts) 63 | 0]) hi 64..83, lo 84..103; the 27-wide direction embedding reuses X's first 4
    K groups for the last layer.  One accumulator (TMEM columns 0..255)."""
    HH, HL, XH, XL = 0, 32, 64, 84        # X K groups live in TENSOR memory: columns 256..335 (hi), 336..415 (lo)
    P = UmmaProgram(pair=True, scheme='f16f8' if precision == PREC_F16F8 else 'bf16x3')

    def over_h(W, first, n_pad=None):
        for q in range(4):
            P.wait(q)
            P.block(W[:, 64 * q:64 * q + 64], HH + 8 * q, HL + 8 * q, 0, first and q == 0, n_pad=n_pad)

    def pad_k(W, k):
        out = W.new_zeros(W.shape[0], k)
        out[:, :W.shape[1]] = W
        return out

    for i in range(6):
        W = p[f'xyz_encoding_{i + 1}.0.weight']
        if i == 0:
            for q in range(4):
                P.wait(q)
            P.block(pad_k(W, 160), XH, XL, 0, True, a_in_tmem=True)
        elif i in (2, 4):                      # cat([input_xyz, h]): X part first (ready at once), then h by quarters
            P.wait(0)
            P.block(pad_k(W[:, :159], 160), XH, XL, 0, True, a_in_tmem=True)
            for q in range(4):
                if q:
                    P.wait(q)
                P.block(W[:, 159 + 64 * q:159 + 64 * q + 64], HH + 8 * q, HL + 8 * q, 0, False)
        else:
            over_h(W, True)
        P.commit()
    over_h(p['xyz_encoding_final.weight'], True)
    P.commit()
    Wd = p['dir_encoding.0.weight']            # (128, 283) on cat([final, dir27])
    over_h(Wd[:, :256], True)
    P.block(pad_k(Wd[:, 256:283], 32), XH, XL, 0, False, a_in_tmem=True)
    P.commit()
    z3 = torch.zeros(3, dtype=torch.float64)
    vec = torch.cat([p[f'xyz_encoding_{i + 1}.0.bias'] for i in range(6)]
                    + [p['xyz_encoding_final.bias'], p['dir_encoding.0.bias'], p['sigma.weight'].reshape(-1),
                       p['sigma.bias'].reshape(-1), z3, p['rgb.0.weight'].reshape(-1), p['rgb.0.bias'].reshape(-1)]
                    ).to(torch.float32).contiguous()
    gemm, prog_dev, prog_host = P.finish(dev)
    return Packed(precision, gemm, vec.to(dev), prog_dev, prog_host, True)


def pack_nerf(module, precision=PREC_FP32):
    if (module.D, module.W, module.in_channels_xyz, module.in_channels_dir, list(module.skips)) != (6, 256, 159, 27, [2, 4]):
        raise NotImplementedError(
            "the fused NeRF kernel is specialised for D=6, W=256, in_channels_xyz=159, "
            "in_channels_dir=27, skips=[2,4] (configs/d2c-vae/srn_cars.yaml:45-49)")
    p = _params64(module)
    dev = module_device(module)
    if precision in (PREC_BF16X3, PREC_F16F8):
        return _pack_nerf_umma(p, precision, dev)
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
    return Packed(precision, torch.cat(segs).to(torch.float32).contiguous().to(dev),
                  torch.cat(vec).to(torch.float32).contiguous().to(dev))
