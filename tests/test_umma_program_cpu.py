"""CPU model of the tcgen05 engine's host contract: the MMA program and the weight stream that
ddmi_b200/packing.py emit TOGETHER are interpreted here the way csrc/umma_engine.cuh does
(UNIT = K steps of one 128 x N block read from the stream in order, WAIT / COMMIT = handshakes)
with the image kernel's epilogue sequence restated in torch.  Running it against the oracle checks,
without a GPU, that stream order, K-group addressing, accumulator columns, accumulate flags, the
CTA-pair half split and the folded biases all line up (the GPU tests then only have to prove the
CUDA side)."""
import math
import os

import pytest
import torch

from ddmi_b200 import _lib, packing
from oracle import cases, ddmi_oracle as orc

torch.set_grad_enabled(False)
NCODE = {0: 128, 1: 256, 2: 16, 3: 64}


def unpack_kstep(words, n, pair):
    """Inverse of packing.umma_kstep_blocks for ONE K step: int16 words -> (hi + lo) fp32 block (n, 16)."""
    halves = 2 if pair else 1
    nloc = n // halves
    per = 2 * (2 * nloc * 8)                   # hi block + lo block of one half
    out = torch.zeros(n, 16)
    for h in range(halves):
        w = words[h * per:(h + 1) * per].view(torch.bfloat16).float().reshape(2, 2, nloc, 8)   # (hi/lo, kgroup, row, 8)
        blk = (w[0] + w[1]).permute(1, 0, 2).reshape(nloc, 16)
        out[h * nloc:(h + 1) * nloc] = blk
    return out


class EngineModel:
    """A operands as fp32 'K group' columns (8 wide), accumulators as (rows, 512) fp32, like shared memory / TMEM."""

    def __init__(self, packed, rows):
        self.ops = packed.program_host.tolist()
        self.stream = packed.gemm.cpu()
        self.pair = packed.pair
        self.pos = 0
        self.pc = 0
        self.A = torch.zeros(rows, 256 * 8)      # K group g -> columns 8g..8g+7 (hi + lo folded together)
        self.acc = torch.zeros(rows, 512)
        self.waits = []

    def split(self, x):                           # what the epilogue writes: bf16 hi + bf16 lo
        hi = x.to(torch.bfloat16).float()
        return hi + (x - hi).to(torch.bfloat16).float()

    def write(self, kg, x):                       # x: (rows, 8k) starting at K group kg
        self.A[:, kg * 8:kg * 8 + x.shape[1]] = self.split(x)

    def run_group(self):
        """Execute ops up to and including the next COMMIT; returns the waits consumed."""
        waits = []
        while True:
            op = self.ops[self.pc]
            self.pc += 1
            kind = op & 3
            if kind == 1:
                waits.append((op >> 2) & 7)
            elif kind == 2:
                return waits, (op >> 2) & 3
            elif kind == 3:
                raise AssertionError("END inside a tile")
            else:
                n = NCODE[(op >> 2) & 3]
                accum, col = (op >> 4) & 1, ((op >> 5) & 7) * 64
                hi_kg, lo_kg, cnt = (op >> 8) & 0xFF, (op >> 16) & 0xFF, ((op >> 24) & 31) + 1
                assert lo_kg > hi_kg                      # lo copy lives behind the hi copy of the same operand
                for j in range(cnt):
                    words = self.stream[self.pos:self.pos + n * 32]
                    self.pos += n * 32
                    W = unpack_kstep(words, n, self.pair)
                    a = self.A[:, (hi_kg + 2 * j) * 8:(hi_kg + 2 * j) * 8 + 16]
                    d = a @ W.t()
                    self.acc[:, col:col + n] = d + (self.acc[:, col:col + n] if (accum or j > 0) else 0)


class EngineModelF16F8(EngineModel):
    """The f16f8 scheme (csrc/umma.cuh): 16-wide steps in 32-wide pairs; the A region holds fp16 K groups (8 wide) at
    [15:8] and per pair four FP8 K groups [r8 r8 a8 a8] (16 wide) at [23:16]; the stream holds per step and CTA half
    [fp16(S W) | e4m3(W) (even step) or e4m3(S W - fp16(S W)) (odd step) of the pair]; accumulators hold S x the product."""
    S = 4096.0

    def __init__(self, packed, rows):
        super().__init__(packed, rows)
        assert self.pair
        self.A16 = torch.zeros(rows, 256 * 8)     # fp16 K group g -> columns 8g..
        self.A8 = torch.zeros(rows, 256 * 16)     # FP8 K group g -> columns 16g..

    @staticmethod
    def q8(x):
        return x.clamp(-448, 448).to(torch.float8_e4m3fn).float()

    def write(self, kg, x):                       # kg: fp16 K group of the first column (0.. = H, 64.. = X), a multiple of 4
        assert kg % 4 == 0
        f8 = 32 + kg if kg < 32 else 72 + (kg - 64)
        a16 = x.to(torch.float16).float()
        self.A16[:, kg * 8:kg * 8 + x.shape[1]] = a16
        r8, a8 = self.q8((x - a16) * self.S), self.q8(x)
        for s in range(x.shape[1] // 32):
            base = (f8 + 4 * s) * 16
            self.A8[:, base:base + 32] = r8[:, 32 * s:32 * s + 32]
            self.A8[:, base + 32:base + 64] = a8[:, 32 * s:32 * s + 32]

    def run_group(self):
        waits = []
        while True:
            op = self.ops[self.pc]
            self.pc += 1
            kind = op & 3
            if kind == 1:
                waits.append((op >> 2) & 7)
            elif kind == 2:
                return waits, (op >> 2) & 3
            elif kind == 3:
                raise AssertionError("END inside a tile")
            else:
                n = NCODE[(op >> 2) & 3]
                nloc = n // 2
                accum, col = (op >> 4) & 1, ((op >> 5) & 7) * 64
                kg16, kg8, cnt = (op >> 8) & 0xFF, (op >> 16) & 0xFF, ((op >> 24) & 31) + 1
                assert cnt % 2 == 0                                     # steps come in 32-wide pairs
                half = (op >> 30) & 1                                   # half-width unit: the stream is grouped per step PAIR
                assert not half or (n == 128 and cnt % 4 == 0)
                blocks = {}
                if half:
                    for pr in range(cnt // 2):
                        for h in range(2):
                            for st in range(2):
                                blocks[(2 * pr + st, h)] = self.stream[self.pos:self.pos + nloc * 64]
                                self.pos += nloc * 64
                else:
                    for j in range(cnt):
                        for h in range(2):
                            blocks[(j, h)] = self.stream[self.pos:self.pos + nloc * 64]
                            self.pos += nloc * 64
                for j in range(cnt):
                    W16, F8 = torch.zeros(n, 16), torch.zeros(n, 32)
                    for h in range(2):
                        b = blocks[(j, h)]
                        rows = slice(h * nloc, (h + 1) * nloc)
                        W16[rows] = b[:nloc * 32].view(torch.float16).float().reshape(2, nloc, 8).permute(1, 0, 2).reshape(nloc, 16)
                        F8[rows] = b[nloc * 32:].view(torch.float8_e4m3fn).float().reshape(2, nloc, 16).permute(1, 0, 2).reshape(nloc, 32)
                    a16 = self.A16[:, (kg16 + 2 * j) * 8:(kg16 + 2 * j) * 8 + 16]
                    a8 = self.A8[:, (kg8 + 2 * j) * 16:(kg8 + 2 * j) * 16 + 32]     # even step: r8, odd step: a8
                    d = (a16 @ W16.t() + a8 @ F8.t()) / self.S
                    self.acc[:, col:col + n] = d + (self.acc[:, col:col + n] if (accum or j > 0) else 0)


@pytest.mark.parametrize("pair,scheme,nsplit", [(True, 'bf16x3', 'off'), (False, 'bf16x3', 'off'), (True, 'f16f8', 'off'),
                                                (True, 'f16f8', 'mixed'), (True, 'f16f8', 'full')])
def test_image_program_and_stream_reproduce_the_oracle(pair, scheme, nsplit, monkeypatch):
    """The epilogue sequence of image_umma_kernel, including the N split: after COMMIT 0 the model overwrites operand
    columns 0..127 with the layer's first output half, runs the rest of the group, and only after COMMIT 1 writes columns
    128..255 -- a program that still read the dead columns after COMMIT 0 would not reproduce the oracle."""
    monkeypatch.setenv('DDMI_B200_NSPLIT', nsplit)
    monkeypatch.setenv('DDMI_B200_IMAGE_TS', '0')                        # the shared-memory-operand kernel's program
    m = cases.build_module('image')
    sd = cases.state_dict32(m)
    coords, planes, si = cases.image_inputs(batch=1, sizes=(8, 16, 32), res=12)
    ref = orc.image_decode(sd, coords, planes, si)                       # (1,3,12,12)
    packed = packing.pack_image(m, si, _lib.PREC_F16F8 if scheme == 'f16f8' else _lib.PREC_BF16X3, pair=pair)
    EngineModel = EngineModelF16F8 if scheme == 'f16f8' else globals()['EngineModel']
    vec = packed.vec
    grid = coords.permute(0, 2, 3, 1)
    X = [torch.nn.functional.grid_sample(p, grid, padding_mode='border', align_corners=False).permute(0, 2, 3, 1).reshape(-1, 64)
         for p in planes]
    E = EngineModel(packed, rows=144)
    lr = lambda v: torch.nn.functional.leaky_relu(v, 0.2)
    E.write(64, X[0])                                                    # X region: K groups 64.. (hi), 72.. (lo)

    def stage(out_of_acc, after_first=None):
        """One GEMM group + its epilogue stage; out_of_acc(lo, hi) -> the stage's output columns [lo, hi)."""
        waits, done = E.run_group()
        assert done == 0
        E.write(0, out_of_acc(0, 128))
        waits2, done = E.run_group()                                     # the second accumulator half
        assert done == 1 and sorted(waits + waits2) == [0, 1, 2, 3, 4]
        E.write(16, out_of_acc(128, 256))

    for blk in range(4):
        bv = vec[blk * 1024:(blk + 1) * 1024]
        stage(lambda lo, hi: lr(E.acc[:, lo:hi] + bv[lo:hi]))           # conv1 (+ skip)
        if blk < 2:
            E.write(64, X[blk + 1])
        stage(lambda lo, hi: lr(E.acc[:, lo:hi] + bv[256 + lo:256 + hi]))   # conv2
        stash = {}

        def conv3(lo, hi, blk=blk, bv=bv, stash=stash):
            h = lr(E.acc[:, lo:hi] + bv[512 + lo:512 + hi]) + E.acc[:, 256 + lo:256 + hi] + (bv[768 + lo:768 + hi] if blk < 3 else 0)
            if blk == 2:
                E.acc[:, 256 + lo:256 + hi] = h / math.sqrt(2.0)          # stash res4's identity skip
            return h
        stage(conv3)
    waits, done = E.run_group()                                          # ToRGB
    waits2, done2 = E.run_group()
    assert (done, done2) == (0, 1) and sorted(waits + waits2) == [0, 1, 2, 3, 4]
    out = (E.acc[:, :3] + vec[4096 + 768:4096 + 771]).t().reshape(1, 3, 12, 12)
    assert E.ops[E.pc] & 3 == 3 and E.pos == E.stream.numel()            # program and stream end together
    assert float((out - ref).abs().max()) < 1e-3                         # (the model keeps acc / S; the kernel keeps acc)
    if nsplit != 'off':                                                  # the split really is in the program
        units_between = 0
        ops = E.ops
        for i, o in enumerate(ops):
            if o & 3 == 2 and (o >> 2) & 3 == 0:
                j = i + 1
                while ops[j] & 3 != 2:
                    units_between += ops[j] & 3 == 0
                    j += 1
        assert units_between >= 11


class EngineModelTS(EngineModelF16F8):
    """A-from-TMEM units (op bit 29): K groups count 4 TMEM columns from the TMEM base; the image kernel keeps H there as
    fp16 at columns 256.. (two values per column) and FP8 at 384.. (per 32 K columns [r8: 8 columns | a8: 8 columns])."""

    def __init__(self, packed, rows):
        super().__init__(packed, rows)
        self.T16 = torch.zeros(rows, 256)
        self.T8 = torch.zeros(rows, 2, 256)       # [:, 0] = r8, [:, 1] = a8

    def write_h(self, k0, x):
        a16 = x.to(torch.float16).float()
        self.T16[:, k0:k0 + x.shape[1]] = a16
        self.T8[:, 0, k0:k0 + x.shape[1]] = self.q8((x - a16) * self.S)
        self.T8[:, 1, k0:k0 + x.shape[1]] = self.q8(x)

    def operands(self, op, j):
        kg16, kg8 = (op >> 8) & 0xFF, (op >> 16) & 0xFF
        if not (op >> 29) & 1:
            return (self.A16[:, (kg16 + 2 * j) * 8:(kg16 + 2 * j) * 8 + 16],
                    self.A8[:, (kg8 + 2 * j) * 16:(kg8 + 2 * j) * 16 + 32])
        c16, c8 = kg16 * 4 + 8 * j, kg8 * 4 + 8 * j                     # TMEM columns of this 16-wide step
        assert 256 <= c16 and c16 + 8 <= 384 and 384 <= c8 and c8 + 8 <= 512
        k16 = (c16 - 256) * 2
        pair, which = divmod(c8 - 384, 16)
        assert which in (0, 8) and which == 8 * (j & 1) and 32 * pair == k16 - 16 * (j & 1)
        return self.T16[:, k16:k16 + 16], self.T8[:, which // 8, 32 * pair:32 * pair + 32]

    def run_group(self):
        waits = []
        while True:
            op = self.ops[self.pc]
            self.pc += 1
            kind = op & 3
            if kind == 1:
                waits.append((op >> 2) & 7)
            elif kind == 2:
                return waits, (op >> 2) & 3
            elif kind == 3:
                raise AssertionError("END inside a tile")
            else:
                n = NCODE[(op >> 2) & 3]
                nloc = n // 2
                accum, col, cnt = (op >> 4) & 1, ((op >> 5) & 7) * 64, ((op >> 24) & 31) + 1
                half = (op >> 30) & 1
                assert cnt % (4 if half else 2) == 0 and (not half or n == 128)
                blocks = {}
                for j0 in range(0, cnt, 2 if half else 1):               # stream order: per step (pair, if half-width), per CTA
                    for h in range(2):
                        for st in range(2 if half else 1):
                            blocks[(j0 + st, h)] = self.stream[self.pos:self.pos + nloc * 64]
                            self.pos += nloc * 64
                for j in range(cnt):
                    W16, F8 = torch.zeros(n, 16), torch.zeros(n, 32)
                    for h in range(2):
                        b = blocks[(j, h)]
                        rows = slice(h * nloc, (h + 1) * nloc)
                        W16[rows] = b[:nloc * 32].view(torch.float16).float().reshape(2, nloc, 8).permute(1, 0, 2).reshape(nloc, 16)
                        F8[rows] = b[nloc * 32:].view(torch.float8_e4m3fn).float().reshape(2, nloc, 16).permute(1, 0, 2).reshape(nloc, 32)
                    a16, a8 = self.operands(op, j)
                    d = (a16 @ W16.t() + a8 @ F8.t()) / self.S
                    self.acc[:, col:col + n] = d + (self.acc[:, col:col + n] if (accum or j > 0) else 0)


def test_image_ts_program_and_stream_reproduce_the_oracle(monkeypatch):
    """The epilogue sequence of image_umma_kernel<.., TS = 1> (H in tensor memory, one accumulator, parked skip, ToRGB in fp32):
    after COMMIT 0 the model overwrites H columns 0..127 in place, after COMMIT 1 columns 128..255."""
    monkeypatch.setenv('DDMI_B200_IMAGE_TS', '1')
    m = cases.build_module('image')
    sd = cases.state_dict32(m)
    coords, planes, si = cases.image_inputs(batch=1, sizes=(8, 16, 32), res=12)
    ref = orc.image_decode(sd, coords, planes, si)
    packed = packing.pack_image(m, si, _lib.PREC_F16F8, pair=True)
    assert packed.ts and packed.vec_host is not None and len(packed.program_host) <= 256
    # csrc/image_ts_issuer.cuh is this op table written out as code: the launcher refuses any other (kImageTsProgramOps / Hash)
    ops = [o & 0xFFFFFFFF for o in packed.program_host.tolist()][:-4]
    h = 0x811C9DC5
    for o in ops:
        for b in range(4):
            h = ((h ^ ((o >> (8 * b)) & 0xFF)) * 0x01000193) & 0xFFFFFFFF
    src = open(os.path.join(os.path.dirname(__file__), '..', 'ddmi_b200', 'csrc', 'image_ts_issuer.cuh')).read()
    assert f'kImageTsProgramOps = {len(ops)};' in src and f'kImageTsProgramHash = 0x{h:08x}u;' in src
    vec = packed.vec.cpu()
    g = coords.permute(0, 2, 3, 1)
    X = [torch.nn.functional.grid_sample(p, g, mode='bilinear', padding_mode='border', align_corners=False)[0]
         .reshape(64, -1).t().contiguous() for p in planes]
    E = EngineModelTS(packed, rows=144)
    lr = lambda v: torch.nn.functional.leaky_relu(v, 0.2)
    E.write(64, X[0])
    parked = torch.zeros(144, 256)

    def stage(out_of_acc, published, publishes=True):
        waits, done = E.run_group()
        assert done == 0
        lo = out_of_acc(0, 128)
        if publishes:
            E.write_h(0, lo)
        waits2, done = E.run_group()
        assert done == 1 and sorted(waits + waits2) == ([0, 1, 2, 3, 4, 5] if published else [4, 5])
        hi = out_of_acc(128, 256)
        if publishes:
            E.write_h(128, hi)
        return torch.cat([lo, hi], dim=1)

    for blk in range(4):
        bv = vec[blk * 1024:(blk + 1) * 1024]
        if blk < 3:
            parked = stage(lambda lo, hi: E.acc[:, lo:hi].clone(), True, publishes=False)      # skip GEMM -> parked
        stage(lambda lo, hi: lr(E.acc[:, lo:hi] + bv[lo:hi]), blk == 3)                         # conv1
        if blk < 2:
            E.write(64, X[blk + 1])
        stage(lambda lo, hi: lr(E.acc[:, lo:hi] + bv[256 + lo:256 + hi]), True)                 # conv2
        h = stage(lambda lo, hi: lr(E.acc[:, lo:hi] + bv[512 + lo:512 + hi]) + parked[:, lo:hi]
                  + (bv[768 + lo:768 + hi] if blk < 3 else 0), True, publishes=blk < 3)         # conv3 + skip
        if blk == 2:
            parked = h / math.sqrt(2.0)
    assert E.ops[E.pc] & 3 == 3 and E.pos == E.stream.numel()
    out = (h @ vec[4096:4096 + 768].reshape(3, 256).t() + vec[4096 + 768:4096 + 771]).t().reshape(1, 3, 12, 12)
    assert float((out - ref).abs().max()) < 1e-3


def test_programs_consume_exactly_their_streams():
    for packed in (packing.pack_occupancy(cases.build_module('occupancy'), _lib.PREC_BF16X3),
                   packing.pack_video(cases.build_module('video'), _lib.PREC_BF16X3),
                   packing.pack_nerf(cases.build_module('nerf'), _lib.PREC_BF16X3),
                   packing.pack_occupancy(cases.build_module('occupancy'), _lib.PREC_F16F8),
                   packing.pack_video(cases.build_module('video'), _lib.PREC_F16F8),
                   packing.pack_nerf(cases.build_module('nerf'), _lib.PREC_F16F8)):
        ops = packed.program_host.tolist()
        assert ops[-4:] == [3, 3, 3, 3] and all(o & 3 != 3 for o in ops[:-4])
        need = sum(NCODE[(o >> 2) & 3] * 64 * (((o >> 24) & 31) + 1) for o in ops if o & 3 == 0)
        assert need == packed.gemm.numel() * packed.gemm.element_size()
        if packed.precision == _lib.PREC_F16F8:     # 16-wide steps come in 32-wide pairs
            assert all((((o >> 24) & 31) + 1) % 2 == 0 for o in ops if o & 3 == 0)
        # every accumulator region is started with accumulate = 0 before it is accumulated into
        assert any(o & 3 == 0 and not (o >> 4) & 1 for o in ops)


def test_operand_barriers_never_run_two_phases_ahead():
    """Video program: between two WAITs on the same operand barrier there is a COMMIT (the E threads only
    re-signal a barrier after consuming a commit that follows its previous WAIT) -- the parity-wait safety rule."""
    ops = packing.pack_video(cases.build_module('video'), _lib.PREC_BF16X3).program_host.tolist()[:-4]
    ops = ops + ops                                  # two consecutive tiles
    last_wait = {}
    for i, o in enumerate(ops):
        if o & 3 == 1:
            b = (o >> 2) & 7
            if b in last_wait:
                assert any(x & 3 == 2 for x in ops[last_wait[b]:i]), (b, i)
            last_wait[b] = i
