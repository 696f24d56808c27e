"""CPU: host logic -- C-ABI library exports, state-dict layout, weight folding,
coordinate helpers, argument validation that needs no GPU."""
import ctypes
import math
import os
import re

import pytest
import torch

import ddmi_b200
from ddmi_b200 import _lib, packing
from oracle import cases, ddmi_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.set_grad_enabled(False)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'ddmi_b200.h')).read()
    declared = set(re.findall(r'DDMI_API\s+(?:const\s+char\*|int64_t|int)\s+(ddmi_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ddmi_abi_version() == _lib.ABI_VERSION
    assert L.ddmi_status_string(1).decode() == 'bad argument'


def test_abi_rejects_bad_arguments_without_gpu():
    L = _lib.lib()
    w = _lib.Weights()
    planes = (_lib.Plane * 3)()
    rc = L.ddmi_decode_image(planes, 1, 64, None, None, 16, ctypes.byref(w), None, None)
    assert rc == 1 and b'planes[0].data is NULL' in L.ddmi_last_error()
    rc = L.ddmi_selftest_umma(None, None, None, 256, 256, None)
    assert rc == 1
    # ABI 10: workspace sizes are pure host arithmetic; the workspace entry points validate before any launch
    assert L.ddmi_video_workspace_bytes(2, 16, 256, 256, _lib.PREC_F16F8) == 2 * 3 * (256 * 256 + 2 * 16 * 256) * 256
    assert L.ddmi_video_workspace_bytes(2, 16, 256, 256, _lib.PREC_BF16X3) == 0      # the table path is f16f8's
    assert L.ddmi_occupancy_lattice_workspace_bytes(3, 128, 128, 128) == 3 * 3 * (3 * 128 * 128) * 256
    assert L.ddmi_occupancy_lattice_workspace_bytes(1, 2048, 2048, 2048) == 0           # more than 2^31 - 1 points
    p9 = (_lib.Plane * 9)()
    rc = L.ddmi_decode_video_ws(p9, 1, 64, None, None, None, 1, 1, 1, ctypes.byref(w), 0, None, None, 0, None)
    assert rc == 1 and b'planes[0].data is NULL' in L.ddmi_last_error()
    rc = L.ddmi_decode_occupancy_lattice(p9, 1, 64, 1, None, 4, 4, 4, 0.1, ctypes.byref(w), None, None, 0, None)
    assert rc == 1 and b'planes[0].data is NULL' in L.ddmi_last_error()


EXPECTED_KEYS = {
    'image': ['time_mlp.1.weight', 'net_res1.conv1.conv.weight', 'net_res1.conv1.conv.modulation.bias',
              'net_res1.conv1.noise.weight', 'net_res1.conv1.activate.bias', 'net_res1.skip.0.weight',
              'net_res4.conv3.conv.modulation.weight', 'torgb.bias', 'torgb.conv.modulation.weight'],
    'occupancy': ['net_p.weight', 'net_res1.fc_0.weight', 'net_res1.shortcut.weight', 'net_res4.fc_1.bias', 'net_out.bias'],
    'video': ['net_res1.fc_0.weight', 'net_res2.shortcut.weight', 'net_out.weight'],
    'nerf': ['xyz_encoding_1.0.weight', 'xyz_encoding_3.0.weight', 'xyz_encoding_final.bias',
             'dir_encoding.0.weight', 'sigma.weight', 'rgb.0.bias'],
}
PARAM_COUNT = {'image': 1880021, 'occupancy': 629825, 'video': 858819, 'nerf': 554116}


@pytest.mark.parametrize("kind", list(EXPECTED_KEYS))
def test_state_dict_matches_reference_inventory(kind):
    m = cases.build_module(kind)
    sd = m.state_dict()
    for k in EXPECTED_KEYS[kind]:
        assert k in sd, k
    assert sum(v.numel() for v in sd.values()) == PARAM_COUNT[kind]
    if kind == 'image':
        assert 'net_res4.skip.0.weight' not in sd
        assert tuple(sd['net_res2.conv1.conv.weight'].shape) == (1, 256, 322, 1, 1)
        assert tuple(sd['net_res2.skip.0.weight'].shape) == (256, 322, 1, 1)
    if kind == 'nerf':
        assert tuple(sd['xyz_encoding_3.0.weight'].shape) == (256, 415)
        assert m.negative_slope == 1.0   # nn.LeakyReLU(True), SURVEY.md F3


def test_fold_image_reproduces_reference_layers():
    """Folded weights applied as plain matmuls == the oracle's modulated convs."""
    m = cases.build_module('image')
    sd = cases.state_dict32(m)
    si = 256 / 96
    f = packing.fold_image(m, si)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(5, 64, generator=g, dtype=torch.float64)
    xin = torch.cat([x0, torch.full((5, 2), si, dtype=torch.float64)], dim=1)
    # oracle block on a (5,66,1,1) "image"
    style = orc._sinusoidal(torch.ones(5) * si, 64)
    style = torch.nn.functional.linear(style, sd['time_mlp.1.weight'], sd['time_mlp.1.bias'])
    style = torch.nn.functional.linear(torch.nn.functional.gelu(style), sd['time_mlp.3.weight'], sd['time_mlp.3.bias'])
    ref = orc._styled_res_block(sd, 'net_res1', xin.float().view(5, 66, 1, 1), style).view(5, 256).double()
    d = f['blocks'][0]
    lr = lambda v: torch.nn.functional.leaky_relu(v, 0.2)
    h = math.sqrt(2) * lr(x0 @ d['W1'].t() + d['b1'])
    h = math.sqrt(2) * lr(h @ d['W2'].t() + d['b2'])
    h = lr(h @ d['W3'].t() + d['b3']) + x0 @ d['Ws'].t() + d['cs']
    assert float((h - ref).abs().max()) < 2e-5


def test_packed_sizes_match_abi_constants():
    for kind, fn, gfl, vfl in (
            ('occupancy', packing.pack_occupancy, (64 * 64 + 64 * 256 * 2) + 2 * (896 * 256) + 2 * 65536, 2881),
            ('video', packing.pack_video, (192 * 192 + 192 * 256 * 2) + 2 * (1152 * 256) + 2 * 65536, 2755),
            ('nerf', packing.pack_nerf, (160 + 256 + 416 + 256 + 416 + 256 + 256) * 256 + 288 * 128, 2564)):
        p = fn(cases.build_module(kind))
        assert p.gemm.numel() == gfl and p.vec.numel() == vfl, kind
    p = packing.pack_image(cases.build_module('image'), 1.0, _lib.PREC_FP32)
    assert p.gemm.numel() == (640 + 2 * 1152 + 768) * 256 and p.vec.numel() == 4879


def test_split_bf16_is_accurate():
    x = torch.randn(4096) * 3
    hi, lo = packing._split_bf16(x)
    err = (hi.float() + lo.float() - x).abs() / x.abs().clamp_min(1e-20)
    assert float(err.max()) < 2 ** -15


def test_umma_kstep_layout():
    W = torch.arange(32 * 32, dtype=torch.float32).reshape(32, 32) / 64.0   # exactly representable in bf16? not all; use hi only
    blob = packing.umma_kstep_blocks(W, 0, 32)
    # per step: hi block (2*32*8) then lo block
    assert blob.numel() == 2 * 2 * (2 * 32 * 8)
    step = blob.reshape(2, 2, 2, 32, 8)   # (kstep, hi/lo, kgroup, n, 8)
    hi = step[:, 0].view(torch.bfloat16).float()
    n, k = 5, 27
    assert float(hi[k // 16, (k % 16) // 8, n, k % 8]) == float(W[n, k].to(torch.bfloat16))


def test_noise_weights_are_packed_with_the_folded_gain():
    """NoiseInjection.weight != 0 (trained checkpoints): the 12 weights ride at the end of the vec blob, conv1 / conv2 with the
    sqrt(2) activation gain the tcgen05 packing folds into weights and biases."""
    from ddmi_b200 import _lib
    m = cases.build_module('image')
    assert not packing.pack_image(m, 1.0, _lib.PREC_FP32).noise_active
    m.net_res2.conv1.noise.weight.data.fill_(0.1)
    m.net_res2.conv3.noise.weight.data.fill_(0.3)
    p = packing.pack_image(m, 1.0, _lib.PREC_FP32)
    assert p.noise_active and p.vec[-12:].tolist() == pytest.approx([0, 0, 0, 0.1, 0, 0.3, 0, 0, 0, 0, 0, 0])
    q = packing.pack_image(m, 1.0, _lib.PREC_F16F8)
    assert q.vec[-12:].tolist() == pytest.approx([0, 0, 0, 0.1 * 2 ** 0.5, 0, 0.3, 0, 0, 0, 0, 0, 0])


def test_philox_noise_stream_is_standard_normal_and_keyed():
    """The documented noise stream (oracle restatement of csrc/common.cuh::philox_normal3): known-answer for Philox4x32-10,
    N(0,1) moments, distinct per layer / item / seed."""
    import numpy as np
    from oracle import ddmi_oracle as orc
    # Random123 known-answer test: counter = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    out = orc.philox4x32_10(np.zeros((1, 4), dtype=np.uint32), (0, 0))[0]
    assert [hex(int(v)) for v in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    z = orc.philox_noise(1234, 2, 50000)
    assert len(z) == 12 and z[0].shape == (2, 1, 50000)
    allz = torch.stack(z)
    assert abs(float(allz.mean())) < 5e-3 and abs(float(allz.std()) - 1.0) < 5e-3
    assert float((z[0] - z[1]).abs().max()) > 1 and float((z[0][0] - z[0][1]).abs().max()) > 1
    assert float((orc.philox_noise(1235, 2, 64)[0] - z[0][:, :, :64]).abs().max()) > 0.5


def test_cpu_tensors_fail_loudly():
    m = cases.build_module('occupancy')
    pts, hdbf = cases.occupancy_inputs(n=16)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(pts, hdbf)


def test_grad_mode_is_rejected():
    m = cases.build_module('nerf')
    with torch.enable_grad():
        with pytest.raises(RuntimeError, match="inference only"):
            m(torch.zeros(4, 186))


def test_coord_helpers():
    c = ddmi_b200.convert_to_coord_format_2d(1, 4, 4, hstart=-.75, hend=.75, wstart=-.75, wend=.75)
    assert tuple(c.shape) == (1, 2, 4, 4)
    assert torch.equal(c[0, 0, 0], torch.linspace(-.75, .75, 4)) and torch.equal(c[0, 1, :, 0], torch.linspace(-.75, .75, 4))
    d = ddmi_b200.convert_to_coord_format_3d(1, 4, 6, 3)
    assert tuple(d['xy'].shape) == (1, 2, 4, 6) and tuple(d['xt'].shape) == (1, 2, 3, 6) and tuple(d['yt'].shape) == (1, 2, 3, 4)
    assert torch.equal(d['xt'][0, 0, :, 0], torch.linspace(-1, 1, 3))      # channel 0 of 'xt' is t (F6)
    assert ddmi_b200.get_scale_injection(1024) == 0.25
    g = ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (2, 3, 4))
    assert tuple(g.shape) == (24, 3) and float(g[1, 2] - g[0, 2]) > 0 and float(g[1, 0] - g[0, 0]) == 0


def test_f16f8_default_falls_back_to_bf16x3_on_out_of_range_weights():
    """f16f8 packs fp16(4096 W): a weight >= 16 cannot be represented.  As the DEFAULT precision that silently would be
    wrong, so the module falls back to bf16x3 (with a warning); an explicit f16f8 request raises."""
    import warnings
    import pytest
    from ddmi_b200 import _lib, packing
    from oracle import cases
    m = cases.build_module('occupancy')
    with torch.no_grad():
        m.net_res2.fc_0.weight[3, 5] = 17.0
    key_of = lambda pr: ('occ', pr, True)
    build_of = lambda pr: packing.pack_occupancy(m, pr, True)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        prec, packed = m._packed_auto(_lib.PREC_F16F8, key_of, build_of)
    assert prec == _lib.PREC_BF16X3 and packed.precision == _lib.PREC_BF16X3 and len(w) == 1
    m.precision = 'f16f8'
    with pytest.raises(packing.F16F8RangeError):
        m._packed_auto(_lib.PREC_F16F8, key_of, build_of)


def test_packed_cache_sees_data_writes_that_keep_the_version_counter():
    """`p.data.copy_(...)` -- the idiom of the reference's EMA swap (models/ema.py LitEma.copy_to / restore) -- changes a
    parameter without bumping its version counter; the cache key carries a content digest, so the packed weights are rebuilt."""
    from ddmi_b200 import _lib, packing
    m = cases.build_module('occupancy')
    built = []
    build = lambda: built.append(1) or packing.pack_occupancy(m, _lib.PREC_FP32)
    a = m._packed(('occ', 0), build)
    assert m._packed(('occ', 0), build) is a and len(built) == 1
    v0 = m.net_out.bias._version
    m.net_out.bias.data.copy_(m.net_out.bias.data + 1.0)
    assert m.net_out.bias._version == v0                      # the write is invisible to the version counter ...
    b = m._packed(('occ', 0), build)
    assert b is not a and len(built) == 2                     # ... but not to the digest
    assert float(b.vec[-1] - a.vec[-1]) == 1.0
    m.invalidate_packed()
    assert m._packed(('occ', 0), build) is not b and len(built) == 3


def test_mismatched_planes_are_rejected_before_any_launch():
    """The C ABI takes ONE (batch, channels) pair for all planes of a decode; a smaller plane must raise, not be over-read."""
    from ddmi_b200.mlp import _check_plane_set
    ok = [torch.zeros(2, 64, 4, 4), torch.zeros(2, 64, 8, 8)]
    _check_plane_set(ok, 64, ['a', 'b'])
    with pytest.raises(RuntimeError, match="batch"):
        _check_plane_set([ok[0], torch.zeros(1, 64, 8, 8)], 64, ['a', 'b'])
    with pytest.raises(RuntimeError, match="channels"):
        _check_plane_set([ok[0], torch.zeros(2, 32, 8, 8)], 64, ['a', 'b'])
