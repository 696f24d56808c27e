"""CPU: the oracle restatement vs the outputs of the REFERENCE itself
(tests/golden/*.pt, produced by oracle/make_golden.py in the build container).
This is the pin of SURVEY.md §8c: the reference has no tests of its own."""
import os

import pytest
import torch

from oracle import cases, ddmi_oracle as orc

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'))


def _flat(x):
    return [t for a in x for t in (a if isinstance(a, (list, tuple)) else [a])]


def _check_inputs(g, inputs):
    cs = cases.checksum(inputs)
    assert abs(cs - g['input_checksum']) <= 1e-9 * abs(cs), (
        "seeded inputs differ from the ones the golden outputs were generated from "
        "(torch RNG drift?) -- regenerate with oracle/make_golden.py")


@pytest.mark.parametrize("tag,kw", [("image_96", dict(batch=2, sizes=(16, 32, 64), res=96)),
                                    ("image_native", dict(batch=1, sizes=(8, 16, 32), res=32))])
def test_image(golden_dir, tag, kw):
    g = _load(golden_dir, tag)
    sd = cases.state_dict32(cases.build_module('image'))
    coords, planes, si = cases.image_inputs(**kw)
    _check_inputs(g, planes + list(sd.values()))
    out = orc.image_decode(sd, coords, planes, si)
    assert out.shape == g['out'].shape
    assert float((out - g['out']).abs().max()) < 2e-5


def test_image_with_noise_injection(golden_dir):
    """Reference run with explicit noise tensors handed to its 12 NoiseInjection modules (non-zero weights)."""
    g = _load(golden_dir, 'image_noise')
    sd = cases.state_dict32(cases.build_module('image_noise'))
    coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
    noise = cases.image_noise_tensors(2, 96)
    _check_inputs(g, planes + noise + list(sd.values()))
    out = orc.image_decode(sd, coords, planes, si, noise=noise)
    assert float((out - g['out']).abs().max()) < 2e-5
    with pytest.raises(ValueError):
        orc.image_decode(sd, coords, planes, si)


def test_occupancy(golden_dir):
    g = _load(golden_dir, 'occupancy')
    sd = cases.state_dict32(cases.build_module('occupancy'))
    pts, hdbf = cases.occupancy_inputs()
    _check_inputs(g, _flat(hdbf) + [pts] + list(sd.values()))
    out = orc.occupancy_logits(sd, pts, hdbf)
    assert float((out - g['out']).abs().max()) < 1e-5
    assert bool(((out > 0) == (g['out'] > 0)).all())


def test_video(golden_dir):
    g = _load(golden_dir, 'video')
    sd = cases.state_dict32(cases.build_module('video'))
    coords, hdbf = cases.video_inputs()
    _check_inputs(g, _flat(hdbf) + list(sd.values()))
    out = orc.video_decode(sd, coords, hdbf)
    assert out.shape == g['out'].shape == (1, 3, 4, 32, 32)
    assert float((out - g['out']).abs().max()) < 1e-5


def test_nerf_mlp(golden_dir):
    g = _load(golden_dir, 'nerf_mlp')
    sd = cases.state_dict32(cases.build_module('nerf'))
    x = cases.nerf_mlp_inputs()
    _check_inputs(g, [x] + list(sd.values()))
    assert float((orc.nerf_mlp(sd, x) - g['out']).abs().max()) < 1e-5
    # LeakyReLU(True) is the identity: the xyz trunk is affine
    s1 = orc.nerf_mlp(sd, x[:, :159], sigma_only=True)
    s2 = orc.nerf_mlp(sd, 2 * x[:, :159], sigma_only=True)
    s0 = orc.nerf_mlp(sd, 0 * x[:, :159], sigma_only=True)
    assert float(((s2 - s0) - 2 * (s1 - s0)).abs().max()) < 1e-3


def test_nerf_render(golden_dir):
    g = _load(golden_dir, 'nerf_render')
    sd = cases.state_dict32(cases.build_module('nerf'))
    res, K, fea, c2w = cases.nerf_inputs()
    _check_inputs(g, list(fea.values()) + list(sd.values()))
    out = orc.nerf_render_rays(sd, g['rays'], fea, 64, True)
    assert float((out - g['out']).abs().max()) < 1e-5
    assert float(g['out'].max() - g['out'].min()) > 0.3   # the fixture is not degenerate


@pytest.mark.parametrize("tag,perturb,lindisp", [("nerf_render_perturb", 1.0, False), ("nerf_render_lindisp", 0., True),
                                                 ("nerf_render_perturb_lindisp", 1.0, True)])
def test_nerf_render_stratified_and_lindisp(golden_dir, tag, perturb, lindisp):
    """perturb > 0 draws torch.rand(z_vals.shape) on the CPU generator (utils/nerf_helpers.py:370): seeded like the
    generator script, the oracle reproduces the reference's samples."""
    g = _load(golden_dir, tag)
    sd = cases.state_dict32(cases.build_module('nerf'))
    res, K, fea, c2w = cases.nerf_inputs()
    _check_inputs(g, list(fea.values()) + list(sd.values()))
    torch.manual_seed(cases.NERF_PERTURB_SEED)
    out = orc.nerf_render_rays(sd, g['rays'], fea, 64, True, perturb=perturb, lindisp=lindisp)
    assert float((out - g['out']).abs().max()) < 1e-5
    plain = _load(golden_dir, 'nerf_render')['out']
    assert float((g['out'] - plain).abs().max()) > 1e-3      # the option changes the image


def test_oracle_fp64_agrees():
    """fp64 evaluation of the oracle bounds the fp32 reference's own rounding noise."""
    sd = cases.state_dict32(cases.build_module('occupancy'))
    pts, hdbf = cases.occupancy_inputs(n=2000)
    o32 = orc.occupancy_logits(sd, pts, hdbf)
    sd64 = {k: v.double() for k, v in sd.items()}
    o64 = orc.occupancy_logits(sd64, pts.double(), tuple([t.double() for t in a] for a in hdbf))
    assert float((o32.double() - o64).abs().max()) < 5e-5


def test_sample_pdf(golden_dir):
    """Hierarchical sampling (utils/nerf_helpers.py:166-209): deterministic u, and the reference's own pytest-mode u."""
    import numpy as np
    g = _load(golden_dir, 'sample_pdf')
    _check_inputs(g, [g['bins'], g['weights']])
    out = orc.sample_pdf(g['bins'], g['weights'], torch.linspace(0., 1., 128).expand(300, 128))
    assert float((out - g['out']).abs().max()) < 1e-6
    np.random.seed(0)
    u = torch.Tensor(np.random.rand(300, 96))
    assert float((orc.sample_pdf(g['bins'], g['weights'], u) - g['out_pytest']).abs().max()) < 1e-6


def test_marching_cubes_oracle_is_the_reference(golden_dir):
    """Occupancy post-step: the Python restatement of libmcubes against the REFERENCE's outputs (oracle/_ref, committed as
    tests/golden/mcubes.pt) -- bit for bit, vertex and triangle order included; and against oracle/_ref itself when built."""
    import numpy as np
    from oracle import mcubes_oracle as mo
    g = torch.load(os.path.join(golden_dir, 'mcubes.pt'))
    for i, name in enumerate(('a', 'b', 'c')):
        c = g[name]
        vol = cases.mcubes_volume(c['shape'], seed=i)
        assert float(vol.double().sum()) == c['vol_sum']
        v, t = mo.marching_cubes(vol.numpy(), c['iso'])
        assert np.array_equal(v, c['vertices'].numpy().reshape(-1, 3)) and np.array_equal(t, c['triangles'].numpy().reshape(-1, 3))
        ref = mo.ref_marching_cubes(vol.numpy(), c['iso'])
        if ref is not None:
            assert np.array_equal(ref[0], v) and np.array_equal(ref[1], t)
    c = g['mesh']
    vol = cases.mcubes_volume(c['shape'], seed=9, noise=0.5, scale=1.5)
    v, t = mo.extract_mesh(vol.numpy(), 0.2, 0.1)
    assert np.array_equal(v, c['vertices'].numpy()) and np.array_equal(t, c['triangles'].numpy())
    # a closed surface: the -1e6 padding makes every edge of the mesh shared by exactly two triangles
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()


def test_plane_tail_oracle_is_the_reference(golden_dir):
    """Plane-producer tail: the oracle's restatement against inputs / outputs captured inside the reference Decoder's forward."""
    from oracle import plane_tail_oracle as po
    g = torch.load(os.path.join(golden_dir, 'plane_tail.pt'))
    for tag in ('plain', 'tanh'):
        c = g[tag]
        assert float((po.tail(c['sd'], c['tail_in'], 32, c['tanh_out']) - c['tail_out']).abs().max()) < 1e-6
        for k, (hin, hout) in c['heads'].items():
            assert float((po.head(c['sd'], int(k[4:]), hin) - hout).abs().max()) < 1e-6
