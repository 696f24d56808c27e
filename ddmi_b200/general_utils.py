"""Caller-side coordinate helpers of the D2C-VAE decode path.

These mirror the *interface* (names, argument meaning, returned layouts) of the
reference's ``utils/general_utils.py`` so the call sites listed in SURVEY.md §3
work unchanged.  They are plain tensor plumbing; the sampling functions of that
file (``singleplane_positional_encoding``, ``triplane_positional_encoding``,
``sample_plane_feature``/``normalize_coordinate``) are NOT here -- they are fused
into the CUDA kernels (``ddmi_b200/csrc``).
"""
import torch


def _axis(start, end, n, device):
    return torch.linspace(start, end, n, device=device)


def convert_to_coord_format_2d(b, h, w, device='cpu', integer_values=False,
                               hstart=-1, hend=1, wstart=-1, wend=1):
    """Regular 2-D query grid ``(b, 2, h, w)``; channel 0 = x (varies along the
    last dim), channel 1 = y.  Reference: utils/general_utils.py:27-35.  Like the
    reference this is only consistent for square grids (it tiles x ``w`` times
    down and y ``h`` times across); we keep that quirk rather than "fix" it."""
    if integer_values:
        xs = torch.arange(w, dtype=torch.float, device=device)
        ys = torch.arange(h, dtype=torch.float, device=device)
    else:
        xs = _axis(wstart, wend, w, device)
        ys = _axis(hstart, hend, h, device)
    x_channel = xs.view(1, 1, 1, w).repeat(b, 1, w, 1)
    y_channel = ys.view(1, 1, h, 1).repeat(b, 1, 1, h)
    return torch.cat((x_channel, y_channel), dim=1)


def convert_to_coord_format_3d(b, h, w, t, device='cpu', hstart=-1, hend=1,
                               wstart=-1, wend=1, tstart=-1, tend=1):
    """Dict of the three 2-D grids a video decode queries.
    Reference: utils/general_utils.py:38-52.  NOTE the channel order of the
    temporal planes is (t, x) / (t, y): grid_sample uses channel 0 as the
    *width* index, so the ``(T, W)`` planes are indexed with transposed axes
    (SURVEY.md F6).  The kernels take these tensors literally."""
    xs = _axis(wstart, wend, w, device)
    ys = _axis(hstart, hend, h, device)
    ts = _axis(tstart, tend, t, device)
    out = {}
    out['xy'] = torch.cat((xs.view(1, 1, 1, w).repeat(b, 1, h, 1),
                           ys.view(1, 1, h, 1).repeat(b, 1, 1, w)), dim=1)
    out['xt'] = torch.cat((ts.view(1, 1, t, 1).repeat(b, 1, 1, w),
                           xs.view(1, 1, 1, w).repeat(b, 1, t, 1)), dim=1)
    out['yt'] = torch.cat((ts.view(1, 1, t, 1).repeat(b, 1, 1, h),
                           ys.view(1, 1, 1, h).repeat(b, 1, t, 1)), dim=1)
    return out


def get_scale_injection(current_res, anchor_res=256):
    """Scale-injection scalar ``si``.  Reference: utils/general_utils.py:204-206."""
    return anchor_res / current_res


def make_3d_grid(bb_min, bb_max, shape):
    """Dense query lattice ``(prod(shape), 3)``, z fastest.
    Semantics of convocc/src/common.py:145-164 (used by
    Generator3D.generate_mesh_fromdiffusion, generation.py:90-97)."""
    axes = [torch.linspace(bb_min[i], bb_max[i], shape[i]) for i in range(3)]
    gx, gy, gz = torch.meshgrid(*axes, indexing='ij')
    return torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)
