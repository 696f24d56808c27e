"""Occupancy post-step on the GPU: the numeric part of the reference's mesh generator
(convocc/src/conv_onet/generation.py `Generator3D`) behind the same names and arguments --
`eval_points` (chunked occupancy queries, :123-144), `marching_cubes` (libmcubes.marching_cubes,
convocc/src/utils/libmcubes/mcubes.pyx:24-29), `extract_mesh` (:146-186) and the dense-grid path of
`generate_mesh_fromdiffusion` (:84-98) -- with everything kept in device memory: the decoded logit grid never visits the host
and marching cubes runs in the library's kernels (ddmi_mcubes_*), producing the reference's mesh bit for bit
(vertex and triangle order included).

Not built: MISE refinement (upsampling_steps > 0, libmise), normals estimation, mesh simplification / refinement, trimesh and
pytorch3d containers -- callers get tensors (vertices float64 (V,3), triangles int64 (T,3)) to wrap as they like."""
import contextlib
import ctypes
import functools
import math

import torch

from . import _lib
from .general_utils import make_3d_grid


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _run_mcubes(grid, isovalue, pad, pad_value, affine):
    if not (torch.is_tensor(grid) and grid.is_cuda):
        raise RuntimeError("marching cubes: the volume must be a CUDA tensor (ddmi_b200 has no CPU path)")
    if grid.dim() != 3:
        raise RuntimeError("Only three-dimensional arrays are supported.")            # pywrapper.cpp:92-93
    g = grid.detach()
    if g.dtype == torch.float64:
        if not bool((g.to(torch.float32).to(torch.float64) == g).all()):
            raise RuntimeError("marching cubes: float64 volumes must hold float32-representable values (decoded logits do)")
    g = g.to(torch.float32).contiguous()
    dev = g.device
    nx, ny, nz = g.shape
    if min(nx, ny, nz) + 2 * pad < 2:                      # no cells: the reference's sweep does not run either
        return (torch.empty((0, 3), dtype=torch.float64, device=dev), torch.empty((0, 3), dtype=torch.int64, device=dev))
    L = _lib.lib()
    nbytes = ctypes.c_uint64()
    _lib.check(L.ddmi_mcubes_workspace_bytes(nx, ny, nz, pad, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ddmi_mcubes_count(g.data_ptr(), nx, ny, nz, pad, float(pad_value), float(isovalue), ws.data_ptr(),
                                       nbytes.value, totals.data_ptr(), _stream_ptr(dev)))
        nv, nc = (int(v) for v in totals.tolist())                                      # the one host read: output sizes
        vertices = torch.empty((nv, 3), dtype=torch.float64, device=dev)
        triangles = torch.empty((nc // 3, 3), dtype=torch.int64, device=dev)
        if nv and nc:
            aff = (ctypes.c_double * 7)(*affine) if affine is not None else None
            _lib.check(L.ddmi_mcubes_emit(g.data_ptr(), nx, ny, nz, pad, float(pad_value), float(isovalue), ws.data_ptr(), aff,
                                          vertices.data_ptr(), triangles.data_ptr(), _stream_ptr(dev)))
    return vertices, triangles


def marching_cubes(volume, isovalue):
    """libmcubes.marching_cubes(volume, isovalue) -> (verts (V,3) float64, faces (T,3) int64), on the volume's device.
    As in the reference, vertex coordinates are in grid units shifted by + 0.5 (generation.py:169 undoes it)."""
    return _run_mcubes(volume, isovalue, 0, 0.0, None)


def extract_mesh(occ_hat, threshold=0.2, padding=0.1):
    """Generator3D.extract_mesh (generation.py:146-186, the vol_bound=None branch) on a device-resident logit grid
    (n_x, n_y, n_z): pad with -1e6, marching cubes at logit(threshold), normalise to the bounding box.  The padding and the
    vertex post-processing are fused into the kernels.  -> (vertices float64 (V,3), triangles int64 (T,3))."""
    n_x, n_y, n_z = occ_hat.shape
    box_size = 1 + padding
    thr = math.log(threshold) - math.log(1. - threshold)
    return _run_mcubes(occ_hat, thr, 1, -1e6, (0.5, 1.0, float(n_x - 1), float(n_y - 1), float(n_z - 1), 0.5, box_size))


def eval_points(p, c, mlp, points_batch_size=100000):
    """Generator3D.eval_points (generation.py:123-144): occupancy logits of points p (N,3) for ONE latent c = (xy, yz, xz)
    plane lists; the result stays on the device.  The reference queries chunks of points_batch_size to bound its memory
    (every chunk materialises (N, 320) features); points are independent, so chunking does not change a single logit, and the
    fused decoder streams 128-point tiles: points_batch_size is accepted and the query goes down in launches of up to 2^23
    points (21 launches of 100k points cost 2.5x the time of one launch of 2.1 M: wave quantisation + per-call host work)."""
    if p.shape[0] == 0:
        return torch.empty(0, device=p.device)
    with mlp.weights_unchanged() if hasattr(mlp, 'weights_unchanged') else contextlib.nullcontext():
        outs = [mlp(pi.unsqueeze(0), c).logits.squeeze(0).to(torch.float32) for pi in torch.split(p, max(points_batch_size, 1 << 23))]
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


@functools.lru_cache(maxsize=4)
def _query_grid(nx, box_size, dev):
    """box_size * make_3d_grid((-0.5,)*3, (0.5,)*3, (nx,)*3) on the device (25 MB at 128^3: built and uploaded once per
    resolution, not once per mesh -- the reference rebuilds it for every latent)."""
    return (box_size * make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)).to(dev)


@functools.lru_cache(maxsize=4)
def _query_axis(nx, box_size, dev):
    """The lattice's axis: row k of box_size * make_3d_grid(...) has z = box_size * linspace(-0.5, 0.5, nx)[k] (common.py:145-164
    builds the grid from these linspace values; the elementwise product gives every point of a lattice line the same fp32)."""
    return (box_size * torch.linspace(-0.5, 0.5, nx)).to(dev)


def generate_mesh(c, mlp, resolution0=128, threshold=0.2, padding=0.1, points_batch_size=100000):
    """Dense-grid path of Generator3D.generate_mesh_fromdiffusion (generation.py:84-98, upsampling_steps == 0) for one
    decoded latent c: query box_size * make_3d_grid((-0.5,)*3, (0.5,)*3, (nx,)*3), reshape to (nx, nx, nx), extract the mesh.
    The query goes down as a LATTICE (MLP3D.decode_logits_lattice: same logits bit for bit, a quarter of the gather bytes);
    decoders without that entry point get the point list through eval_points.
    -> (vertices, triangles, value_grid), all on the planes' device."""
    dev = c[0][0].device
    nx = resolution0
    box_size = 1 + padding
    if hasattr(mlp, 'decode_logits_lattice'):
        ax = _query_axis(nx, box_size, str(dev))
        value_grid = mlp.decode_logits_lattice((ax, ax, ax), c)[0]
    else:
        pointsf = _query_grid(nx, box_size, str(dev))
        value_grid = eval_points(pointsf, c, mlp, points_batch_size).reshape(nx, nx, nx)
    vertices, triangles = extract_mesh(value_grid, threshold, padding)
    return vertices, triangles, value_grid
