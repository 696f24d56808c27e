"""Multi-GPU sharding of a decode (one process per GPU, torch.distributed).

The path has no exchange step: every (batch item, coordinate) pair is independent
given that item's planes, so ranks take disjoint work units and no
collective runs on the hot path.  Units are (item, slab) pairs: the query rows of an
item (image rows, point ranges, whole rays) are cut into `slabs` equal slabs, with
`slabs` chosen so that `batch * slabs` divides evenly over the ranks -- every rank
owns the same number of units (no 2:1 imbalance).  The one optional collective
assembles the full signal: each rank contributes ONLY its own units, packed into one
equal-sized buffer (`all_gather_into_tensor`, or `gather` to a root) -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  The reference decodes on rank 0 only
(tools/ldm/image.py:179-189); this is the repo's own contract (BASELINE north_star).
"""
import math
from typing import List, Tuple

import torch
import torch.distributed as dist

Unit = Tuple[int, int, int]       # (item, row0, row1)


def slabs_for(batch: int, rows: int, world: int) -> int:
    """Slabs per item so that batch * slabs is a multiple of world (1 when the batch already divides), capped by rows."""
    return max(1, min(rows, world // math.gcd(batch, world)))


def plan_units(batch: int, rows: int, world: int) -> List[List[Unit]]:
    """Deal (item, row0, row1) units to `world` ranks: item-major order, consecutive blocks, so a rank's units are
    neighbours (whole items when the batch divides).  Every unit is owned by exactly one rank, the union covers
    batch x rows, and ranks own equally many units whenever rows >= the slab count."""
    if batch < 1 or rows < 1 or world < 1:
        raise ValueError("batch, rows and world must be >= 1")
    slabs = slabs_for(batch, rows, world)
    units = []
    for item in range(batch):
        for s in range(slabs):
            r0, r1 = rows * s // slabs, rows * (s + 1) // slabs
            if r1 > r0:
                units.append((item, r0, r1))
    out: List[List[Unit]] = [[] for _ in range(world)]
    base, extra = divmod(len(units), world)
    k = 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out[r] = units[k:k + n]
        k += n
    return out


def _world_rank(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def assemble(results, plan, shape, rows_dim, group=None, gather='all', root=0):
    """Assemble per-rank unit results into the full signal.

    results: this rank's list of (unit, tensor) where `tensor` has the item's shape with the row axis (`rows_dim` of the
    full per-item shape) cut to the unit's rows.  shape: full output shape (batch, ...).  Each rank packs its units into one
    buffer padded to the largest unit (equal size on every rank) and ONE collective moves exactly those bytes:
    `gather='all'` -> all_gather_into_tensor, every rank returns the full tensor; `'root'` -> dist.gather, only `root`
    returns it (others None)."""
    world, rank = _world_rank(group)
    ref = results[0][1] if results else None
    device = ref.device if ref is not None else torch.device('cpu')
    dtype = ref.dtype if ref is not None else torch.float32
    full = torch.empty(shape, device=device, dtype=dtype)

    def place(dst, unit, t):
        item, r0, r1 = unit
        dst[item].narrow(rows_dim - 1, r0, r1 - r0).copy_(t.narrow(rows_dim - 1, 0, r1 - r0))

    if world == 1:
        for unit, t in results:
            place(full, unit, t)
        return full
    per_rank = max(len(p) for p in plan)
    max_rows = max(r1 - r0 for p in plan for (_, r0, r1) in p)
    unit_shape = list(shape[1:])
    unit_shape[rows_dim - 1] = max_rows
    send = torch.zeros([per_rank] + unit_shape, device=device, dtype=dtype)
    for k, (unit, t) in enumerate(results):
        send[k].narrow(rows_dim - 1, 0, unit[2] - unit[1]).copy_(t)
    if gather == 'all':
        recv = torch.empty([world * per_rank] + unit_shape, device=device, dtype=dtype)
        dist.all_gather_into_tensor(recv, send, group=group)
    elif gather == 'root':
        parts = [torch.empty_like(send) for _ in range(world)] if rank == root else None
        dist.gather(send, parts, dst=root, group=group)
        if rank != root:
            return None
        recv = torch.cat(parts)
    else:
        raise ValueError("gather must be 'all' or 'root'")
    for r in range(world):
        for k, unit in enumerate(plan[r]):
            place(full, unit, recv[r * per_rank + k])
    return full


def _finish(results, plan, shape, rows_dim, group, gather):
    if not gather:
        return results
    return assemble(results, plan, shape, rows_dim, group, 'all' if gather is True else gather)


def decode_image_sharded(mlp, coords, hdbf, si=1, group=None, gather=False, **kw):
    """Image decode with the batch (then row slabs) sharded over the process group.

    `hdbf` holds the FULL batch on every rank (planes are <= 26 MB per item; replicate or broadcast them once); each rank
    decodes only its units.  Returns this rank's list of ((item, row0, row1), tensor (3, row1-row0, w)); with
    gather=True / 'all' the assembled (b, 3, h, w) on every rank, with gather='root' on rank 0 only (None elsewhere)."""
    world, rank = _world_rank(group)
    b = hdbf[0].shape[0]
    _, _, h, w = coords.shape
    plan = plan_units(b, h, world)
    results = []
    mine = plan[rank]
    k = 0
    while k < len(mine):                       # consecutive whole items go down as ONE batched launch
        item, r0, r1 = mine[k]
        j = k
        if (r0, r1) == (0, h):
            while j + 1 < len(mine) and mine[j + 1] == (mine[j][0] + 1, 0, h):
                j += 1
        out = mlp(coords[:, :, r0:r1, :], hdbf=[p[item:mine[j][0] + 1] for p in hdbf], si=si, **kw)
        for i in range(k, j + 1):
            results.append((mine[i], out[i - k]))
        k = j + 1
    return _finish(results, plan, (b, 3, h, w), 2, group, gather)


def decode_video_sharded(mlp, coords, hdbf, group=None, gather=False, **kw):
    """Video decode sharded over batch ITEMS (the reference's forward takes the (t,h,w) of its output from the planes, so an
    item is the unit; SkyTimelapse batches are 16).  Returns [((item, 0, t), tensor (3,t,h,w))] or the assembled
    (b,3,t,h,w)."""
    world, rank = _world_rank(group)
    xy, yt, xt = hdbf
    b, _, h, w = xy[-1].shape
    t = yt[-1].shape[2]
    base, extra = divmod(b, world)
    plan, k = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        plan.append([(i, 0, t) for i in range(k, k + n)])
        k += n
    results = []
    if plan[rank]:
        i0, i1 = plan[rank][0][0], plan[rank][-1][0] + 1
        out = mlp(coords, tuple([p[i0:i1] for p in axis] for axis in hdbf), **kw)
        results = [(u, out[j]) for j, u in enumerate(plan[rank])]
    return _finish(results, plan, (b, 3, t, h, w), 2, group, gather)


def decode_occupancy_sharded(mlp, points, hdbf, group=None, gather=False):
    """Occupancy logits with items, then point ranges, sharded.  points (B,N,3) or (1,N,3) shared by all items.
    Returns [((item, n0, n1), logits (n1-n0,))] or the assembled (B,N)."""
    world, rank = _world_rank(group)
    b = hdbf[0][0].shape[0]
    n = points.shape[1]
    plan = plan_units(b, n, world)
    results = []
    for item, n0, n1 in plan[rank]:
        pi = points[item if points.shape[0] > 1 else 0, n0:n1][None]
        c = tuple([p[item:item + 1] for p in axis] for axis in hdbf)
        results.append(((item, n0, n1), mlp.decode_logits(pi, c)[0]))
    return _finish(results, plan, (b, n), 1, group, gather)


def decode_occupancy_lattice_sharded(mlp, axes, hdbf, group=None, gather=False):
    """Occupancy logits on the query lattice axes = (xs, ys, zs) (MLP3D.decode_logits_lattice: the mesh generator's dense grid)
    with items, then x-slabs of the lattice, sharded: a rank decodes the sub-lattice (xs[i0:i1], ys, zs) of its items.
    Returns [((item, i0, i1), logits (i1-i0, ny, nz))] or the assembled (B, nx, ny, nz)."""
    world, rank = _world_rank(group)
    b = hdbf[0][0].shape[0]
    xs, ys, zs = axes
    nx, ny, nz = int(xs.numel()), int(ys.numel()), int(zs.numel())
    plan = plan_units(b, nx, world)
    results = []
    for item, i0, i1 in plan[rank]:
        c = tuple([p[item:item + 1] for p in axis] for axis in hdbf)
        results.append(((item, i0, i1), mlp.decode_logits_lattice((xs[i0:i1], ys, zs), c)[0]))
    return _finish(results, plan, (b, nx, ny, nz), 1, group, gather)


def render_rays_sharded(module, rays, fea, N_samples, white_bkgd, group=None, gather=False, **kw):
    """NeRF render with objects, then ray ranges (whole rays: compositing is per ray), sharded.
    Returns [((object, ray0, ray1), rgb (ray1-ray0, 3))] or the assembled (B, N_rays, 3)."""
    from . import nerf_helpers as nh
    world, rank = _world_rank(group)
    b = fea['xy'].shape[0]
    n = rays.shape[0]
    plan = plan_units(b, n, world)
    results = []
    for item, r0, r1 in plan[rank]:
        fb = {k: v[item:item + 1] for k, v in fea.items()}
        results.append(((item, r0, r1), nh.render_rays_fused(rays[r0:r1], fb, module, N_samples, white_bkgd, **kw)[0]))
    return _finish(results, plan, (b, n, 3), 1, group, gather)


def all_gather_outputs(full, plan, group=None):
    """Compatibility wrapper (round-1 API): `full` holds this rank's units at their final positions; assemble every rank's
    units on every rank, moving only the owned slabs."""
    world, rank = _world_rank(group)
    if world == 1:
        return full
    results = [((item, r0, r1), full[item, :, r0:r1]) for item, r0, r1 in plan[rank]]
    return assemble(results, plan, tuple(full.shape), 2, group, 'all')
