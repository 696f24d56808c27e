"""Multi-GPU sharding of a decode (one process per GPU, torch.distributed).

The path has no exchange step: every (batch item, coordinate) pair is independent
given that item's planes, so ranks take disjoint work units and no
collective runs on the hot path.  Units are (item, row-slab) pairs: batch items are
dealt out first; when there are fewer items than ranks each item's query rows are
split into slabs (whole rows for image grids, whole rays for NeRF) so every rank
has work.  `all_gather_outputs` is the one optional collective: it assembles the
full signal on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def plan_units(batch: int, rows: int, world: int) -> List[List[Tuple[int, int, int]]]:
    """Deal (item, row0, row1) units to `world` ranks.

    * batch >= world: items are split as evenly as possible, rows untouched;
    * batch <  world: each item's rows are cut into `ceil(world / batch)` slabs (never
      finer than one row) and the slabs are dealt round-robin.
    Every unit is owned by exactly one rank and the union covers batch x rows.
    """
    if batch < 1 or rows < 1 or world < 1:
        raise ValueError("batch, rows and world must be >= 1")
    out: List[List[Tuple[int, int, int]]] = [[] for _ in range(world)]
    if batch >= world:
        base, extra = divmod(batch, world)
        item = 0
        for r in range(world):
            for _ in range(base + (1 if r < extra else 0)):
                out[r].append((item, 0, rows))
                item += 1
        return out
    slabs = min(rows, -(-world // batch))
    units = []
    for item in range(batch):
        for s in range(slabs):
            r0, r1 = rows * s // slabs, rows * (s + 1) // slabs
            if r1 > r0:
                units.append((item, r0, r1))
    for i, u in enumerate(units):
        out[i % world].append(u)
    return out


def decode_image_sharded(mlp, coords, hdbf, si=1, group=None, gather=False):
    """Image decode with the batch (then row slabs) sharded over the process group.

    `hdbf` holds the FULL batch on every rank (planes are <= 26 MB per item; replicate
    or broadcast them once); each rank decodes only its units.  Returns this rank's
    list of ((item, row0, row1), tensor (3, row1-row0, w)); with `gather=True` returns
    the assembled (b, 3, h, w) on every rank instead.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b = hdbf[0].shape[0]
    _, _, h, w = coords.shape
    mine = plan_units(b, h, world)[rank]
    results = []
    for item, r0, r1 in mine:
        sub = coords[:, :, r0:r1, :]
        out = mlp(sub, hdbf=[p[item:item + 1] for p in hdbf], si=si)   # (1,3,r1-r0,w)
        results.append(((item, r0, r1), out[0]))
    if not gather:
        return results
    full = torch.zeros((b, 3, h, w), device=hdbf[0].device, dtype=torch.float32)
    for (item, r0, r1), t in results:
        full[item, :, r0:r1] = t
    return all_gather_outputs(full, plan_units(b, h, world), group)


def all_gather_outputs(full, plan, group=None):
    """Assemble per-rank partial (b,3,h,w) tensors into the full signal on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return full
    world = dist.get_world_size(group)
    parts = [torch.empty_like(full) for _ in range(world)]
    dist.all_gather(parts, full.contiguous(), group=group)
    out = torch.zeros_like(full)
    for r in range(world):
        for item, r0, r1 in plan[r]:
            out[item, :, r0:r1] = parts[r][item, :, r0:r1]
    return out
