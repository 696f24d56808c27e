"""CPU: the N>1 host logic under a world_size-2 gloo group (no GPU needed).
The decoder itself is stubbed by the CPU oracle here -- this tests the sharding /
gather plumbing only, never the product kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddmi_b200 import sharding


def test_plan_units_covers_everything_once():
    for batch, rows, world in [(64, 1024, 8), (4, 256, 8), (1, 7, 4), (3, 5, 2), (5, 1, 8), (16, 128, 1)]:
        plan = sharding.plan_units(batch, rows, world)
        assert len(plan) == world
        seen = torch.zeros(batch, rows, dtype=torch.int32)
        for units in plan:
            for item, r0, r1 in units:
                assert 0 <= r0 < r1 <= rows
                seen[item, r0:r1] += 1
        assert bool((seen == 1).all()), (batch, rows, world)
        sizes = [sum(r1 - r0 for _, r0, r1 in u) for u in plan]
        if batch * rows >= world:
            assert min(sizes) > 0
        if rows >= world:                      # enough rows to cut: every rank owns the same number of units, rows within 1 slab
            assert len({len(u) for u in plan}) == 1, (batch, rows, world)
            assert max(sizes) - min(sizes) <= len(plan[0]), (batch, rows, world, sizes)
    assert [len(u) for u in sharding.plan_units(3, 64, 4)] == [3, 3, 3, 3]      # was 2:1 imbalanced with round-robin slabs
    assert sharding.plan_units(64, 1024, 8)[3] == [(i, 0, 1024) for i in range(24, 32)]   # whole items when the batch divides


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import cases, ddmi_oracle as orc
    torch.set_grad_enabled(False)
    sd = cases.state_dict32(cases.build_module('image'))
    coords, planes, si = cases.image_inputs(batch=3, sizes=(4, 8, 16), res=12)

    class CpuStub:   # stands in for the CUDA decoder: same call signature
        def __call__(self, c, hdbf, si=1):
            return orc.image_decode(sd, c, hdbf, si)

    full = sharding.decode_image_sharded(CpuStub(), coords, planes, si=si, gather=True)
    ref = orc.image_decode(sd, coords, planes, si)
    mine = sharding.decode_image_sharded(CpuStub(), coords, planes, si=si, gather=False)
    err = float((full - ref).abs().max())
    at_root = sharding.decode_image_sharded(CpuStub(), coords, planes, si=si, gather='root')
    assert (at_root is None) == (rank != 0)
    if rank == 0:
        err = max(err, float((at_root - ref).abs().max()))
    # round-1 API: partial full-size tensors in, assembled tensor out (only the owned slabs travel)
    partial = torch.zeros_like(ref)
    for (item, r0, r1), t in mine:
        partial[item, :, r0:r1] = t
    err = max(err, float((sharding.all_gather_outputs(partial, sharding.plan_units(3, 12, world)) - ref).abs().max()))

    # occupancy: items then point ranges
    sdo = cases.state_dict32(cases.build_module('occupancy'))
    pts, hdbf = cases.occupancy_inputs(batch=3, sizes=(4, 8, 16), n=50)

    class OccStub:
        def decode_logits(self, p, c):
            return orc.occupancy_logits(sdo, p, c)

        def decode_logits_lattice(self, axes, c):
            nx, ny, nz = (a.numel() for a in axes)
            return orc.occupancy_logits(sdo, torch.cartesian_prod(*axes)[None], c).reshape(-1, nx, ny, nz)

    lo = sharding.decode_occupancy_sharded(OccStub(), pts, hdbf, gather=True)
    err = max(err, float((lo - orc.occupancy_logits(sdo, pts, hdbf)).abs().max()))
    # lattice queries: items then x-slabs of the lattice
    axes = (torch.linspace(-.5, .5, 5), torch.linspace(-.4, .5, 3), torch.linspace(-.5, .3, 4))
    la = sharding.decode_occupancy_lattice_sharded(OccStub(), axes, hdbf, gather=True)
    ref_l = orc.occupancy_logits(sdo, torch.cartesian_prod(*axes)[None].expand(3, -1, -1), hdbf).reshape(3, 5, 3, 4)
    assert la.shape == ref_l.shape
    err = max(err, float((la - ref_l).abs().max()))

    # video: batch items
    sdv = cases.state_dict32(cases.build_module('video'))
    cv, hv = cases.video_inputs(batch=3, T=2, sizes=(4, 4, 8))

    class VidStub:
        def __call__(self, c, h):
            return orc.video_decode(sdv, c, h)

    vo = sharding.decode_video_sharded(VidStub(), cv, hv, gather=True)
    err = max(err, float((vo - orc.video_decode(sdv, cv, hv)).abs().max()))
    q.put((rank, err, [u for u, _ in mine]))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_decode_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    units = []
    for rank, err, mine in res:
        assert err < 1e-5, (rank, err)
        units += mine
    assert sorted(units) == [(0, 0, 6), (0, 6, 12), (1, 0, 6), (1, 6, 12), (2, 0, 6), (2, 6, 12)]
