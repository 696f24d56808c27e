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
    q.put((rank, float((full - ref).abs().max()), [u for u, _ in mine]))
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
    assert sorted(units) == [(0, 0, 12), (1, 0, 12), (2, 0, 12)]
