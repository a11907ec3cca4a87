"""world_size-2 gloo test (CPU) of the host-side multi-rank logic bench.py and the tests rely on:
communicator-id broadcast, slab ownership of the reference's initial particle stream (every rank
walks the whole stream and keeps its own particles, src/interfaces/particles.cpp:47-57), and the
max / sum reductions of the bench contract."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import xpic_b200 as X
from oracle import oracle as O


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch

        ids = [os.urandom(128) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        n = (6, 6, 7)  # odd plane count: slabs of 4 and 3 planes
        o = O.Oracle(n)
        sid = o.add_species(Np=10)
        o.set_particles_maxwell(sid, 0.1, True)
        pts, pid = o.get_particles(sid)
        z0, nzl = X.slab_range(n[2], rank, world)
        cz = np.floor(pts[:, 2] / 0.5).astype(int)
        mine = (cz >= z0) & (cz < z0 + nzl)
        owners = np.array([X.owner_rank(int(k), n[2], world) for k in cz])
        assert np.array_equal(mine, owners == rank)
        cnt = torch.tensor([float(mine.sum())], dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.put((rank, ids[0], int(cnt.item()), len(pid), float(t.item()), z0, nzl))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and len(res[0][1]) == 128  # same communicator id on both ranks
    assert res[0][2] == res[0][3] == res[1][2]  # every particle of the stream has exactly one owner
    assert res[0][4] == res[1][4] == 11.0  # max over ranks
    assert (res[0][5], res[0][6]) == (0, 4) and (res[1][5], res[1][6]) == (4, 3)
