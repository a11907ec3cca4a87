"""Decomposition-independence check, run under torchrun with N >= 2 ranks (one GPU each):
the z-slab run must reproduce the single-GPU run of the same box (the reference demands the
same of its 1- and 2-rank ctest runs, tests/ecsim/CMakeLists.txt:14-16).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import xpic_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = (12, 8, 8 * world)
    steps = int(os.environ.get("XPIC_CHECK_STEPS", "6"))
    ok = True
    # the three schemes in a periodic box, then ECSIM with an open z boundary (particles leave through the outer slabs)
    # ... and ECSIM with the slabs depositing in batches of planes (the staging mode of large slabs)
    cases = ((X.ECSIM, O.ECSIM, False, None), (X.ECSIMCORR, O.ECSIMCORR, False, None), (X.ECCAPFIM, O.ECCAPFIM, False, None), (X.ECSIM, O.ECSIM, True, None),
             (X.ECSIM, O.ECSIM, False, "5.2e-3"))
    for scheme, oscheme, open_z, stage_gb in cases:
        ids = [X.comm_unique_id() if rank == 0 else None]  # one communicator id per context
        dist.broadcast_object_list(ids, src=0)
        o = O.Oracle(n)  # only used for the reference's mt19937 initial particles
        sid = o.add_species(Np=20)
        o.set_particles_maxwell(sid, 2.0 if open_z else 0.1, True)
        pts, pid = o.get_particles(sid)
        rng = np.random.default_rng(5)
        B0 = 0.05 * rng.standard_normal(o.n3)
        if stage_gb:
            os.environ["XPIC_STAGE_GB"] = stage_gb  # 5 planes of 12 x 8 cells: batches of 3 planes
        slab = X.Simulation(n, scheme=scheme, device=local, rank=rank, nranks=world, comm_id=ids[0], track_ids=True, open_z=open_z)
        os.environ.pop("XPIC_STAGE_GB", None)
        slab.add_species(Np=20, capacity=len(pid))
        mine = slab.add_particles(0, pts, pid)
        lo, hi = 3 * n[0] * n[1] * slab.z0, 3 * n[0] * n[1] * (slab.z0 + slab.nzl)
        slab.set_field("B", B0[lo:hi])
        for w in (0, 1):
            slab.solver_set(w, 1e-12, 1e-50, 1000, 30, 4)
        slab.nonlinear_set(atol=1e-13, rtol=1e-30, particle_tol=1e-14)
        for _ in range(steps):
            slab.step()
        cnt = torch.tensor([slab.particle_count(0)], device="cuda")
        dist.all_reduce(cnt)
        parts = [None] * world
        fe, fb = slab.get_field("E"), slab.get_field("B")
        gathered = [None] * world
        dist.all_gather_object(gathered, (fe, fb, slab.get_particles(0)))
        if rank == 0:
            single = X.Simulation(n, scheme=scheme, device=local, track_ids=True, open_z=open_z)
            single.add_species(Np=20, capacity=len(pid))
            single.add_particles(0, pts, pid)
            single.set_field("B", B0)
            for w in (0, 1):
                single.solver_set(w, 1e-12, 1e-50, 1000, 30, 4)
            single.nonlinear_set(atol=1e-13, rtol=1e-30, particle_tol=1e-14)
            for _ in range(steps):
                single.step()
            E = np.concatenate([g[0] for g in gathered])
            B = np.concatenate([g[1] for g in gathered])
            P = np.concatenate([g[2][0] for g in gathered])
            I = np.concatenate([g[2][1] for g in gathered])
            order = np.argsort(I)
            Ps, Is = single.get_particles(0)
            so = np.argsort(Is)
            eE = np.linalg.norm(E - single.get_field("E")) / np.linalg.norm(single.get_field("E"))
            eB = np.linalg.norm(B - single.get_field("B")) / np.linalg.norm(single.get_field("B"))
            eP = np.linalg.norm(P[order] - Ps[so]) / np.linalg.norm(Ps[so])
            left = len(pid) - len(Is)
            good = int(cnt.item()) == len(Is) and (left > 0) == open_z and np.array_equal(I[order], Is[so]) and max(eE, eB, eP) < 1e-9
            print(f"scheme {scheme}{' open z' if open_z else ''}{' batched staging' if stage_gb else ''}: ranks {world} particles {int(cnt.item())}/{len(pid)} relerr E {eE:.2e} B {eB:.2e} particles {eP:.2e} -> {'OK' if good else 'FAIL'}",
                  flush=True)
            ok = ok and good
            single.close()
        slab.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
