"""BASELINE config 4: the implicit Maxwell operator (L + M, fixed-offset stencil layout = 3x3-block rows
without indices) on a deposited Maxwellian plasma, grid sweep: SpMV bandwidth against the measured HBM copy
bandwidth and the cost of the preconditioned GMRES solve.

    python tools/spmv_sweep.py [ppc] [grids] > profiles/rNN_spmv_sweep.json                # one GPU: 64 .. 256
    python -m torch.distributed.run --nproc-per-node 8 ... tools/spmv_sweep.py 8 512        # 512^3 in 8 z-slabs

Algorithmic bytes: 3000 B per cell and SpMV (369 coefficients x 8 B + 24 B x + 24 B y).
The operator's values depend on ppc, its layout and size do not; a small ppc and the batched staging of the
deposit (XPIC_STAGE_GB) keep the 256^3 case (49.5 GB operator) inside one GPU's memory.  On several GPUs the
time of an SpMV is the maximum over the ranks (CUDA events, barrier on both sides) and includes the halo exchange."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("XPIC_STAGE_GB", "40")
import xpic_b200 as X

ppc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grids = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 96, 128, 160, 192, 256]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for n in grids:
    comm_id = None
    if world > 1:
        ids = [X.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm_id = ids[0]
    sim = X.Simulation((n, n, n), scheme=X.ECSIM, track_ids=False, device=local, rank=rank, nranks=world, comm_id=comm_id)
    sid = sim.add_species(Np=ppc, capacity=int(sim.ncl * ppc * 1.25) + 65536)
    sim.set_particles_maxwellian(sid, n * n * n * ppc, T=0.1, seed=20261018)
    sim.solver_set(0, 1e-7, 1e-7, 100, 30, 6)
    sim.run_steps(2)  # deposits the operator, settles the fields
    if world > 1:
        dist.barrier()
    ms_a = allmax(sim.spmv_bench(X.binding.OP_A, 100))
    ms_m = allmax(sim.spmv_bench(X.binding.OP_M, 100))
    t_solve = allmax(sim.kernel_bench(3, 3))
    its = sim.solver_info(0)[0]
    cells = n ** 3
    rows.append({"grid": f"{n}^3", "cells": cells, "gpus": world, "ppc": ppc, "spmv_LM_ms": ms_a, "spmv_LM_GBs": 3000.0 * cells / ms_a / 1e6,
                 "frac_of_measured_hbm": 3000.0 * cells / ms_a / 1e6 / (peak * world), "spmv_M_ms": ms_m, "spmv_M_GBs": 48.0 * cells / ms_m / 1e6,
                 "gmres_iterations": its, "gmres_solve_ms": t_solve, "operator_GB": 369 * 8 * cells / 1e9})
    if rank == 0:
        print(rows[-1], file=sys.stderr, flush=True)
    sim.close()
if rank == 0:
    print(json.dumps({"peak_hbm_gbs_per_gpu": peak, "gpus": world, "krylov": "GMRES(30) rtol=atol=1e-7, Chebyshev(M) degree 6", "rows": rows}, indent=1))
if world > 1:
    dist.destroy_process_group()
