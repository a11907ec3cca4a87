"""BASELINE config 4: the implicit Maxwell operator (L + M, fixed-offset stencil layout = 3x3-block rows
without indices) on a deposited Maxwellian plasma, grid sweep on one GPU: SpMV bandwidth against the
measured HBM copy bandwidth and the cost of the preconditioned GMRES solve.

    python tools/spmv_sweep.py [ppc] > profiles/rNN_spmv_sweep.json

Algorithmic bytes: 3000 B per cell and SpMV (369 coefficients x 8 B + 24 B x + 24 B y).
The operator's values depend on ppc, its layout and size do not; a small ppc keeps the 192^3 case
(20.9 GB operator + 75 GB deposit staging) inside one GPU's memory."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpic_b200 as X

ppc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for n in (64, 96, 128, 160, 192):
    sim = X.Simulation((n, n, n), scheme=X.ECSIM, track_ids=False)
    sid = sim.add_species(Np=ppc, capacity=int(sim.ncl * ppc * 1.25) + 65536)
    sim.set_particles_maxwellian(sid, n * n * n * ppc, T=0.1, seed=20261018)
    sim.solver_set(0, 1e-7, 1e-7, 100, 30, 6)
    sim.run_steps(2)  # deposits the operator, settles the fields
    ms_a = sim.spmv_bench(X.binding.OP_A, 100)
    ms_m = sim.spmv_bench(X.binding.OP_M, 100)
    t_solve = sim.kernel_bench(3, 3)
    its = sim.solver_info(0)[0]
    cells = n ** 3
    rows.append({"grid": f"{n}^3", "cells": cells, "ppc": ppc, "spmv_LM_ms": ms_a, "spmv_LM_GBs": 3000.0 * cells / ms_a / 1e6,
                 "frac_of_measured_hbm": 3000.0 * cells / ms_a / 1e6 / peak, "spmv_M_ms": ms_m, "spmv_M_GBs": 48.0 * cells / ms_m / 1e6,
                 "gmres_iterations": its, "gmres_solve_ms": t_solve, "operator_GB": 369 * 8 * cells / 1e9})
    print(rows[-1], file=sys.stderr, flush=True)
    sim.close()
print(json.dumps({"peak_hbm_gbs": peak, "krylov": "GMRES(30) rtol=atol=1e-7, Chebyshev(M) degree 6", "rows": rows}, indent=1))
