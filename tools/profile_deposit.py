"""Times the moment-deposition variants on the BASELINE configs[1] state (or XPIC_BENCH_GRID): xb_kernel_bench(1) per
variant after two real steps (so the particles are in their steady-state, partially mixed order), plus the in-step
family clocks.  Also the driver of the ncu capture of k_cell_moments (profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpic_b200 as X

grid = tuple(int(v) for v in os.environ.get("XPIC_BENCH_GRID", "128,128,128").split(","))
ppc = int(os.environ.get("XPIC_BENCH_PPC", "64"))
variants = [int(v) for v in os.environ.get("XPIC_DEPOSIT_VARIANTS", "0,3,2").split(",")]
sim = X.Simulation(grid, scheme=X.ECSIM, track_ids=False)
sid = sim.add_species(Np=ppc, capacity=int(sim.ncl * ppc * 1.25) + 65536)
sim.set_particles_maxwellian(sid, grid[0] * grid[1] * grid[2] * ppc, T=0.1, seed=20261018)
sim.solver_set(0, 1e-7, 1e-7, 100, 30, 6)
sim.run_steps(2)
out = {"grid": grid, "ppc": ppc}
for v in variants:
    sim.set_option(0, v)
    out[f"deposit_variant_{v}_ms"] = sim.kernel_bench(1, 3)
if os.environ.get("XPIC_BACKOFF_SWEEP"):
    sim.set_option(0, 0)
    for ns in (0, 32, 64, 128, 256, 512):
        sim.set_option(5, ns)
        out[f"deposit_ws_backoff_{ns}ns_ms"] = sim.kernel_bench(1, 3)
    sim.set_option(5, 64)
sim.set_option(0, variants[0])
sim.family_profile(True)
ms = sim.run_steps(3)
fam = sim.family_profile_read()
out["ms_per_step"] = ms / 3
out["families_ms_per_step"] = {k: v[1] / 3 for k, v in fam.items()}
out["stage_ms"] = {k: round(1e3 * v[0] / max(v[1], 1), 3) for k, v in sim.timing().items()}
out["second_push_ms"] = sim.kernel_bench(2, 3)
out["solve_ms"] = sim.kernel_bench(3, 3)
print(json.dumps(out))
