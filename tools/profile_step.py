"""Small driver for ncu: builds the BASELINE configs[1] state (or XPIC_BENCH_GRID) and runs a few
ECSIM steps through the C ABI.  Used only for profiling; bench.py is the measurement."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpic_b200 as X

grid = tuple(int(v) for v in os.environ.get("XPIC_BENCH_GRID", "128,128,128").split(","))
ppc = int(os.environ.get("XPIC_BENCH_PPC", "64"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scheme = {"ecsim": X.ECSIM, "ecsimcorr": X.ECSIMCORR, "eccapfim": X.ECCAPFIM}[os.environ.get("XPIC_SCHEME", "ecsim")]
sim = X.Simulation(grid, scheme=scheme, track_ids=False)
sid = sim.add_species(Np=ppc, capacity=int(sim.ncl * ppc * 1.25) + 65536)
sim.set_particles_maxwellian(sid, grid[0] * grid[1] * grid[2] * ppc, T=0.1, seed=20261018)
p = int(os.environ.get("XPIC_BENCH_PRECOND", "6"))
sim.solver_set(0, 1e-7, 1e-7, 100, 30, p)
sim.solver_set(1, 1e-7, 1e-7, 100, 30, p)
if os.environ.get("XPIC_PROFILE_RANGE"):
    # ncu --profile-from-start off: only whole steady-state steps are captured (two warm-up steps first)
    import ctypes

    rt = ctypes.CDLL("libcudart.so")
    sim.run_steps(2)
    rt.cudaProfilerStart()
    ms = sim.run_steps(steps)
    rt.cudaProfilerStop()
else:
    ms = sim.run_steps(steps)
if scheme == X.ECCAPFIM:
    print("nonlinear", sim.nonlinear_info())
print("steps", steps, "ms/step", ms / steps, "its", sim.solver_info(0)[0], {k: round(1e3 * v[0] / max(v[1], 1), 3) for k, v in sim.timing().items()})
if os.environ.get("XPIC_KERNEL_BENCH"):
    for what, name in ((0, "sort (no move)"), (1, "deposit (fields + cell blocks + gather)"), (2, "second push"), (3, "solve")):
        print(f"kernel_bench {name}: {sim.kernel_bench(what, 5):.3f} ms")
