#!/usr/bin/env python
"""bench.py -- ECSIM particle-steps/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port; PETSc reference unbuildable here)

Workload: BASELINE.json configs[1] at N = 1 (ECSIM 3D 128^3 cells x 64 ppc, fp64); for N > 1 the
per-GPU work is kept (weak scaling): 256 x 256 x (32 N) cells in z-slabs of 32 planes, which at
N = 8 is BASELINE configs[2]'s 256^3 x 64 ppc box.  A "step" is one timestep_implementation():
push + re-binning + moment deposition (current, mass matrices) + GMRES solve + Boris update +
field update.  Inputs are synthetic: Maxwellian electrons (T = 0.1 keV, q = -1, m = 1, n = 1,
dx = 0.5, dt = 1.5, periodic box, E = B = 0 at t = 0) from a counter-based generator.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ecsim_particle_steps_per_s"
UNIT = "particle-steps/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload(n_gpus, scheme="ecsim"):
    grid = os.environ.get("XPIC_BENCH_GRID")
    ppc = env_int("XPIC_BENCH_PPC", 32 if scheme == "eccapfim" else 64)
    if grid:
        n = tuple(int(v) for v in grid.split(","))
    elif scheme == "eccapfim":
        n = (192, 192, 24 * n_gpus)  # BASELINE configs[4]: 192^3 x 32 ppc on 8 GPUs, the same slab per GPU for fewer
    elif n_gpus == 1:
        n = (128, 128, 128)
    else:
        n = (256, 256, 32 * n_gpus)
    return n, ppc


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_sample(steps=4, warm=1, n=(32, 32, 32), ppc=64, scheme="ecsim"):
    """The oracle (C++ port of the reference's algorithm with its OpenMP loop structure; the PETSc
    reference cannot be built in this image) timed on all host cores on a bounded sample of the workload.
    Returns (particle-steps/s, particles, s/step, solver iterations [eccapfim: residual evaluations], cores)."""
    from oracle import oracle as O

    cores = O.max_threads()
    O.set_threads(cores)
    o = O.Oracle(n)
    sid = o.add_species(Np=ppc)
    N = o.set_particles_maxwell(sid, 0.1, True)
    o.solver_set(0, 1e-7, 1e-7, 100, 30)
    o.solver_set(1, 1e-7, 1e-7, 100, 30)
    # eccapfim: the oracle's NGMRES restatement on the preconditioned residual (~25 evaluations per step;
    # the reference's unpreconditioned NGMRES takes ~105, golden convergence_history.txt)
    o.snes_set(atol=1e-7, rtol=1e-7, precond=1, shift=0.5)
    code = {"ecsim": O.ECSIM, "ecsimcorr": O.ECSIMCORR, "eccapfim": O.ECCAPFIM}[scheme]
    for _ in range(warm):
        o.step(code)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(code)
    dt = time.perf_counter() - t0
    O.set_threads(1)
    its = o.snes_info()["fevals"] if scheme == "eccapfim" else o.solver_info(0)[0]
    return N * steps / dt, N, dt / steps, its, cores


def cpu_baseline_subprocess(scheme, steps=4, warm=1):
    """The CPU arm in its own process (keeps the oracle's shared object out of the GPU arm's address space)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--scheme", scheme, "--steps", str(steps), "--warmup", str(warm)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
        for ln in reversed(out.stdout.splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"unavailable": "no JSON line from the CPU arm: " + out.stderr[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def measure_fp64_peak():
    """tools/fp64_peak.cu (built into xpic_b200/_build/fp64_peak): DFMA / DMMA issue rates of this GPU, CUDA events."""
    exe = os.path.join(ROOT, "xpic_b200", "_build", "fp64_peak")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    n, ppc_w = workload(args.gpus, args.scheme)
    n_s, ppc = ((24, 24, 24), ppc_w) if args.scheme == "eccapfim" else ((32, 32, 32), 64)
    value, N, sps, its, cores = cpu_baseline_sample(steps=args.steps, warm=args.warmup, n=n_s, ppc=ppc, scheme=args.scheme)
    dt = sps * args.steps
    sample = (f"each step = one {args.scheme.upper()} step of a {n_s[0]}^3-cell x {ppc} ppc sample of the workload ({N} particles), {cores} OpenMP threads"
              + (f", {its} residual evaluations per step" if args.scheme == "eccapfim" else ""))
    line = {
        "impl": "reference", "metric": METRIC if args.scheme == "ecsim" else f"{args.scheme}_particle_steps_per_s", "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.scheme.upper()} 3D {n[0]}x{n[1]}x{n[2]} cells x {ppc_w} ppc fp64 (timed on the sample below)",
                   "solver_iterations_per_step": its},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "xpic needs MPI + PETSc, neither is in this image (no network): the arm times oracle/ (C++ restatement; ecsim: GMRES(30) unpreconditioned)",
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(local_rank):
    """Moves this process onto the CPUs NVML lists as local to its GPU; returns the previous affinity (None when NVML
    is not available: the run is then simply not bound)."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)).encode())
        except Exception:  # noqa: BLE001
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before
    except Exception:  # noqa: BLE001
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scheme", default="ecsim", choices=["ecsim", "ecsimcorr", "eccapfim"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short ECSIMCorr / EC-CAPFIM measurements (BASELINE configs 3, 5) appended to the ECSIM line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # libraries (NCCL's version banner, torch warnings) write to fd 1; the contract is ONE JSON line on
    # stdout, so everything else goes to stderr and the line is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import xpic_b200 as X

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def maxreduce(v):
        return reduce(v, dist.ReduceOp.MAX) if world > 1 else v

    def sumreduce(v):
        return reduce(v, dist.ReduceOp.SUM) if world > 1 else v

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    precond = env_int("XPIC_BENCH_PRECOND", 8)
    codes = {"ecsim": X.ECSIM, "ecsimcorr": X.ECSIMCORR, "eccapfim": X.ECCAPFIM}

    def make_sim(scheme_name):
        comm_id = None
        if world > 1:
            ids = [X.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            comm_id = ids[0]
        n, ppc = workload(world, scheme_name)
        sim = X.Simulation(n, d=(0.5, 0.5, 0.5), dt=1.5, scheme=codes[scheme_name], device=local_rank, rank=rank, nranks=world, comm_id=comm_id,
                           track_ids=False)
        total = n[0] * n[1] * n[2] * ppc
        sid = sim.add_species(q=-1.0, m=1.0, n=1.0, Np=ppc, capacity=int(sim.ncl * ppc * 1.25) + 65536)
        mine = sim.set_particles_maxwellian(sid, total, T=0.1, seed=20261018, tov=True)
        sim.solver_set(0, 1e-7, 1e-7, 100, 30, precond)  # the reference's tolerances (ecsim/simulation.h:15-18)
        sim.solver_set(1, 1e-7, 1e-7, 100, 30, precond)
        return sim, n, ppc, int(sumreduce(float(mine)))

    def total_energy(sim):
        # device reductions, summed over all ranks inside the library (Energy diagnostic, diagnostics/energy.cpp:43-107)
        return sim.field_energy("E") + sim.field_energy("B") + sim.scalar("kinetic")

    def conservation_checks(sim, nparticles, steps=3):
        """Untimed: `steps` more steps with the invariants the reference's tests assert (tests/common.h:30-89 compares
        the energy tables; dE + dB + dK at solver tolerance) evaluated around each of them."""
        e = [total_energy(sim)]
        reason_min = None
        for _ in range(steps):
            sim.step()
            e.append(total_energy(sim))
            r = sim.nonlinear_info()["reason"] if sim.scheme == X.ECCAPFIM else sim.solver_info(0)[2]
            reason_min = r if reason_min is None else min(reason_min, r)
        drift = max(abs(b - a) for a, b in zip(e[:-1], e[1:]))
        count = int(sumreduce(float(sim.particle_count(0))))
        return {"steps_checked": steps, "energy_total": e[-1], "energy_drift_max": drift, "energy_drift_max_rel": drift / e[0] if e[0] else None,
                "particles_conserved": count == nparticles, "solver_reason_min": reason_min,
                "note": "max over the checked steps of |d(wE + wB + wK)| (all ranks); the solver stops at rtol = atol = 1e-7 (reference defaults)"}

    # ================= the headline measurement: args.scheme ========================================================
    sim, n, ppc, nparticles = make_sim(args.scheme)
    scheme = codes[args.scheme]

    # ---- resident arm: W warm-up steps, then exactly K timed steps ---------------------------------
    sim.run_steps(args.warmup)
    fp64 = measure_fp64_peak() if rank == 0 else None  # before the timed region, while this rank's GPU is idle
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count()
    sim.timing_reset()
    sim.family_profile(True)
    if args.scheme == "eccapfim":
        sim.nonlinear_profile(1)
    barrier()
    ms = sim.run_steps(args.steps)
    barrier()
    cap = None
    if args.scheme == "eccapfim":
        passes, pass_ms = sim.nonlinear_profile(0)
        cap = dict(sim.nonlinear_info(), passes=passes, pass_ms=pass_ms / max(passes, 1))
    launches = sim.launch_count() - launches0
    families = sim.family_profile_read()
    spmv_n, spmv_ms = families["spmv"]
    sim.family_profile(False)
    stage_s = {k: v[0] / max(v[1], 1) for k, v in sim.timing().items()}
    its = sim.solver_info(0)[0]
    ms = maxreduce(ms)
    clocks = sampler.stop() if rank == 0 else None
    value = nparticles * args.steps / (ms * 1e-3)

    # ---- end-to-end arm: the same K steps through xb_step_host with pinned HOST buffers ------------
    # (allocated while the process sits on the CPUs next to its GPU, so that the copies do not cross sockets)
    affinity = bind_near_gpu(local_rank)
    E = torch.zeros(sim.nown, dtype=torch.float64).pin_memory()
    B = torch.zeros(sim.nown, dtype=torch.float64).pin_memory()
    B0 = torch.zeros(sim.nown, dtype=torch.float64).pin_memory()
    K = torch.zeros(1, dtype=torch.float64).pin_memory()
    E.numpy()[:] = sim.get_field("E")
    B.numpy()[:] = sim.get_field("B")
    barrier()
    ms_e2e = sim.run_steps_host(args.steps, E.numpy(), B.numpy(), B0.numpy(), K.numpy())
    barrier()
    ms_e2e = maxreduce(ms_e2e)
    e2e_value = nparticles * args.steps / (ms_e2e * 1e-3)
    h2d = 3 * sim.nown * 8 * world
    d2h = (2 * sim.nown * 8 + 8) * world
    del E, B, B0
    if affinity:
        os.sched_setaffinity(0, affinity)  # the CPU baseline below uses every host thread again

    checks = conservation_checks(sim, nparticles)

    # ---- roofline of the kernel the metric names (the operator SpMV inside GMRES) -----------------
    spmv_avg_ms = spmv_ms / max(spmv_n, 1)
    alg_bytes = 3000.0 * sim.ncl  # 369 coefficients * 8 B + 24 B x + 24 B y per cell
    achieved = alg_bytes / (spmv_avg_ms * 1e-3) / 1e9 if spmv_n else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tuple(tj.get("grid", [])) == tuple(n) and world == 1:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
    roofline = {"kernel": "k_spmv<L+M>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": peak_source,
                "launches_timed": spmv_n, "avg_launch_ms": spmv_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "spmv_share_of_step": (spmv_ms / ms) if ms else None}
    if cap is not None:
        # eccapfim: the dominant kernel is the particle pass of every residual evaluation (k_cap_push);
        # its algorithmic HBM bytes are 96 B per particle (read r, v of the start state, write the pushed
        # state) -- the kernel is bound by fp64 / shared-memory work, the HBM fraction says how far from a stream it is
        npart_rank = float(nparticles) / world
        pass_bytes = 96.0 * npart_rank
        ach = pass_bytes / (cap["pass_ms"] * 1e-3) / 1e9 if cap["pass_ms"] else None
        roofline = {"kernel": "k_cap_push (+ current halo add)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                    "traffic": None, "peak_source": peak_source,
                    "launches_timed": cap["passes"], "avg_launch_ms": cap["pass_ms"], "algorithmic_bytes_per_launch": pass_bytes,
                    "share_of_step": cap["passes"] * cap["pass_ms"] / ms if ms else None,
                    "particle_pushes_per_s": npart_rank * world * cap["passes"] / (ms * 1e-3) if ms else None}

    # ---- the kernel families INSIDE the timed steps (CUDA events around every launch group, rank 0's GPU) ---------
    kernels = None
    roofline_dominant = None
    if args.scheme == "ecsim":
        npart_rank = float(nparticles) / world
        fp64_peak = float(fp64["dmma_tflops"]) if fp64 else float(os.environ.get("XPIC_FP64_TFLOPS", "37.0"))
        fp64_src = ("tools/fp64_peak.cu run on this GPU before the timed region: mma.sync.m8n8k4.f64 issue rate (DFMA rate %.1f TFLOP/s)" % fp64["dfma_tflops"]
                    if fp64 else "vendor fp64 figure for B200 (xpic_b200/_build/fp64_peak missing)")
        per = {k: (v[1] / args.steps, v[0] // args.steps) for k, v in families.items()}  # ms per step, launch groups per step
        t_sort, t_dep, t_push, t_pre = per["sort"][0], per["moments"][0], per["second_push"][0], per["precond"][0]
        kernels = [
            {"name": "re-binning inside the step (move + key pass, scan, scatter" + (", migration" if world > 1 else "") + ")", "ms": t_sort, "bound": "hbm",
             "achieved": 152.0 * npart_rank / t_sort / 1e6, "unit": "GB/s", "algorithmic": "152 B/particle: read r,v + key, write r,v"},
            {"name": "moments (k_cell_moments_ws: field records + DMMA accumulator tiles; k_gather_tiles: rows)", "ms": t_dep, "bound": "fp64", "achieved": 1200.0 * npart_rank / t_dep / 1e9,
             "unit": "TFLOP/s", "algorithmic": "1200 flop/particle (576 + 24 FMA); HBM: 48 B/particle in, 15.9 KB/cell of accumulator tiles staged out and in, 3 KB/cell rows out"},
            {"name": "second push (tile-staged gather + Boris)", "ms": t_push, "bound": "hbm", "achieved": 72.0 * npart_rank / t_push / 1e6, "unit": "GB/s",
             "algorithmic": "72 B/particle: read r,v, write v"},
            {"name": "operator SpMV (all launches of a step)", "ms": spmv_ms / args.steps, "bound": "hbm", "achieved": achieved, "unit": "GB/s",
             "algorithmic": f"{spmv_n // args.steps} launches x 3000 B/cell"},
            {"name": f"Chebyshev(M) preconditioner, degree {precond} (all applications of a step)", "ms": t_pre, "bound": "hbm",
             "achieved": (per["precond"][1] * ((precond - 1) * 96.0 + 48.0) * sim.ncl / t_pre / 1e6) if t_pre else None, "unit": "GB/s",
             "algorithmic": f"{per['precond'][1]} applications x ({precond - 1} x 96 + 48) B/cell"},
        ]
        for k in kernels:
            if k["achieved"] is not None:
                k["frac"] = k["achieved"] / (peak if k["bound"] == "hbm" else fp64_peak)
        kernels[1]["peak"] = fp64_peak
        kernels[1]["peak_source"] = fp64_src
        stage_sum = {"first_push": t_sort + t_dep, "advance_fields_spmv_plus_precond": spmv_ms / args.steps + t_pre, "second_push": t_push,
                     "moments_parts": {"accumulator_tiles_owned_planes": per["moments_cells"][0], "boundary_plane_exchange_left_over": per["moments_ghost"][0],
                                       "row_gather": per["moments_rows"][0]},
                     "sort_parts": {"move_and_key_pass": per["sort_keys"][0], "migration_counts_payloads_arrivals": per["sort_migrate"][0],
                                    "scan_and_scatter": per["sort_scatter"][0]}}
        moments_traffic = None  # DRAM bytes of the family from the committed ncu capture (same grid, one GPU), else null
        mpath = os.path.join(ROOT, "profiles", "moments_traffic.json")
        if os.path.exists(mpath) and world == 1:
            try:
                with open(mpath) as f:
                    mj = json.load(f)
                if tuple(mj.get("grid", [])) == tuple(n) and mj.get("ppc") == ppc:
                    moments_traffic = mj.get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                pass
        # by time the moment deposition is the dominant kernel family of the step (the SpMV above is the kernel
        # BASELINE.json's metric names); its roof is the fp64 tensor / FMA rate, not HBM
        roofline_dominant = {"kernel": "moment deposition: k_cell_moments_ws (fp64 DMMA m8n8k4, fused field records, accumulator tiles) + k_gather_tiles", "bound": "tensor",
                             "achieved": kernels[1]["achieved"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": kernels[1]["achieved"] / fp64_peak,
                             "traffic": moments_traffic, "peak_source": fp64_src, "avg_launch_ms": t_dep,
                             "algorithmic_flops_per_launch": 1200.0 * npart_rank, "share_of_step": t_dep / (ms / args.steps) if ms else None,
                             "hbm_view": {"algorithmic_bytes": 48.0 * npart_rank + 2976.0 * sim.ncl, "achieved_GBs": (48.0 * npart_rank + 2976.0 * sim.ncl) / t_dep / 1e6,
                                          "frac_of_hbm_peak": (48.0 * npart_rank + 2976.0 * sim.ncl) / t_dep / 1e6 / peak},
                             "family_ms_sum_vs_stage_clock": stage_sum}
    sim.close()
    del sim
    torch.cuda.empty_cache()

    # ================= BASELINE configs 3 and 5, short runs on the same GPUs (driver-visible lines) ==================
    extra = None
    if args.scheme == "ecsim" and not args.no_extra and not os.environ.get("XPIC_BENCH_GRID"):
        extra = {}
        for name, (w_, k_) in (("ecsimcorr", (3, 5)), ("eccapfim", (3, 3))):
            try:
                sx, nx_, ppcx, npx = make_sim(name)
                sx.run_steps(w_)
                barrier()
                msx = maxreduce(sx.run_steps(k_))
                barrier()
                info = {"metric": f"{name}_particle_steps_per_s", "value": npx * k_ / (msx * 1e-3), "unit": UNIT, "ms_per_step": msx / k_, "steps": k_, "warmup": w_,
                        "n_gpus": world, "workload": f"{name.upper()} 3D {nx_[0]}x{nx_[1]}x{nx_[2]} cells x {ppcx} ppc fp64", "particles": npx,
                        "checks": conservation_checks(sx, npx, steps=2)}
                if name == "eccapfim":
                    info["nonlinear"] = sx.nonlinear_info()
                else:
                    info["krylov_iterations_per_step"] = [sx.solver_info(0)[0], sx.solver_info(1)[0]]
                extra[name] = info
                sx.close()
                del sx
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                extra[name] = {"error": repr(e)[:300]}
                barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess(args.scheme, steps=2 if args.scheme == "eccapfim" else 4, warm=1)
        line = {
            "metric": METRIC if args.scheme == "ecsim" else f"{args.scheme}_particle_steps_per_s", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.scheme.upper()} 3D {n[0]}x{n[1]}x{n[2]} cells x {ppc} ppc fp64, periodic Maxwellian plasma", "particles": nparticles,
                       "parallelism": f"z-slabs x{world}", "krylov": f"GMRES(30) rtol=atol=1e-7, Chebyshev(M) degree {precond} right preconditioner",
                       "krylov_iterations_per_step": its, "l2": "inputs (6.2 GB operator, 6.4 GB particles per GPU) exceed the 126 MB L2; no flush needed",
                       "stage_ms": {k: 1e3 * v for k, v in stage_s.items()},
                       **({"nonlinear": {"solver": "Anderson(10) on x - P F(x), P = Chebyshev(12) of ((1+sigma) I + dt^2/4 curl curl)^-1; atol = rtol = 1e-7 (reference)",
                                         "iterations_last_step": cap["iterations"], "residual_evaluations_last_step": cap["fevals"],
                                         "picard_iterations_per_particle_last_evaluation": cap["avg_cn"], "path_pieces_per_particle": cap["avg_cells"],
                                         "picard": "warm-started from the previous residual evaluation of the step (first evaluation cold: ~2.7 iterations per particle)",
                                         "reference_residual_evaluations_per_step": 105}} if cap else {})},
            "roofline": roofline, "roofline_dominant": roofline_dominant, "kernels": kernels, "fp64_peak": fp64, "checks": checks, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "boundary": "xb_step_host: E, B, B0 uploaded from pinned host buffers (B under the re-binning, E and B0 under the particle stages), E, B and kinetic energy downloaded every step; particles resident"},
            "other_configs": extra,
            "gpu_launches": launches, "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
