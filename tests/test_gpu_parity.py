"""GPU parity tests: every kernel family and the whole ECSIM / ECSIMCorr step, through the C ABI,
against the CPU oracle on the same seeded inputs.  Tolerances: fp64 round-off for single kernels,
the north-star bound 1e-8 relative for state after 10 steps (BASELINE.json)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import by_id, csr_to_stencil, csr_to_stencil_open, make_pair, rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def X():
    import xpic_b200

    return xpic_b200


def test_library_loaded_and_device_present(X):
    import torch

    assert torch.cuda.is_available()
    assert os.path.exists(X.library_path())
    s = X.Simulation((8, 8, 8))
    assert s.launch_count() == 0
    s.close()


def test_curl_matches_oracle(X):
    o, s = make_pair(n=(10, 8, 6), Np=1)
    f = np.random.default_rng(1).standard_normal(o.n3)
    for positive in (True, False):
        assert rel_err(s.curl(f, positive), o.curl(f, positive)) < 1e-14


def test_spmv_constant_operator_matches_oracle(X):
    o, s = make_pair(n=(10, 8, 6), Np=1)
    x = np.random.default_rng(2).standard_normal(o.n3)
    assert rel_err(s.spmv(x, op=2), o.spmv(x, L=False, M=True)) < 1e-13


@pytest.mark.parametrize("n", [(10, 10, 10), (33, 6, 5)])
def test_spmv_with_oracle_operator(X, n):
    o, s = make_pair(n=n, Np=8, seed_fields=3)
    o.deposit()
    s.operator_upload(csr_to_stencil(o, X.coef_table()))
    x = np.random.default_rng(4).standard_normal(o.n3)
    assert rel_err(s.spmv(x, op=1), o.spmv(x, L=True, M=False)) < 1e-13
    assert rel_err(s.spmv(x, op=3), o.spmv(x, L=True, M=True)) < 1e-13


@pytest.mark.parametrize("n,Np", [((10, 10, 10), 100), ((12, 9, 7), 13)])
def test_deposit_matches_oracle(X, n, Np):
    o, s = make_pair(n=n, Np=Np, seed_fields=5)
    o.deposit()
    s.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    got = s.operator_download()
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-13
    assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12
    # the slots are complete: L x agrees too (nothing of the oracle's pattern was dropped)
    x = np.random.default_rng(6).standard_normal(o.n3)
    assert rel_err(s.spmv(x, op=1), o.spmv(x, L=True, M=False)) < 1e-12


def test_deposit_scalar_and_tensor_core_kernels_agree(X):
    o, s = make_pair(n=(9, 8, 7), Np=37, seed_fields=15)
    o.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    for variant in (0, 4, 1):  # 0: DMMA variant tiles, 4: DMMA cell blocks folded in shared memory, 1: scalar FMA cell blocks
        s.set_option(0, variant)
        s.deposit()
        assert np.max(np.abs(s.operator_download() - ref)) / np.max(np.abs(ref)) < 1e-13, variant
        assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12, variant


def test_deposit_two_species(X):
    o, s = make_pair(n=(8, 8, 8), Np=20, seed_fields=7, species=((-1.0, 1.0, 1.0), (+1.0, 100.0, 1.0)))
    o.deposit()
    s.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    assert np.max(np.abs(s.operator_download() - ref)) / np.max(np.abs(ref)) < 1e-13
    assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12


def test_solve_matches_oracle(X):
    o, s = make_pair(n=(10, 10, 10), Np=20, seed_fields=8)
    o.deposit()
    s.deposit()
    b = np.random.default_rng(9).standard_normal(o.n3)
    x = s.solve(b, which=0, op=3)
    it, rn, reason = s.solver_info(0)
    assert reason > 0
    assert rel_err(o.spmv(x), b) < 1e-10
    # Chebyshev-preconditioned solve gives the same answer in fewer iterations
    s.solver_set(0, 1e-12, 1e-50, 1000, 30, 6)
    x2 = s.solve(b, which=0, op=3)
    it2, _, reason2 = s.solver_info(0)
    assert reason2 > 0 and it2 < it
    assert rel_err(x2, x) < 1e-9


def _compare_state(o, s, tol):
    for name in ("E", "B"):
        assert rel_err(s.get_field(name), o.get_field(name)) < tol, name
    po, io = by_id(*o.get_particles(0))
    pg, ig = by_id(*s.get_particles(0))
    assert np.array_equal(io, ig)
    assert rel_err(pg[:, :3], po[:, :3]) < tol
    assert rel_err(pg[:, 3:], po[:, 3:]) < tol


def test_ecsim_10_steps_state_parity(X):
    o, s = make_pair(n=(10, 10, 10), Np=100, scheme=X.ECSIM)
    for _ in range(10):
        o.step(O.ECSIM)
        s.step()
    _compare_state(o, s, 1e-8)


def test_ecsimcorr_10_steps_state_parity(X):
    o, s = make_pair(n=(10, 10, 10), Np=100, scheme=X.ECSIMCORR)
    for _ in range(10):
        o.step(O.ECSIMCORR)
        s.step()
    _compare_state(o, s, 1e-8)
    for name in ("pred_w", "corr_w", "lambda_dK"):
        assert abs(s.scalar(name) - o.scalar(name)) < 1e-10, name


def test_ecsimcorr_esirkepov_variants_agree(X):
    """Atomic-free tensor-core Esirkepov deposit vs the per-particle reduction kernel vs the oracle."""
    res = {}
    for variant in (0, 1):
        o, s = make_pair(n=(9, 7, 6), Np=30, scheme=X.ECSIMCORR, seed_fields=21)
        s.set_option(1, variant)
        for _ in range(3):
            s.step()
        res[variant] = (s.get_field("currJe"), s.get_field("E"), s.scalar("pred_w"))
        if variant == 0:
            for _ in range(3):
                o.step(O.ECSIMCORR)
            assert rel_err(res[0][0], o.get_field("currJe")) < 1e-10
            assert rel_err(res[0][1], o.get_field("E")) < 1e-10
            assert abs(res[0][2] - o.scalar("pred_w")) < 1e-12
    assert rel_err(res[0][0], res[1][0]) < 1e-11
    assert rel_err(res[0][1], res[1][1]) < 1e-11


def test_ecsim_golden_energy_rows(X):
    """The CUDA path against the reference's own golden file (curl_sign = -1)."""
    _, s = make_pair(n=(10, 10, 10), Np=100, scheme=X.ECSIM, curl_sign=-1)
    _, gold = O.read_table(os.path.join(GOLDEN, "ecsim_ex1", "energy.txt"))
    rows = [(0.0, 0.0, s.scalar("kinetic"))]
    for _ in range(12):
        s.step()
        wE, wB = s.field_energies()
        rows.append((wE, wB, s.scalar("kinetic")))
    rows = np.array(rows)
    np.testing.assert_allclose(rows, gold[:13, 1:4], rtol=2e-6, atol=1e-10)
    assert np.max(np.abs(np.diff(rows.sum(axis=1)))) < 1e-12  # dE + dB + dK at solver tolerance


def test_ecsimcorr_golden_energy_rows(X):
    _, s = make_pair(n=(10, 10, 10), Np=100, scheme=X.ECSIMCORR, curl_sign=-1)
    _, gold = O.read_table(os.path.join(GOLDEN, "ecsimcorr_ex1", "energy.txt"))
    _, gc = O.read_table(os.path.join(GOLDEN, "ecsimcorr_ex1", "energy_conservation.txt"))
    rows = [(0.0, 0.0, s.scalar("kinetic"))]
    cwd = []
    for _ in range(8):
        s.step()
        wE, wB = s.field_energies()
        rows.append((wE, wB, s.scalar("kinetic")))
        cwd.append(s.scalar("lambda_dK"))
    np.testing.assert_allclose(np.array(rows), gold[:9, 1:4], rtol=2e-6, atol=1e-10)
    np.testing.assert_allclose(np.array(cwd), gc[1:9, 4], rtol=5e-6)


def test_energy_conservation_32cubed(X):
    """Size-independent property at BASELINE config 1 size (32^3 x 100 ppc): |dE + dB + dK| at
    solver tolerance, with the default (reference) tolerances and the production preconditioner."""
    n = (32, 32, 32)
    s = X.Simulation(n, scheme=X.ECSIM, track_ids=False)
    sid = s.add_species(Np=100)
    rng = np.random.default_rng(11)
    N = 100 * 32**3
    pts = np.empty((N, 6))
    pts[:, :3] = rng.random((N, 3)) * 16.0
    pts[:, 3:] = rng.standard_normal((N, 3)) * np.sqrt(0.1 / 511.0)
    assert s.add_particles(sid, pts) == N
    s.solver_set(0, 1e-7, 1e-7, 100, 30, 6)
    tot = []
    for _ in range(4):
        wE, wB = s.field_energies()
        tot.append(wE + wB + s.scalar("kinetic"))
        s.step()
    wE, wB = s.field_energies()
    tot.append(wE + wB + s.scalar("kinetic"))
    assert np.max(np.abs(np.diff(tot))) < 1e-9 * tot[0] + 1e-9
    assert s.launch_count() > 0


def test_host_program_reproduces_golden_tables(X, tmp_path):
    """The C++ host mirror (config.json in, temporal/*.txt out) against the reference's golden
    tables: same file format (tests/common.h:30-89 compares column by column), same numbers."""
    import json
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "xpic_b200", "_build", "xpic_b200.out")
    if not os.path.exists(exe):
        pytest.skip("host program not built")
    for name, opts in (("ecsim_ex1", ["-ksp_rtol", "1e-11", "-ksp_atol", "1e-50"]),
                       ("ecsimcorr_ex1", ["-predict_ksp_rtol", "1e-11", "-predict_ksp_atol", "1e-50", "-correct_ksp_rtol", "1e-11", "-correct_ksp_atol", "1e-50"])):
        cfg = json.load(open(os.path.join(ROOT, "configs", name + ".json")))
        cfg["OutputDirectory"] = str(tmp_path / name)
        path = tmp_path / (name + ".json")
        path.write_text(json.dumps(cfg))
        subprocess.run([exe, str(path), "-curl_sign", "-1", "-ksp_max_it", "500"] + opts, check=True, capture_output=True, timeout=300)
        for table in ("energy.txt", "energy_conservation.txt"):
            tg, gold = O.read_table(os.path.join(GOLDEN, name, table))
            to, out = O.read_table(str(tmp_path / name / "temporal" / table))
            assert to == tg  # identical headers (titles are truncated to the column width)
            rows = out.shape[0]
            assert rows == 11
            if table == "energy.txt":
                np.testing.assert_allclose(out[:, 1:4], gold[:rows, 1:4], rtol=2e-6, atol=1e-10)
            else:
                np.testing.assert_allclose(out[:, 1:4], gold[:rows, 1:4], rtol=1e-4, atol=2e-9)
                assert np.max(np.abs(out[:, -2 if name == "ecsimcorr_ex1" else -1])) < 1e-11


def test_runs_are_bit_reproducible_without_ids(X):
    """Atomic-free deposits + canonical order inside the bins: two runs give identical bits."""
    out = []
    for _ in range(2):
        s = X.Simulation((12, 10, 8), scheme=X.ECSIMCORR, track_ids=False)
        s.set_option(2, 1)  # canonical order inside the bins (default only when ids are tracked)
        sid = s.add_species(Np=40)
        s.set_particles_maxwellian(sid, 12 * 10 * 8 * 40, T=0.1, seed=7)
        for _ in range(4):
            s.step()
        out.append((s.get_field("E").copy(), s.get_field("B").copy(), s.scalar("kinetic")))
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


def _charge_density(pts, n, d, q_np):
    """ParticlesChargeDensity::collect (src/diagnostics/charge_conservation.cpp:33-97): 2nd-order spline,
    3 nodes per axis from ceil(p - 1.5), periodic fold."""
    nx, ny, nz = n
    rho = np.zeros((nz, ny, nx))

    def spline2(s):
        s = np.abs(s)
        return np.where(s <= 0.5, 0.75 - s * s, np.where(s < 1.5, 0.5 * (1.5 - s) ** 2, 0.0))

    p = pts[:, :3] / np.array(d)
    start = np.ceil(p - 1.5).astype(int)
    for k in range(3):
        wz = spline2(p[:, 2] - (start[:, 2] + k))
        for j in range(3):
            wy = spline2(p[:, 1] - (start[:, 1] + j))
            for i in range(3):
                wx = spline2(p[:, 0] - (start[:, 0] + i))
                np.add.at(rho, ((start[:, 2] + k) % nz, (start[:, 1] + j) % ny, (start[:, 0] + i) % nx), q_np * wx * wy * wz)
    return rho


def test_ecsimcorr_charge_conservation_property(X):
    """d(rho)/dt + div J = 0 to round-off for the Esirkepov current (the reference's
    charge_conservation.txt golden holds 2e-13 .. 9e-13 in the 1-norm for 10^3 x 100 ppc)."""
    n, d, dt, Np = (10, 10, 10), (0.5, 0.5, 0.5), 1.5, 100
    _, s = make_pair(n=n, Np=Np, scheme=X.ECSIMCORR)
    q_np = -1.0 * 1.0 / Np
    worst = 0.0
    for _ in range(3):
        rho0 = _charge_density(s.get_particles(0)[0], n, d, q_np)
        s.step()
        rho1 = _charge_density(s.get_particles(0)[0], n, d, q_np)
        J = s.get_field("currJe_sort").reshape(n[2], n[1], n[0], 3)
        div = ((J[..., 0] - np.roll(J[..., 0], 1, axis=2)) / d[0] + (J[..., 1] - np.roll(J[..., 1], 1, axis=1)) / d[1] +
               (J[..., 2] - np.roll(J[..., 2], 1, axis=0)) / d[2])  # Divergence::create_negative, utils/operators.cpp:306-318
        res = (rho1 - rho0) / dt + div
        worst = max(worst, float(np.abs(res).sum()))
    assert worst < 5e-12


def test_empty_and_ragged_inputs(X):
    # no particles at all: the step is a vacuum Maxwell step with zero fields; nothing diverges
    s = X.Simulation((8, 6, 5), scheme=X.ECSIM)
    s.add_species(Np=10)
    s.step()
    assert np.all(s.get_field("E") == 0.0) and s.solver_info(0)[2] > 0
    s.close()
    # vacuum with a seeded field: pure field update, matches the oracle
    o, s = make_pair(n=(7, 6, 9), Np=1, seed_fields=31)
    # make_pair added 7*6*9 particles; compare 2 steps anyway (ragged grid: nothing divides 4 or 32)
    for _ in range(2):
        o.step(O.ECSIM)
        s.step()
    assert rel_err(s.get_field("E"), o.get_field("E")) < 1e-10
    # all particles in one cell (several 32-particle chunks, every other cell empty), crossing the periodic boundary
    import xpic_b200

    n = (8, 8, 8)
    o = O.Oracle(n)
    sid = o.add_species(Np=50)
    rng = np.random.default_rng(3)
    pts = np.empty((150, 6))
    pts[:, :3] = np.array([0.02, 3.97, 0.01]) + rng.random((150, 3)) * 0.03  # hugging x = 0, y = Ly, z = 0
    pts[:, 3:] = rng.standard_normal((150, 3)) * 0.05
    pts[:, 3] -= 0.05  # drift across x = 0
    o.set_particles(sid, pts)
    s = xpic_b200.Simulation(n, scheme=xpic_b200.ECSIM, track_ids=True)
    s.add_species(Np=50)
    assert s.add_particles(0, pts, np.arange(150, dtype=np.uint64)) == 150
    for sim in (o, s):
        sim.solver_set(0, 1e-12, 1e-50, 500, 30)
    for _ in range(4):
        o.step(O.ECSIM)
        s.step()
    po, io = by_id(*o.get_particles(0))
    pg, ig = by_id(*s.get_particles(0))
    assert rel_err(pg, po) < 1e-9 and rel_err(s.get_field("E"), o.get_field("E")) < 1e-9


def test_energy_conservation_64cubed_default_tolerances(X):
    """Size-independent property at a larger size with the production solver settings."""
    s = X.Simulation((64, 64, 64), scheme=X.ECSIM, track_ids=False)
    sid = s.add_species(Np=64)
    s.set_particles_maxwellian(sid, 64**4, T=0.1, seed=5)
    s.solver_set(0, 1e-7, 1e-7, 100, 30, 6)
    tot = []
    for _ in range(4):
        wE, wB = s.field_energies()
        tot.append(wE + wB + s.scalar("kinetic"))
        s.step()
    wE, wB = s.field_energies()
    tot.append(wE + wB + s.scalar("kinetic"))
    assert np.max(np.abs(np.diff(tot))) < 1e-9 * tot[0]
    assert s.solver_info(0)[0] <= 15


@pytest.mark.parametrize("name,scheme", [("ecsim_ex1", 0), ("ecsimcorr_ex1", 1)])
def test_golden_field_dumps_and_all_rows_100_steps(X, name, scheme):
    """The reference's whole test run (100 steps, tests/ecsim/ecsim_ex1.cpp:32-71) on the GPU: every row
    of energy.txt, and the E / B dumps at t = 50 and t = 100 (float32 files)."""
    _, s = make_pair(n=(10, 10, 10), Np=100, scheme=scheme, curl_sign=-1, rtol=1e-10)
    _, gold = O.read_table(os.path.join(GOLDEN, name, "energy.txt"))
    rows = [(0.0, 0.0, s.scalar("kinetic"))]
    g0 = np.fromfile(os.path.join(GOLDEN, name, "density_000.f32"), dtype=np.float32).astype(np.float64)
    assert np.max(np.abs(s.density(0) - g0)) < 2e-7 * np.max(g0)
    for t in range(1, 101):
        s.step()
        wE, wB = s.field_energies()
        rows.append((wE, wB, s.scalar("kinetic")))
        if t in (50, 100):
            for f in ("E", "B"):
                g = np.fromfile(os.path.join(GOLDEN, name, f"{f}_{t:03d}.f32"), dtype=np.float32).astype(np.float64)
                assert rel_err(s.get_field(f), g) < 2e-5, (f, t)
            # DistributionMoment density dump (electrons/density/050, 100): the particle positions of the run
            g = np.fromfile(os.path.join(GOLDEN, name, f"density_{t:03d}.f32"), dtype=np.float32).astype(np.float64)
            assert np.max(np.abs(s.density(0) - g)) < 1e-4 * np.max(g), t
    # the reference solved to 1e-7 and chaos amplifies the difference slowly: 7 digits early, 5 at the end
    np.testing.assert_allclose(np.array(rows)[:31], gold[:31, 1:4], rtol=3e-6, atol=1e-10)
    np.testing.assert_allclose(np.array(rows), gold[:, 1:4], rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_step_host_equals_resident_step(X, scheme):
    """xb_step_host (host E, B, B0 in; E, B, kinetic energies out; the copies of E and B0 overlap the
    particle stages) gives the same state as xb_step on resident fields."""
    sims = []
    o = O.Oracle((9, 8, 7))
    sid = o.add_species(Np=12)
    o.set_particles_maxwell(sid, 0.1, True)
    pts, ids = o.get_particles(sid)
    rng = np.random.default_rng(77)
    E0, B0 = 0.01 * rng.standard_normal(o.n3), 0.03 * rng.standard_normal(o.n3)
    for _ in range(2):
        s = X.Simulation((9, 8, 7), scheme=scheme, track_ids=True)
        s.add_species(Np=12)
        s.add_particles(0, pts, ids)
        s.solver_set(0, 1e-12, 1e-50, 500, 30, 4)
        s.solver_set(1, 1e-12, 1e-50, 500, 30, 4)
        s.nonlinear_set(atol=1e-13, rtol=1e-30, particle_tol=1e-14)
        sims.append(s)
    a, b = sims
    a.set_field("E", E0)
    a.set_field("B", B0)
    E, B, Bz = E0.copy(), B0.copy(), np.zeros(o.n3)
    K = np.zeros(1)
    for _ in range(3):
        a.step()
        b.step_host(E, B, Bz, K)
    pa, ia = by_id(*a.get_particles())
    pb, ib = by_id(*b.get_particles())
    assert np.array_equal(ia, ib) and np.array_equal(b.get_field("E"), E)
    # bit-identical for all three schemes: eccapfim's current is accumulated in fixed point (integer additions commute),
    # the other deposits are atomic-free
    assert np.array_equal(a.get_field("E"), E) and np.array_equal(a.get_field("B"), B)
    assert K[0] == a.scalar("kinetic") and np.array_equal(pa, pb)


def test_device_charge_density_matches_host_restatement(X):
    n, d = (9, 8, 7), (0.5, 0.5, 0.5)
    _, s = make_pair(n=n, Np=23, scheme=X.ECSIMCORR)
    rho = s.charge_density(0).reshape(n[2], n[1], n[0])
    ref = _charge_density(s.get_particles(0)[0], n, d, -1.0 / 23)
    assert np.max(np.abs(rho - ref)) < 1e-13 * np.max(np.abs(ref))
    assert abs(rho.sum() - (-1.0 / 23) * s.particle_count(0)) < 1e-10  # the form factor sums to one


@pytest.mark.parametrize("scheme,current", [(1, "currJe"), (2, "J")])
def test_device_charge_conservation_diagnostic(X, scheme, current):
    """ChargeConservation on the device for the two charge-conserving schemes: the golden tables
    (tests/{ecsimcorr,eccapfim}/expected/*/temporal/charge_conservation.txt) hold 1e-13 .. 9e-13 in the
    1-norm and ~3e-14 in the 2-norm for this set-up; the residual is round-off of sums of O(1e-2) numbers."""
    n, Np = (10, 10, 10), 100
    _, s = make_pair(n=n, Np=Np, scheme=scheme)
    s.charge_density(0)  # ChargeConservation::initialize
    gold = O.read_table(os.path.join(GOLDEN, "ecsimcorr_ex1" if scheme == 1 else "eccapfim_ex1", "charge_conservation.txt"))[1]
    for t in range(1, 4):
        s.step()
        norms = s.charge_conservation(current)
        assert norms.shape == (2, 2)
        assert np.all(norms[:, 0] < 5e-12) and np.all(norms[:, 1] < 5e-13), norms
        # same order of magnitude as the reference's own round-off
        assert norms[0, 0] < 10 * gold[t, 1] and norms[0, 1] < 10 * gold[t, 2]
    # the diagnostic does see a violation: ecsim's implicit current is not charge conserving
    if scheme == 1:
        _, e = make_pair(n=n, Np=Np, scheme=X.ECSIM)
        e.charge_density(0)
        e.step()
        assert e.charge_conservation("currJe")[0, 0] > 1e-3


def test_host_backup_and_restart_round_trip(X, tmp_path):
    """SimulationBackup in the host mirror: PETSc binary Vec images of E, B, B0 (big-endian, class id 1211214),
    raw big-endian particles; a run restarted from the backup of step 5 continues the tables and ends in the
    state of the uninterrupted run (src/diagnostics/simulation_backup.cpp:27-183)."""
    import json
    import shutil
    import struct
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "xpic_b200", "_build", "xpic_b200.out")
    if not os.path.exists(exe):
        pytest.skip("host program not built")
    cfg = json.load(open(os.path.join(ROOT, "configs", "ecsimcorr_ex1.json")))
    cfg["SimulationBackup"] = {"diagnose_period": 7.5}
    a, b = tmp_path / "a", tmp_path / "b"
    cfg["OutputDirectory"] = str(a)
    (tmp_path / "a.json").write_text(json.dumps(cfg))
    subprocess.run([exe, str(tmp_path / "a.json")], check=True, capture_output=True, timeout=300)
    assert sorted(os.listdir(a / "simulation_backup")) == ["10", "5"]  # only the last two periods are kept
    raw = (a / "simulation_backup" / "5" / "E").read_bytes()
    assert len(raw) == 8 + 3000 * 8 and struct.unpack(">ii", raw[:8]) == (1211214, 3000)
    assert struct.unpack(">i", (a / "simulation_backup" / "5" / "electrons.numparts").read_bytes()) == (100000,)
    assert (a / "simulation_backup" / "5" / "electrons").stat().st_size == 100000 * 48
    # restart from step 5 in a fresh output directory that only holds that backup
    (b / "simulation_backup").mkdir(parents=True)
    shutil.copytree(a / "simulation_backup" / "5", b / "simulation_backup" / "5")
    cfg["OutputDirectory"] = str(b)
    cfg["SimulationBackup"] = {"diagnose_period": 7.5, "load_from": 5}
    (tmp_path / "b.json").write_text(json.dumps(cfg))
    res = subprocess.run([exe, str(tmp_path / "b.json")], check=True, capture_output=True, timeout=300, text=True)
    assert "successfully loaded" in res.stdout
    for table in ("energy.txt", "energy_conservation.txt", "charge_conservation.txt", "momentum_conservation.txt"):
        ta, ra = O.read_table(str(a / "temporal" / table))
        tb, rb = O.read_table(str(b / "temporal" / table))
        assert ta == tb and ra.shape == rb.shape, table
        np.testing.assert_allclose(rb, ra, rtol=1e-6, atol=1e-11, err_msg=table)
    for f in ("E", "B"):
        fa = np.fromfile(a / f / "10", dtype=np.float32)
        fb = np.fromfile(b / f / "10", dtype=np.float32)
        assert rel_err(fb.astype(np.float64), fa.astype(np.float64)) < 1e-6


# ---- round 2: sizes and code paths the small boxes above never reach ---------------
def _sorted_rows(pts):
    return pts[np.lexsort(pts.T[::-1])]


@pytest.mark.parametrize("variant", [0, 4, 3, 2])
def test_deposit_variants_at_tiling_size(X, variant):
    """40 x 20 x 12 cells: wider than the SpMV tile (32), the push tiles (16) and not a multiple of either, so
    CTAs own full tiles and partial ones, x-neighbour tiles exist.  Every moment kernel against the oracle's CSR."""
    o, s = make_pair(n=(40, 20, 12), Np=16, seed_fields=31)
    O.set_threads(O.max_threads())
    try:
        o.deposit()
    finally:
        O.set_threads(1)
    s.set_option(0, variant)
    s.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    assert np.max(np.abs(s.operator_download() - ref)) / np.max(np.abs(ref)) < 1e-13
    assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12
    x = np.random.default_rng(6).standard_normal(o.n3)
    assert rel_err(s.spmv(x, op=3), o.spmv(x, L=True, M=True)) < 1e-12


def test_ecsim_state_parity_at_tiling_size_and_production_path(X):
    """3 ECSIM steps at 40 x 20 x 12 x 16 ppc against the oracle (fields and id-matched particles < 1e-8), and the
    production configuration (track_ids = False, reference tolerances replaced by the same tight ones, Chebyshev
    preconditioner) against the id-tracked run on the same particles."""
    n = (40, 20, 12)
    o, s = make_pair(n=n, Np=16, scheme=X.ECSIM, precond=6)
    pts0, _ = o.get_particles(0)
    O.set_threads(O.max_threads())
    try:
        for _ in range(3):
            o.step(O.ECSIM)
            s.step()
    finally:
        O.set_threads(1)
    _compare_state(o, s, 1e-8)
    p = X.Simulation(n, scheme=X.ECSIM, track_ids=False)
    p.add_species(Np=16)
    assert p.add_particles(0, pts0) == len(pts0)
    p.solver_set(0, 1e-12, 1e-50, 1000, 30, 6)
    for _ in range(3):
        p.step()
    for name in ("E", "B"):
        assert rel_err(p.get_field(name), s.get_field(name)) < 1e-10, name
    a, b = _sorted_rows(p.get_particles(0)[0]), _sorted_rows(s.get_particles(0)[0])
    assert rel_err(a, b) < 1e-10
    p.close()


@pytest.mark.parametrize("scheme", ["ecsim", "ecsimcorr"])
def test_state_parity_with_cell_sizes_that_are_not_powers_of_two(X, scheme):
    """d = (0.3, 0.45, 0.7): r / d is a true division (gather.cuh to_cells), the other arithmetic path of every
    weight computation; 3 steps, fields and id-matched particles within 1e-8."""
    gs, os_ = (X.ECSIM, O.ECSIM) if scheme == "ecsim" else (X.ECSIMCORR, O.ECSIMCORR)
    o, s = make_pair(n=(11, 9, 7), Np=30, scheme=gs, d=(0.3, 0.45, 0.7), dt=0.8)
    for _ in range(3):
        o.step(os_)
        s.step()
    _compare_state(o, s, 1e-8)


def test_deposit_with_cell_sizes_that_are_not_powers_of_two(X):
    o, s = make_pair(n=(11, 9, 7), Np=30, seed_fields=41, d=(0.3, 0.45, 0.7), dt=0.8)
    o.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    for variant in (0, 4, 2, 1):
        s.set_option(0, variant)
        s.deposit()
        assert np.max(np.abs(s.operator_download() - ref)) / np.max(np.abs(ref)) < 1e-13, variant
        assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12, variant


def test_deposit_many_particles_per_cell_and_empty_cells(X):
    """Ragged bins: 300 particles per cell in a quarter of the box (several record rounds per cell, octants larger
    than a round), nothing elsewhere (empty cells, empty octants)."""
    n = (8, 8, 8)
    o = O.Oracle(n)
    s = X.Simulation(n, track_ids=True)
    sid = o.add_species(Np=300)
    s.add_species(Np=300)
    rng = np.random.default_rng(3)
    N = 300 * 4 * 4 * 8
    pts = np.empty((N, 6))
    pts[:, 0] = rng.random(N) * 2.0
    pts[:, 1] = rng.random(N) * 2.0 + 1.0
    pts[:, 2] = rng.random(N) * 4.0
    pts[:, 3:] = rng.standard_normal((N, 3)) * 0.02
    assert o.set_particles(sid, pts) == N
    pts, ids = o.get_particles(sid)
    assert s.add_particles(0, pts, ids) == N
    f = 0.05 * np.random.default_rng(4).standard_normal(o.n3)
    o.set_field("B", f)
    s.set_field("B", f)
    o.deposit()
    ref = csr_to_stencil(o, X.coef_table())
    for variant in (0, 4, 3, 2):
        s.set_option(0, variant)
        s.deposit()
        assert np.max(np.abs(s.operator_download() - ref)) / np.max(np.abs(ref)) < 1e-13, variant
        assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12, variant


def test_decomposition_independence_on_two_gpus(X):
    """The z-slab run equals the single-GPU run (tests/ecsim/CMakeLists.txt:14-16 demands the same of the reference's
    1- and 2-rank ctest runs).  Needs two GPUs: skipped on a single-GPU box."""
    import subprocess
    import sys

    import torch

    from conftest import ROOT

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, XPIC_CHECK_STEPS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-> OK") == 5


def test_open_z_boundary_state_parity(X):
    """da_boundary_z = DM_BOUNDARY_NONE / GHOSTED: no nodes outside the box (matrix entries and deposits dropped,
    zero ghost values), particles that leave through z are removed (src/interfaces/particles.cpp:100-103,
    src/utils/operators.cpp:12-43).  Operators, deposit and 10 ECSIM steps against the oracle with the same switch."""
    n = (9, 8, 10)
    o = O.Oracle(n, open_z=True)
    s = X.Simulation(n, scheme=X.ECSIM, track_ids=True, open_z=True)
    sid = o.add_species(Np=30)
    o.set_particles_maxwell(sid, T=2.0, tov=True)  # hot enough that particles reach the faces within ten steps
    s.add_species(Np=30)
    pts, ids = o.get_particles(sid)
    assert s.add_particles(0, pts, ids) == len(ids)
    for which in (0, 1):
        o.solver_set(which, 1e-12, 1e-50, 1000, 30)
        s.solver_set(which, 1e-12, 1e-50, 1000, 30, 0)
    rng = np.random.default_rng(17)
    for name, amp in (("E", 0.02), ("B", 0.05)):
        f = amp * rng.standard_normal(o.n3)
        o.set_field(name, f)
        s.set_field(name, f)
    x = rng.standard_normal(o.n3)
    for positive in (True, False):
        assert rel_err(s.curl(x, positive), o.curl(x, positive)) < 1e-14
    assert rel_err(s.spmv(x, op=2), o.spmv(x, L=False, M=True)) < 1e-13
    o.deposit()
    s.deposit()
    ref, exists = csr_to_stencil_open(o, X.coef_table())
    assert np.max(np.abs(np.where(exists, s.operator_download(), 0.0) - ref)) / np.max(np.abs(ref)) < 1e-13
    assert rel_err(s.get_field("currI"), o.get_field("currI")) < 1e-12
    assert rel_err(s.spmv(x, op=3), o.spmv(x, L=True, M=True)) < 1e-12
    n0 = o.particle_count()
    for _ in range(10):
        o.step(O.ECSIM)
        s.step()
    assert o.particle_count() < n0  # the test does exercise the removal
    assert s.particle_count() == o.particle_count()
    _compare_state(o, s, 1e-8)


def test_host_program_on_two_gpus(X, tmp_path):
    """The C++ host program as one process per GPU (`-rank r -nranks 2`, the reference's `mpiexec -n 2 ... -da_processors_z 2`,
    tests/ecsim/CMakeLists.txt:14-16): the energy tables of the 2-rank run equal the reference's golden tables, the
    field dump is one file written slab-wise by both ranks."""
    import json
    import subprocess

    import torch

    from conftest import ROOT

    exe = os.path.join(ROOT, "xpic_b200", "_build", "xpic_b200.out")
    if torch.cuda.device_count() < 2 or not os.path.exists(exe):
        pytest.skip("needs two GPUs and the host program")
    cfg = json.load(open(os.path.join(ROOT, "configs", "ecsim_ex1.json")))
    cfg["OutputDirectory"] = str(tmp_path / "two")
    cfg["mpi"] = {"da_processors_z": 2}
    cfg.setdefault("Diagnostics", []).append({"diagnostic": "LogView", "level": "EachTimestep"})
    path = tmp_path / "two.json"
    path.write_text(json.dumps(cfg))
    opts = ["-curl_sign", "-1", "-ksp_max_it", "500", "-ksp_rtol", "1e-11", "-ksp_atol", "1e-50"]
    procs = [subprocess.Popen([exe, str(path), "-rank", str(r), "-nranks", "2", "-device", str(r)] + opts, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    for table in ("energy.txt", "energy_conservation.txt"):
        tg, gold = O.read_table(os.path.join(GOLDEN, "ecsim_ex1", table))
        to, out = O.read_table(str(tmp_path / "two" / "temporal" / table))
        assert to == tg and out.shape[0] == 11
        if table == "energy.txt":
            np.testing.assert_allclose(out[:, 1:4], gold[:11, 1:4], rtol=2e-6, atol=1e-10)
        else:
            assert np.max(np.abs(out[:, -1])) < 1e-11
    assert os.path.exists(str(tmp_path / "two" / "E" / "10")), os.listdir(str(tmp_path / "two"))
    dump = np.fromfile(str(tmp_path / "two" / "E" / "10"), dtype=np.float32)
    gold = np.fromfile(os.path.join(GOLDEN, "ecsim_ex1", "E_010.f32"), dtype=np.float32) if os.path.exists(os.path.join(GOLDEN, "ecsim_ex1", "E_010.f32")) else None
    assert dump.size == 3 * 10 * 10 * 10
    if gold is not None:
        np.testing.assert_allclose(dump, gold, rtol=2e-4, atol=2e-7)
    log = open(str(tmp_path / "two" / "log-EachTimestep.txt")).read().splitlines()
    assert log[0].split()[:3] == ["Timestep", "Total_[sec]", "Main_Stage"] and len(log) == 11


@pytest.mark.parametrize("name", ["density", "current", "momentum_flux", "momentum_flux_cyl", "momentum_flux_diag", "momentum_flux_diag_cyl"])
def test_distribution_moments_and_regions(X, name):
    """Every moment of DistributionMoment (src/diagnostics/distribution_moment.cpp:212-313) on the device, whole box and a
    region that does not span x and z, against the formulas evaluated with numpy on the same particles."""
    from helpers import cell_moment

    n = (9, 8, 7)
    o, s = make_pair(n=n, Np=12, T=5.0)
    pts, _ = o.get_particles(0)
    for start, size in ((None, None), ((2, 0, 1), (5, 8, 4))):
        got = s.distribution_moment(name, 0, start, size).reshape(n[2], n[1], n[0], -1)
        ref = cell_moment(pts, n, (0.5, 0.5, 0.5), 1.0 / 12, name, start=start or (0, 0, 0), size=size)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 1e-12 * max(np.max(np.abs(ref)), 1e-30), (name, start)


@pytest.mark.parametrize("projector", ["vx_vy", "vz_vxy", "vr_vphi"])
def test_velocity_distribution(X, projector):
    """VelocityDistribution::collect (src/diagnostics/velocity_distribution.cpp:116-166) on the device, for a box and a
    cylinder, against the same rule evaluated with numpy: a particle counts when the centre of its cell is inside the
    geometry, its projected velocity is binned with ROUND_STEP, both axes use the x range."""
    n = (9, 8, 7)
    d = 0.5
    o, s = make_pair(n=n, Np=12, T=40.0)
    pts, _ = o.get_particles(0)
    r, v = pts[:, :3], pts[:, 3:]
    dv, vmin, vmax = (0.05, 0.04), (-0.6, -0.2), (0.7, 0.9)
    start, size = int(np.round(vmin[0] / dv[0])), int(np.round((vmax[0] - vmin[0]) / dv[0]))
    for geometry, p in (("box", (0.6, 0.0, 1.1, 3.9, 3.2, 3.0)), ("cylinder", (2.25, 2.0, 1.75, 1.6, 2.2, 0.0))):
        cell = np.floor(r / d).astype(int)
        centre = (cell + 0.5) * d
        if geometry == "box":
            lo, hi = np.array(p[:3]), np.array(p[3:])
            a0, a1 = np.floor(lo / d).astype(int), np.floor(hi / d).astype(int)
            inside = np.all((lo <= centre) & (centre < hi), axis=1)
        else:
            c, rad, h = np.array(p[:3]), p[3], p[4]
            lo, hi = c - np.array([rad, rad, 0.5 * h]), c + np.array([rad, rad, 0.5 * h])
            a0, a1 = np.floor(lo / d).astype(int), np.floor(hi / d).astype(int)
            inside = (np.abs(centre[:, 2] - c[2]) < 0.5 * h) & ((centre[:, 0] - c[0]) ** 2 + (centre[:, 1] - c[1]) ** 2 <= rad * rad)
        inside &= np.all((cell >= a0) & (cell < a1), axis=1)
        if projector == "vx_vy":
            pa, pb = v[:, 0], v[:, 1]
        elif projector == "vz_vxy":
            pa, pb = v[:, 2], np.hypot(v[:, 0], v[:, 1])
        else:
            x, y = r[:, 0] - 0.5 * n[0] * d, r[:, 1] - 0.5 * n[1] * d
            rr = np.hypot(x, y)
            pa, pb = (x * v[:, 0] + y * v[:, 1]) / rr, (-y * v[:, 0] + x * v[:, 1]) / rr
        ia = np.round(pa / dv[0]).astype(int) - start  # np.round is half-to-even; an exact tie of a random velocity does not occur
        ib = np.round(pb / dv[1]).astype(int) - start
        keep = inside & (ia >= 0) & (ia < size) & (ib >= 0) & (ib < size)
        ref = np.zeros((size, size))
        np.add.at(ref, (ib[keep], ia[keep]), 1.0 / 12)
        got_start, got = s.velocity_distribution(projector, geometry, p, dv, vmin, vmax)
        assert got_start == start and got.shape == ref.shape
        assert keep.sum() > 100
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 * ref.max())
    s.close()


def test_host_program_runs_the_reference_root_configuration(X, tmp_path):
    """The set-up of the reference's root config.json (eccapfim Langmuir wave: a 2 x 2 x 32 box, 1000 particles per cell,
    MaxwellCosinePerturbation, three LogView levels), shortened to 20 steps, plus a 2D FieldView plane and a
    DistributionMoment region: the host program accepts the schema, conserves energy at the solver tolerance and writes the
    files the reference would."""
    import json
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "xpic_b200", "_build", "xpic_b200.out")
    if not os.path.exists(exe):
        pytest.skip("host program not built")
    cfg = json.load(open(os.path.join(ROOT, "configs", "langmuir_eccapfim.json")))
    cfg["OutputDirectory"] = str(tmp_path / "lw")
    path = tmp_path / "lw.json"
    path.write_text(json.dumps(cfg))
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    out = tmp_path / "lw"
    titles, en = O.read_table(str(out / "temporal" / "energy_conservation.txt"))
    assert en.shape[0] == 21
    _, e = O.read_table(str(out / "temporal" / "energy.txt"))
    total = e[:, 1] + e[:, 2] + e[:, 3]
    assert np.max(np.abs(np.diff(total))) < 2e-6 * total[0]  # SNES atol = rtol = 1e-7, Picard tolerance 0.5e-7 per particle
    assert e[-1, 1] > 10 * e[1, 1] or e[:, 1].max() > 1e-4  # the perturbation drives a field: wE grows from zero
    for name in ("log-EachTimestep.txt", "log-DiagnosePeriodAvg.txt", "log-AllTimestepsSummary.txt"):
        assert (out / name).exists()
    assert len(open(str(out / "log-EachTimestep.txt")).read().splitlines()) == 21
    plane = np.fromfile(str(out / "E_planeX_0001" / "20"), dtype=np.float32)
    assert plane.size == 3 * 1 * 2 * 32  # one cell thick in x
    cur = np.fromfile(str(out / "electrons" / "current" / "20"), dtype=np.float32)
    assert cur.size == 3 * 2 * 2 * 8 and np.abs(cur).max() > 0


@pytest.mark.parametrize("open_z", [False, True])
def test_batched_staging_equals_whole_slab_staging(X, open_z, monkeypatch):
    """Large slabs deposit in batches of P planes through a staging area of P + 2 planes (XPIC_STAGE_GB): the operator
    and the current are bit-identical to the whole-slab staging, for every kernel variant."""
    n = (12, 8, 10)
    o, s0 = make_pair(n=n, Np=15, seed_fields=51)
    pts, ids = o.get_particles(0)
    f = s0.get_field("B")

    def build(stage_gb):
        if stage_gb is None:
            monkeypatch.delenv("XPIC_STAGE_GB", raising=False)
        else:
            monkeypatch.setenv("XPIC_STAGE_GB", stage_gb)
        s = X.Simulation(n, track_ids=True, open_z=open_z)
        s.add_species(Np=15)
        assert s.add_particles(0, pts, ids) == len(ids)
        s.set_field("B", f)
        return s

    whole, batched = build(None), build("6.2e-3")  # 6 planes of 12 x 8 cells x 10.6 KB: batches of 4, 4, 2 planes
    for variant in (0, 4, 3, 2):
        res = []
        for s in (whole, batched):
            s.set_option(0, variant)
            s.deposit()
            res.append((s.operator_download(), s.get_field("currI")))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]), variant
    if not open_z:
        o.deposit()
        ref = csr_to_stencil(o, X.coef_table())
        assert np.max(np.abs(res[1][0] - ref)) / np.max(np.abs(ref)) < 1e-13
    whole.close()
    batched.close()
    s0.close()


def _within(geometry, p, r):
    if geometry == "box":
        return np.all((np.array(p[:3]) <= r) & (r < np.array(p[3:6])), axis=-1)
    d = r - np.array(p[:3])
    return (np.abs(d[..., 2]) < 0.5 * p[4]) & (d[..., 0] ** 2 + d[..., 1] ** 2 <= p[3] ** 2)


@pytest.mark.parametrize("geometry,params", [("box", (1.0, 1.0, 0.5, 4.0, 3.5, 3.0)), ("cylinder", (2.5, 2.25, 2.0, 1.5, 3.0))])
def test_fields_damping_matches_the_reference_formulas(X, geometry, params):
    """FieldsDamping::execute with DampForBox / DampForCylinder (src/commands/fields_damping.cpp:16-112), evaluated with numpy."""
    n, d, coef = (10, 9, 8), (0.5, 0.5, 0.5), 0.8
    s = X.Simulation(n, d=d)
    rng = np.random.default_rng(23)
    E, B, B0 = (rng.standard_normal(s.nown) for _ in range(3))
    for name, f in (("E", E), ("B", B), ("B0", B0)):
        s.set_field(name, f)
    taken = s.fields_damping(geometry, params, coef)
    z, y, x = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    r = np.stack([(x + 0.5) * d[0], (y + 0.5) * d[1], (z + 0.5) * d[2]], axis=-1)
    L = np.array(n) * np.array(d)
    if geometry == "box":
        damp = np.ones(r.shape[:-1])
        for i in range(3):
            hi, lo = r[..., i] > params[3 + i], r[..., i] < params[i]
            t_hi = (r[..., i] - params[3 + i]) / (L[i] - params[3 + i]) - 1.0
            t_lo = r[..., i] / params[i] - 1.0 if params[i] > 0 else np.zeros_like(damp)
            damp = damp * np.where(hi, 1.0 - coef * t_hi ** 2, np.where(lo, 1.0 - coef * t_lo ** 2, 1.0))
    else:
        rr = np.hypot(r[..., 0] - params[0], r[..., 1] - params[1])
        width, delta = params[0] - params[3], rr - params[3]
        delta0 = width * (1.0 + 1.0 / np.sqrt(coef))
        damp = np.where(rr < params[3], 1.0, np.where(delta < delta0, 1.0 - coef * (delta / width - 1.0) ** 2, 0.0))
    damp = np.where(_within(geometry, params, r), 1.0, damp)[..., None]
    Eg, Bg, B0g = (f.reshape(n[2], n[1], n[0], 3) for f in (E, B, B0))
    ref_taken = np.sum(0.5 * Eg ** 2 * (1 - damp ** 2)) + np.sum(0.5 * (Bg - B0g) ** 2 * (1 - damp ** 2))
    assert np.any(damp != 1.0)
    np.testing.assert_allclose(s.get_field("E").reshape(Eg.shape), Eg * damp, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(s.get_field("B").reshape(Eg.shape), (Bg - B0g) * damp + B0g, rtol=1e-14, atol=1e-14)
    assert abs(taken - ref_taken) < 1e-11 * ref_taken
    s.close()


@pytest.mark.parametrize("geometry,params", [("box", (1.0, 0.5, 1.0, 4.0, 3.5, 3.5)), ("cylinder", (2.5, 2.5, 2.5, 1.6, 3.0))])
def test_remove_particles_matches_the_reference_rule(X, geometry, params):
    """RemoveParticles::execute (src/commands/remove_particles.cpp:11-45): cells whose corner lies outside the geometry are emptied."""
    o, s = make_pair(n=(10, 10, 10), Np=8)
    pts, ids = o.get_particles(0)
    corner = np.floor(pts[:, :3] / 0.5) * 0.5
    keep = _within(geometry, params, corner)
    removed, energy = s.remove_particles(geometry, params)
    assert removed == int((~keep).sum()) and 0 < removed < len(ids)
    ref_energy = np.sum(0.5 * 1.0 * np.sum(pts[~keep, 3:] ** 2, axis=1) * (1.0 / 8))
    assert abs(energy - ref_energy) < 1e-11 * ref_energy
    assert s.particle_count() == int(keep.sum())
    got = np.sort(s.get_particles(0)[1])
    assert np.array_equal(got, np.sort(ids[keep]))
    s.step()  # the store is consistent: a step runs on what is left
    assert s.particle_count() == int(keep.sum())
    s.close()


def test_host_program_open_trap_with_step_presets(X, tmp_path):
    """An open-trap set-up in the reference's schema: open z boundary, SetMagneticField with coils, particles in a
    cylinder, and the StepPresets InjectParticles / RemoveParticles / FieldsDamping before every step
    (src/interfaces/simulation.cpp:82-84).  The host program runs it; B0 is the analytic two-coil field (spot-checked on
    the axis against the closed form of a current loop), the tables are complete."""
    import json
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "xpic_b200", "_build", "xpic_b200.out")
    if not os.path.exists(exe):
        pytest.skip("host program not built")
    cfg = json.load(open(os.path.join(ROOT, "configs", "open_trap_ecsim.json")))
    cfg["OutputDirectory"] = str(tmp_path / "trap")
    path = tmp_path / "trap.json"
    path.write_text(json.dumps(cfg))
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = tmp_path / "trap"
    _, e = O.read_table(str(out / "temporal" / "energy.txt"))
    assert e.shape[0] == 13 and np.all(np.isfinite(e))
    assert r.stdout.count("Particles have been removed from") == 24 and r.stdout.count("Fields are damped") == 12
    B0 = np.fromfile(str(out / "B0" / "00"), dtype=np.float32).reshape(24, 12, 12, 3)
    # B_z of two loops on the axis: I R^2 ... the reference's quadrature of I R (R - r cos) / d^3 at r -> 0 gives 2 pi I R^2 / (z^2 + R^2)^1.5
    z = np.arange(24) * 0.5
    on_axis = sum(2 * np.pi * 4.0 * 3.0 ** 2 / ((z - z0) ** 2 + 3.0 ** 2) ** 1.5 for z0 in (-2.0, 14.0))
    # B_z sits at (x + 1/2, y + 1/2, z): the node next to the axis is a quarter cell diagonal away from it
    np.testing.assert_allclose(B0[:, 5, 5, 2], on_axis, rtol=5e-2)
    dens = np.fromfile(str(out / "electrons" / "density" / "12"), dtype=np.float32)
    assert dens.size == 24 * 12 * 12 and dens.max() > 0
    # VelocityDistribution "vz_vxy": 60 x 60 bins from -30 (both axes, as the reference sizes them); |v_perp| >= 0 fills rows >= 30 only
    fv = np.fromfile(str(out / "electrons" / "vz_vxy" / "12"), dtype=np.float32).reshape(60, 60)
    assert fv.sum() > 0 and fv[:29].sum() == 0
