"""ctypes binding of include/xpic_b200.h.

`Simulation` mirrors the members of the reference's ecsim::Simulation / ecsimcorr::Simulation
that other subsystems touch (src/interfaces/simulation.h:21-72, src/impls/ecsim/simulation.h:27-40):
named vectors E, B, B0, Ep, Ec, currI, currJe, one particle sort per add_species(), step() ==
timestep_implementation().  Errors are raised as XpicB200Error carrying xb_last_error().
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = os.path.join(_HERE, "_build", "libxpic_b200.so")

ECSIM, ECSIMCORR, ECCAPFIM = 0, 1, 2
FIELDS = {"E": 0, "B": 1, "B0": 2, "Ep": 3, "Ec": 4, "currI": 5, "currJe": 6, "currI_sort": 7, "currJe_sort": 8, "J": 9, "J_sort": 10, "Ehk": 11}
SCALARS = {"kinetic": 0, "pred_w": 1, "corr_w": 2, "pred_dK": 3, "corr_dK": 4, "lambda_dK": 5, "energy_member": 6, "j_diff_norm": 7}
STAGES = ["clear_sources", "first_push", "advance_fields", "second_push", "correct_fields", "final_update"]
FAMILIES = ["sort", "moments", "second_push", "spmv", "precond", "moments_cells", "moments_ghost", "moments_rows", "sort_keys", "sort_migrate", "sort_scatter"]
OP_L, OP_M, OP_A = 1, 2, 3


class XpicB200Error(RuntimeError):
    pass


class _Grid(C.Structure):
    _fields_ = [("n", C.c_int32 * 3), ("d", C.c_double * 3), ("dt", C.c_double), ("curl_sign", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("track_ids", C.c_int32), ("boundary", C.c_int32 * 3)]


def slab_range(nz, rank, nranks):
    """Planes [z0, z0 + nzl) of z-slab `rank` (DMDA's default split: the first nz % nranks slabs are
    one plane thicker).  Must agree with xb_create (csrc/api.cu)."""
    base, rem = divmod(int(nz), int(nranks))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def owner_rank(z_cell, nz, nranks):
    """Rank owning global cell plane z_cell (interfaces::Particles::add_particle keeps a particle on the
    rank whose box contains floor(z / dz), src/interfaces/particles.cpp:47-57)."""
    base, rem = divmod(int(nz), int(nranks))
    split = rem * (base + 1)
    return z_cell // (base + 1) if z_cell < split else rem + (z_cell - split) // base


def library_path():
    return _LIB


def build_library(force=False, jobs=8):
    """Compile xpic_b200/csrc for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", _CSRC, "-s", "-j", str(jobs)]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return _LIB


_lib = None

# name -> (restype, argtypes); every symbol include/xpic_b200.h declares
_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
SYMBOLS = {
    "xb_last_error": (C.c_char_p, []),
    "xb_version": (C.c_int, []),
    "xb_operator_ncoef": (C.c_int, []),
    "xb_operator_coef_info": (C.c_int, [C.c_int] + [C.POINTER(C.c_int)] * 5),
    "xb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "xb_create": (C.c_int, [C.POINTER(_Grid), C.c_void_p, C.POINTER(C.c_void_p)]),
    "xb_destroy": (C.c_int, [C.c_void_p]),
    "xb_species_add": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int64, _i32p]),
    "xb_particles_append": (C.c_int, [C.c_void_p, C.c_int32, _dp, _u64p, C.c_int64, _i64p]),
    "xb_particles_maxwellian": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, _dp, C.c_uint64, C.c_int32, _i64p]),
    "xb_particles_count": (C.c_int, [C.c_void_p, C.c_int32, _i64p]),
    "xb_particles_download": (C.c_int, [C.c_void_p, C.c_int32, _dp, _u64p, C.c_int64, _i64p]),
    "xb_field_upload": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_field_download": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_solver_set": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32]),
    "xb_solver_info": (C.c_int, [C.c_void_p, C.c_int32, _i32p, _dp, _i32p]),
    "xb_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "xb_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "xb_step_host": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp]),
    "xb_run_steps": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_run_steps_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp]),
    "xb_spmv_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "xb_spmv_profile_read": (C.c_int, [C.c_void_p, _i64p, _dp]),
    "xb_family_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "xb_family_profile_read": (C.c_int, [C.c_void_p, C.c_int32, _i64p, _dp]),
    "xb_field_energy": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_field_sums": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_fields_damping": (C.c_int, [C.c_void_p, C.c_int32, _dp, C.c_double, _dp]),
    "xb_particles_remove": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp]),
    "xb_scalar": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_particle_moments": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "xb_timing": (C.c_int, [C.c_void_p, C.c_int32, _dp, _i64p]),
    "xb_timing_reset": (C.c_int, [C.c_void_p]),
    "xb_launch_count": (C.c_int, [C.c_void_p, _i64p]),
    "xb_spmv": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp]),
    "xb_spmv_bench": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_operator_download": (C.c_int, [C.c_void_p, _dp]),
    "xb_operator_upload": (C.c_int, [C.c_void_p, _dp]),
    "xb_set_option": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "xb_deposit": (C.c_int, [C.c_void_p]),
    "xb_solve": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp]),
    "xb_curl": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp]),
    "xb_kernel_bench": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_nonlinear_set": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32]),
    "xb_nonlinear_info": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _dp, _dp, _dp]),
    "xb_nonlinear_history": (C.c_int, [C.c_void_p, _dp, C.c_int32, _i32p]),
    "xb_nonlinear_profile": (C.c_int, [C.c_void_p, C.c_int32, _i64p, _dp]),
    "xb_eccapfim_function": (C.c_int, [C.c_void_p, _dp, _dp]),
    "xb_charge_density": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "xb_charge_conservation": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "xb_momentum": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "xb_distribution_moment": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "xb_distribution_moment_region": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p, _dp]),
    "xb_velocity_distribution_size": (C.c_int, [_dp, _dp, _dp, _i32p, _i32p]),
    "xb_velocity_distribution": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp]),
}


def load_library():
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise XpicB200Error(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
        L = C.CDLL(_LIB, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise XpicB200Error(load_library().xb_last_error().decode("utf-8", "replace"))


def _as_dp(a):
    return a.ctypes.data_as(_dp)


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    _check(load_library().xb_comm_unique_id(buf))
    return bytes(buf)


def coef_table():
    """(369, 5) int array: c1, c2, dx, dy, dz of every coefficient slot."""
    L = load_library()
    n = L.xb_operator_ncoef()
    out = np.zeros((n, 5), dtype=np.int64)
    v = [C.c_int() for _ in range(5)]
    for k in range(n):
        _check(L.xb_operator_coef_info(k, *[C.byref(x) for x in v]))
        out[k] = [x.value for x in v]
    return out


class Simulation:
    def __init__(self, n, d=(0.5, 0.5, 0.5), dt=1.5, scheme=ECSIM, curl_sign=+1, device=0, rank=0, nranks=1, comm_id=None,
                 track_ids=True, open_z=False):
        self._L = load_library()
        self.n = tuple(int(v) for v in n)
        self.d = tuple(float(v) for v in d)
        self.dt = float(dt)
        self.scheme = scheme
        self.rank, self.nranks = rank, nranks
        self.z0, self.nzl = slab_range(self.n[2], rank, nranks)
        self.ncl = self.n[0] * self.n[1] * self.nzl
        self.nown = 3 * self.ncl
        g = _Grid()
        g.n[:] = self.n
        g.d[:] = self.d
        g.dt = self.dt
        g.curl_sign = curl_sign
        g.device = device
        g.rank = rank
        g.nranks = nranks
        g.track_ids = 1 if track_ids else 0
        g.boundary[:] = (0, 0, 1 if open_z else 0)  # da_boundary_z = DM_BOUNDARY_NONE / GHOSTED
        self.track_ids = bool(track_ids)
        h = C.c_void_p()
        uid = C.create_string_buffer(comm_id, 128) if comm_id is not None else None
        _check(self._L.xb_create(C.byref(g), uid, C.byref(h)))
        self._h = h
        self.nsorts = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.xb_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # -- particles (interfaces::Particles) -----------------------------------------------------
    def add_species(self, q=-1.0, m=1.0, n=1.0, Np=100, capacity=None):
        if capacity is None:
            capacity = int(self.ncl * Np * 1.5) + 4096
        sid = C.c_int32()
        _check(self._L.xb_species_add(self._h, q, m, n, Np, capacity, C.byref(sid)))
        self.nsorts += 1
        return sid.value

    def add_particles(self, sid, aos6, ids=None):
        a = np.ascontiguousarray(aos6, dtype=np.float64).reshape(-1, 6)
        added = C.c_int64()
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            idp = ids.ctypes.data_as(_u64p)
        _check(self._L.xb_particles_append(self._h, sid, _as_dp(a), idp, a.shape[0], C.byref(added)))
        return added.value

    def set_particles_maxwellian(self, sid, total, T=0.1, seed=1, tov=True):
        """SetParticles{CoordinateInBox, MaxwellianMomentum} with a counter-based generator; `total`
        counts particles of the GLOBAL box, the return value those that landed in this rank's slab."""
        Tv = np.array([T, T, T] if np.isscalar(T) else T, dtype=np.float64)
        added = C.c_int64()
        _check(self._L.xb_particles_maxwellian(self._h, sid, int(total), _as_dp(Tv), int(seed), int(tov), C.byref(added)))
        return added.value

    def particle_count(self, sid=0):
        n = C.c_int64()
        _check(self._L.xb_particles_count(self._h, sid, C.byref(n)))
        return n.value

    def get_particles(self, sid=0):
        n = self.particle_count(sid)
        a = np.empty((n, 6), dtype=np.float64)
        ids = np.empty(n, dtype=np.uint64) if self.track_ids else None
        cnt = C.c_int64()
        _check(self._L.xb_particles_download(self._h, sid, _as_dp(a), ids.ctypes.data_as(_u64p) if ids is not None else None, n, C.byref(cnt)))
        return a, ids

    # -- named vectors -------------------------------------------------------------------------
    def get_field(self, name, sid=0):
        out = np.empty(self.nown, dtype=np.float64)
        _check(self._L.xb_field_download(self._h, FIELDS[name], sid, _as_dp(out)))
        return out

    def set_field(self, name, arr, sid=0):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        assert a.size == self.nown
        _check(self._L.xb_field_upload(self._h, FIELDS[name], sid, _as_dp(a)))

    # -- solver --------------------------------------------------------------------------------
    def solver_set(self, which=0, rtol=1e-7, atol=1e-7, maxit=100, restart=30, precond=0):
        _check(self._L.xb_solver_set(self._h, which, rtol, atol, maxit, restart, precond))

    def solver_info(self, which=0):
        it, rn, re = C.c_int32(), C.c_double(), C.c_int32()
        _check(self._L.xb_solver_info(self._h, which, C.byref(it), C.byref(rn), C.byref(re)))
        return it.value, rn.value, re.value

    # -- eccapfim's nonlinear solve (SNES in the reference) -------------------------------------
    def nonlinear_set(self, atol=1e-7, rtol=1e-7, stol=1e-7, maxit=1000, depth=10, cheb_degree=12, particle_tol=0.5e-7, particle_maxit=30):
        _check(self._L.xb_nonlinear_set(self._h, atol, rtol, stol, maxit, depth, cheb_degree, particle_tol, particle_maxit))

    def nonlinear_info(self):
        it, fe, re = C.c_int32(), C.c_int32(), C.c_int32()
        fn, ai, ac = C.c_double(), C.c_double(), C.c_double()
        _check(self._L.xb_nonlinear_info(self._h, C.byref(it), C.byref(fe), C.byref(re), C.byref(fn), C.byref(ai), C.byref(ac)))
        return {"iterations": it.value, "fevals": fe.value, "reason": re.value, "fnorm": fn.value, "avg_cn": ai.value, "avg_cells": ac.value}

    def nonlinear_history(self):
        n = C.c_int32()
        buf = np.empty(2048)
        _check(self._L.xb_nonlinear_history(self._h, _as_dp(buf), buf.size, C.byref(n)))
        return buf[: min(n.value, buf.size)].copy()

    def nonlinear_profile(self, enable=-1):
        """(particle passes timed, their summed ms) so far; enable in {0, 1} then resets and switches collecting."""
        n, ms = C.c_int64(), C.c_double()
        _check(self._L.xb_nonlinear_profile(self._h, int(enable), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def eccapfim_function(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.empty(self.nown, dtype=np.float64)
        _check(self._L.xb_eccapfim_function(self._h, _as_dp(x), _as_dp(f)))
        return f

    # -- device-side diagnostics ----------------------------------------------------------------
    def charge_density(self, sid=0):
        """ParticlesChargeDensity::collect: rho of sort sid on the owned nodes (also remembered for charge_conservation)."""
        rho = np.empty(self.ncl, dtype=np.float64)
        _check(self._L.xb_charge_density(self._h, sid, _as_dp(rho)))
        return rho

    def charge_conservation(self, current="currJe"):
        """ChargeConservation::add_columns: (nsorts + 1, 2) array of the 1- and 2-norms of d rho / dt + div J."""
        out = np.zeros(2 * (self.nsorts + 1))
        _check(self._L.xb_charge_conservation(self._h, {"currJe": 0, "J": 1}[current], _as_dp(out)))
        return out.reshape(-1, 2)

    MOMENTS = {"density": (0, 1), "current": (1, 3), "momentum_flux": (2, 6), "momentum_flux_cyl": (3, 6), "momentum_flux_diag": (4, 3),
               "momentum_flux_diag_cyl": (5, 3)}

    def distribution_moment(self, name, sid=0, start=None, size=None):
        """DistributionMoment::collect: (ncl, components) array of the named moment on the owned cells; start / size (cells)
        restrict the particles to a region."""
        mid, ms = self.MOMENTS[name]
        out = np.empty((self.ncl, ms), dtype=np.float64)
        if start is None and size is None:
            _check(self._L.xb_distribution_moment(self._h, sid, mid, _as_dp(out)))
        else:
            st = (C.c_int32 * 3)(*(start if start is not None else (0, 0, 0)))
            sz = (C.c_int32 * 3)(*(size if size is not None else self.n))
            _check(self._L.xb_distribution_moment_region(self._h, sid, mid, st, sz, _as_dp(out)))
        return out

    PROJECTORS = {"vx_vy": 0, "vz_vxy": 1, "vr_vphi": 2}

    def velocity_distribution(self, projector, geometry, p, dv, vmin=(-1.0, -1.0), vmax=(1.0, 1.0), sid=0):
        """VelocityDistribution::collect: (start, f) with f[second projection bin][first projection bin] summed over all
        ranks; bins start..start+len(f) on both axes (the reference sizes both from vmin[0], vmax[0], dv[0])."""
        p = np.ascontiguousarray(np.resize(np.asarray(p, dtype=np.float64), 6))
        dv, vmin, vmax = (np.ascontiguousarray(v, dtype=np.float64) for v in (dv, vmin, vmax))
        start, size = C.c_int32(0), C.c_int32(0)
        _check(self._L.xb_velocity_distribution_size(_as_dp(dv), _as_dp(vmin), _as_dp(vmax), C.byref(start), C.byref(size)))
        out = np.zeros((size.value, size.value), dtype=np.float64)
        _check(self._L.xb_velocity_distribution(self._h, sid, self.PROJECTORS[projector], {"box": 0, "cylinder": 1}[geometry], _as_dp(p), _as_dp(dv),
                                                _as_dp(vmin), _as_dp(vmax), _as_dp(out)))
        return start.value, out

    def density(self, sid=0):
        """DistributionMoment "density": cell-centred number density of sort sid on the owned cells."""
        out = np.empty(self.ncl, dtype=np.float64)
        _check(self._L.xb_distribution_moment(self._h, sid, 0, _as_dp(out)))
        return out

    def momentum(self, sid=0):
        """MomentumConservation::calculate: (P[3], QE[3]) of sort sid with the present E."""
        out = np.zeros(6)
        _check(self._L.xb_momentum(self._h, sid, _as_dp(out)))
        return out[:3], out[3:]

    # -- stepping ------------------------------------------------------------------------------
    def step(self, scheme=None):
        _check(self._L.xb_step(self._h, self.scheme if scheme is None else scheme))

    def stage(self, stage, scheme=None):
        idx = STAGES.index(stage) if isinstance(stage, str) else stage
        _check(self._L.xb_stage(self._h, self.scheme if scheme is None else scheme, idx))

    def step_host(self, E, B, B0=None, kinetic=None, scheme=None):
        """E, B (in/out), B0 (in): host arrays (pinned torch tensors' numpy views work)."""
        _check(self._L.xb_step_host(self._h, self.scheme if scheme is None else scheme, _as_dp(E), _as_dp(B),
                                    _as_dp(B0) if B0 is not None else None, _as_dp(kinetic) if kinetic is not None else None))

    def run_steps(self, k, scheme=None):
        """k steps, state resident; returns device-timeline ms (CUDA events on the launching stream)."""
        ms = C.c_double()
        _check(self._L.xb_run_steps(self._h, self.scheme if scheme is None else scheme, k, C.byref(ms)))
        return ms.value

    def run_steps_host(self, k, E, B, B0=None, kinetic=None, scheme=None):
        ms = C.c_double()
        _check(self._L.xb_run_steps_host(self._h, self.scheme if scheme is None else scheme, k, _as_dp(E), _as_dp(B),
                                         _as_dp(B0) if B0 is not None else None, _as_dp(kinetic) if kinetic is not None else None, C.byref(ms)))
        return ms.value

    def spmv_profile(self, enable=True):
        _check(self._L.xb_spmv_profile(self._h, int(enable)))

    def spmv_profile_read(self):
        n, ms = C.c_int64(), C.c_double()
        _check(self._L.xb_spmv_profile_read(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def family_profile(self, enable=True):
        """CUDA-event timing of the kernel families inside the step (FAMILIES); enabling resets the counters."""
        _check(self._L.xb_family_profile(self._h, int(enable)))

    def family_profile_read(self):
        out = {}
        for i, name in enumerate(FAMILIES):
            n, ms = C.c_int64(), C.c_double()
            _check(self._L.xb_family_profile_read(self._h, i, C.byref(n), C.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    def field_energy(self, name, sid=0):
        """0.5 |v|^2 over ALL ranks, reduced on the device (Energy::calculate_energy)."""
        out = C.c_double()
        _check(self._L.xb_field_energy(self._h, FIELDS[name], sid, C.byref(out)))
        return out.value

    def fields_damping(self, geometry, params, coefficient):
        """FieldsDamping::execute; geometry "box" (min[3], max[3]) or "cylinder" (center[3], radius, height); returns the damped energy."""
        p = np.zeros(6)
        p[: len(params)] = params
        out = C.c_double()
        _check(self._L.xb_fields_damping(self._h, {"box": 0, "cylinder": 1}[geometry], _as_dp(p), float(coefficient), C.byref(out)))
        return out.value

    def remove_particles(self, geometry, params, sid=0):
        """RemoveParticles::execute; returns (removed particles, removed kinetic energy) over all ranks."""
        p = np.zeros(6)
        p[: len(params)] = params
        out = np.zeros(2)
        _check(self._L.xb_particles_remove(self._h, sid, {"box": 0, "cylinder": 1}[geometry], _as_dp(p), _as_dp(out)))
        return int(out[0]), float(out[1])

    def scalar(self, name, sid=0):
        out = C.c_double()
        _check(self._L.xb_scalar(self._h, sid, SCALARS[name], C.byref(out)))
        return out.value

    def particle_moments(self, sid=0):
        out = np.zeros(5)
        _check(self._L.xb_particle_moments(self._h, sid, _as_dp(out)))
        return out

    def field_energies(self):
        E, B = self.get_field("E"), self.get_field("B")
        return 0.5 * float(E @ E), 0.5 * float(B @ B)

    def timing(self):
        out = {}
        for i, name in enumerate(STAGES):
            s, n = C.c_double(), C.c_int64()
            _check(self._L.xb_timing(self._h, i, C.byref(s), C.byref(n)))
            out[name] = (s.value, n.value)
        return out

    def timing_reset(self):
        _check(self._L.xb_timing_reset(self._h))

    def launch_count(self):
        n = C.c_int64()
        _check(self._L.xb_launch_count(self._h, C.byref(n)))
        return n.value

    # -- hooks ---------------------------------------------------------------------------------
    def spmv(self, x, op=OP_A):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.nown, dtype=np.float64)
        _check(self._L.xb_spmv(self._h, op, _as_dp(x), _as_dp(y)))
        return y

    def spmv_bench(self, op=OP_A, reps=100):
        ms = C.c_double()
        _check(self._L.xb_spmv_bench(self._h, op, reps, C.byref(ms)))
        return ms.value

    def operator_download(self):
        out = np.empty((self._L.xb_operator_ncoef(), self.ncl), dtype=np.float64)
        _check(self._L.xb_operator_download(self._h, _as_dp(out)))
        return out

    def operator_upload(self, coef):
        a = np.ascontiguousarray(coef, dtype=np.float64)
        assert a.shape == (self._L.xb_operator_ncoef(), self.ncl)
        _check(self._L.xb_operator_upload(self._h, _as_dp(a)))

    def set_option(self, what, value):
        _check(self._L.xb_set_option(self._h, what, value))

    def deposit(self):
        _check(self._L.xb_deposit(self._h))

    def solve(self, b, which=0, op=OP_A):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.nown, dtype=np.float64)
        _check(self._L.xb_solve(self._h, which, op, _as_dp(b), _as_dp(x)))
        return x

    def curl(self, f, positive=True):
        f = np.ascontiguousarray(f, dtype=np.float64)
        y = np.empty(self.nown, dtype=np.float64)
        _check(self._L.xb_curl(self._h, int(positive), _as_dp(f), _as_dp(y)))
        return y

    def kernel_bench(self, what, reps=5):
        ms = C.c_double()
        _check(self._L.xb_kernel_bench(self._h, what, reps, C.byref(ms)))
        return ms.value
