"""ctypes binding of the CPU oracle (oracle/xpic_oracle.cpp).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from xpic_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libxpic_oracle.so")

ECSIM, ECSIMCORR, ECCAPFIM = 0, 1, 2
FIELDS = {"E": 0, "B": 1, "B0": 2, "Ep": 3, "Ec": 4, "currI": 5, "currJe": 6, "currI_sort": 7, "currJe_sort": 8, "J": 9, "J_sort": 10, "Ehk": 11}
SCALARS = {"energy": 0, "pred_w": 1, "corr_w": 2, "pred_dK": 3, "corr_dK": 4, "lambda_dK": 5, "energy_member": 6, "j_diff_norm": 7}


def build(force=False):
    src = os.path.join(_HERE, "xpic_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        dp = C.POINTER(C.c_double)
        L.xo_set_threads.argtypes = [C.c_int]
        L.xo_max_threads.restype = C.c_int
        L.xo_create.restype = C.c_void_p
        L.xo_create.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int]
        L.xo_destroy.argtypes = [C.c_void_p]
        L.xo_add_species.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        L.xo_set_particles_maxwell.restype = C.c_long
        L.xo_set_particles_maxwell.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_long]
        L.xo_set_particles.restype = C.c_long
        L.xo_set_particles.argtypes = [C.c_void_p, C.c_int, dp, C.c_long]
        L.xo_particle_count.restype = C.c_long
        L.xo_particle_count.argtypes = [C.c_void_p, C.c_int]
        L.xo_get_particles.argtypes = [C.c_void_p, C.c_int, dp, C.POINTER(C.c_uint64)]
        L.xo_step.argtypes = [C.c_void_p, C.c_int]
        L.xo_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.xo_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.xo_solver_set.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.xo_solver_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), dp, C.POINTER(C.c_int)]
        L.xo_scalar.restype = C.c_double
        L.xo_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.xo_deposit.argtypes = [C.c_void_p]
        L.xo_csr_nnz.restype = C.c_long
        L.xo_csr_nnz.argtypes = [C.c_void_p, C.c_int]
        L.xo_csr_export.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32), dp]
        L.xo_spmv.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.xo_curl.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.xo_set_open_z.argtypes = [C.c_void_p, C.c_int]
        L.xo_interpolate.argtypes = [C.c_void_p, dp, dp, dp]
        L.xo_boris_update_vEB.argtypes = [C.c_double, C.c_double, dp, dp, dp]
        L.xo_esirkepov.argtypes = [C.c_void_p, dp, dp, C.c_double, dp]
        L.xo_snes_set.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double]
        L.xo_snes_set_particle_tol.argtypes = [C.c_void_p, C.c_double]
        L.xo_snes_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp]
        L.xo_snes_history.restype = C.c_int
        L.xo_snes_history.argtypes = [C.c_void_p, dp, C.c_int]
        L.xo_eccapfim_function.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.xo_crank_nicolson_uniform.argtypes = [C.c_double, C.c_double, dp, dp, dp, dp]
        L.xo_cell_traversal.restype = C.c_int
        L.xo_cell_traversal.argtypes = [C.c_void_p, dp, dp, dp, C.c_int]
        _lib = L
    return _lib


def set_threads(n):
    """OpenMP threads used by the particle / matrix / Krylov loops (default 1: deterministic)."""
    lib().xo_set_threads(int(n))


def max_threads():
    return lib().xo_max_threads()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One simulation box (all axes periodic), mirroring ecsim::Simulation / ecsimcorr::Simulation."""

    def __init__(self, n, d=(0.5, 0.5, 0.5), dt=1.5, curl_sign=+1, open_z=False):
        self.n = tuple(int(v) for v in n)
        self.d = tuple(float(v) for v in d)
        self.dt = float(dt)
        self.n3 = 3 * self.n[0] * self.n[1] * self.n[2]
        self._h = lib().xo_create(*self.n, *self.d, self.dt, int(curl_sign))
        if open_z:  # da_boundary_z = DM_BOUNDARY_NONE / GHOSTED
            lib().xo_set_open_z(self._h, 1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().xo_destroy(self._h)
            self._h = None

    def add_species(self, q=-1.0, m=1.0, n=1.0, Np=100):
        return lib().xo_add_species(self._h, q, m, n, Np)

    def set_particles_maxwell(self, sid, T=0.1, tov=True, count=-1):
        T = (T, T, T) if np.isscalar(T) else T
        return lib().xo_set_particles_maxwell(self._h, sid, T[0], T[1], T[2], int(tov), count)

    def set_particles(self, sid, aos6):
        a = np.ascontiguousarray(aos6, dtype=np.float64).reshape(-1, 6)
        return lib().xo_set_particles(self._h, sid, _dp(a), a.shape[0])

    def particle_count(self, sid=0):
        return lib().xo_particle_count(self._h, sid)

    def get_particles(self, sid=0):
        n = self.particle_count(sid)
        a = np.empty((n, 6), dtype=np.float64)
        ids = np.empty(n, dtype=np.uint64)
        lib().xo_get_particles(self._h, sid, _dp(a), ids.ctypes.data_as(C.POINTER(C.c_uint64)))
        return a, ids

    def step(self, scheme=ECSIM):
        lib().xo_step(self._h, scheme)

    def get_field(self, name, sid=0):
        out = np.empty(self.n3, dtype=np.float64)
        lib().xo_get_field(self._h, FIELDS[name], sid, _dp(out))
        return out

    def set_field(self, name, arr, sid=0):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        assert a.size == self.n3
        lib().xo_set_field(self._h, FIELDS[name], sid, _dp(a))

    def solver_set(self, which=0, rtol=1e-7, atol=1e-7, maxit=100, restart=30):
        lib().xo_solver_set(self._h, which, rtol, atol, maxit, restart)

    def solver_info(self, which=0):
        it, rn, re = C.c_int(), C.c_double(), C.c_int()
        lib().xo_solver_info(self._h, which, C.byref(it), C.byref(rn), C.byref(re))
        return it.value, rn.value, re.value

    def snes_set(self, atol=1e-7, rtol=1e-7, stol=1e-7, maxit=1000, precond=0, shift=0.0):
        """eccapfim nonlinear solver (reference defaults: eccapfim/simulation.h:14-19).  precond=1 solves
        P F = 0 with P = ((1 + shift) I + dt^2/4 curl curl)^-1 -- not in the reference."""
        lib().xo_snes_set(self._h, atol, rtol, stol, maxit, int(precond), float(shift))

    def snes_set_particle_tol(self, tol=0.5e-7):
        lib().xo_snes_set_particle_tol(self._h, float(tol))

    def snes_info(self):
        it, fe, re = C.c_int(), C.c_int(), C.c_int()
        ai, ac = C.c_double(), C.c_double()
        lib().xo_snes_info(self._h, C.byref(it), C.byref(fe), C.byref(re), C.byref(ai), C.byref(ac))
        return {"iterations": it.value, "fevals": fe.value, "reason": re.value, "avg_cn": ai.value, "avg_cells": ac.value}

    def snes_history(self):
        buf = np.empty(2048)
        n = lib().xo_snes_history(self._h, _dp(buf), buf.size)
        return buf[: min(n, buf.size)].copy()

    def eccapfim_function(self, x, prepare=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.empty(self.n3)
        lib().xo_eccapfim_function(self._h, _dp(x), _dp(f), int(prepare))
        return f

    def cell_traversal(self, end, start):
        e = np.ascontiguousarray(end, dtype=np.float64)
        s0 = np.ascontiguousarray(start, dtype=np.float64)
        out = np.empty((64, 3))
        n = lib().xo_cell_traversal(self._h, _dp(e), _dp(s0), _dp(out), 64)
        return out[:n].copy()

    def scalar(self, name, sid=0):
        return lib().xo_scalar(self._h, sid, SCALARS[name])

    def field_energies(self):
        E, B = self.get_field("E"), self.get_field("B")
        return 0.5 * float(E @ E), 0.5 * float(B @ B)

    def deposit(self):
        lib().xo_deposit(self._h)

    def csr(self, which=0):
        nnz = lib().xo_csr_nnz(self._h, which)
        rp = np.empty(self.n3 + 1, dtype=np.int64)
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.float64)
        lib().xo_csr_export(self._h, which, rp.ctypes.data_as(C.POINTER(C.c_int64)), col.ctypes.data_as(C.POINTER(C.c_int32)), _dp(val))
        return rp, col, val

    def spmv(self, x, L=True, M=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n3, dtype=np.float64)
        lib().xo_spmv(self._h, (1 if L else 0) | (2 if M else 0), _dp(x), _dp(y))
        return y

    def curl(self, f, positive=True):
        f = np.ascontiguousarray(f, dtype=np.float64)
        y = np.empty(self.n3, dtype=np.float64)
        lib().xo_curl(self._h, int(positive), _dp(f), _dp(y))
        return y

    def interpolate(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        e, b = np.zeros(3), np.zeros(3)
        lib().xo_interpolate(self._h, _dp(r), _dp(e), _dp(b))
        return e, b

    def esirkepov(self, old_r, new_r, alpha):
        o = np.ascontiguousarray(old_r, dtype=np.float64)
        n = np.ascontiguousarray(new_r, dtype=np.float64)
        J = np.zeros(self.n3)
        lib().xo_esirkepov(self._h, _dp(o), _dp(n), alpha, _dp(J))
        return J


def boris_update_vEB(dt, qm, E, B, v):
    E = np.ascontiguousarray(E, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    lib().xo_boris_update_vEB(dt, qm, _dp(E), _dp(B), _dp(v))
    return v


def crank_nicolson_uniform(dt, qm, E, B, r, v):
    E = np.ascontiguousarray(E, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    r = np.array(r, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    lib().xo_crank_nicolson_uniform(dt, qm, _dp(E), _dp(B), _dp(r), _dp(v))
    return r, v


def read_table(path):
    """Parse a reference temporal/*.txt table -> (titles, ndarray)."""
    with open(path) as f:
        titles = f.readline().split()
        rows = [[float(v) for v in line.split()] for line in f if line.strip()]
    return titles, np.array(rows)
