// eccapfim.cu -- the fully implicit energy- and charge-conserving step (BASELINE config 5).
//
// Reference code replaced (paths relative to the reference tree):
//   eccapfim::Simulation::timestep_implementation / init_iteration / calc_iteration / after_iteration
//                                            src/impls/eccapfim/simulation.cpp:36-130
//   Simulation::form_iteration / form_current / form_function      simulation.cpp:132-241
//   eccapfim::Particles::form_iteration (Picard-iterated Crank-Nicolson mover with path splitting)
//                                            src/impls/eccapfim/particles.cpp:30-181
//   cell_traversal                           src/impls/eccapfim/cell_traversal.cpp:3-77
//   ImplicitEsirkepov::{Shape::setup, interpolate, decompose}   src/algorithms/implicit_esirkepov.cpp:11-117
//   Shape::setup(r) + SimpleInterpolation (B at the segment midpoint)  src/utils/shape.cpp:34-45,81-107,
//                                            src/algorithms/simple_interpolation.cpp:8-38
//   SNESSolve (PETSc NGMRES, un-vendored)    simulation.cpp:75, init_snes_solver :358-390
//
// The nonlinear system F(E^{n+1/2}) = 0 is the reference's (form_function :228-236):
//   F(x) = x + dt^2/4 curl^- curl^+ x - E^n + dt/2 J(x) - dt/2 curl^- B^n ,
// J(x) = the current of all particles re-pushed from their start-of-step state through x.
// PETSc's NGMRES iteration path is not restated (the golden convergence history was produced by an
// older revision of the reference whose residual is scaled by 2/dt); the solver here
// is Anderson acceleration of the preconditioned fixed-point map x <- x - P F(x) with
// P ~= ((1 + sigma) I + dt^2/4 curl curl)^-1 (Chebyshev polynomial, sigma = sum_s (omega_ps dt)^2 / 4):
// one residual evaluation per iteration and ~10 instead of the reference's ~105 per step.  The stop
// test is SNESConvergedDefault's on the true |F| with the reference's tolerances (simulation.h:14-19).
//
// Particle pass (k_cap_push_tasks, the default; k_cap_push = one thread per particle, cross-check): one CTA
// per 16 x-consecutive cells; the E^{n+1/2,k} and B^n nodes a particle of these cells can reach within one
// cell of motion are staged in shared memory, the current is accumulated in a shared tile and flushed once
// (fast particles fall back to global loads / reductions).  Owners advance one particle each through
// emit -> wait -> consume, all threads evaluate the emitted path pieces as dense tasks (see the kernel).
// The Picard iteration of evaluation k + 1 starts from the velocity evaluation k converged to (option 4).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"
#include "operators.cuh"

namespace xb {

int cheb_solve_shifted(xb_ctx* c, int deg, double diag, const double* u, double* z);  // krylov.cu
int reduce_finish(xb_ctx* c, int nv, double* host_out);                                // fields.cu
int migrate_and_sort(xb_ctx* c, Species& s, double dt_move);                           // migrate.cu

// periodic index of a node a few cells outside [0, n): compare / add instead of the runtime modulo of wrapi()
__device__ __forceinline__ int wrap_near(int i, int n)
{
  while (i < 0) i += n;
  while (i >= n) i -= n;
  return i;
}

constexpr int CAP_CELLS = 8;
constexpr int CAP_THREADS = 256;
constexpr int CAP_LO = 2;  // staged nodes below the first cell (x), the row (y) and the plane (z)
constexpr int CAP_NX = CAP_CELLS + 5, CAP_NY = 6, CAP_NZ = 6;
constexpr int CAP_VOL = CAP_NX * CAP_NY * CAP_NZ;  // nodes per component, layout [c][z][y][x]

struct CapArgs {
  const double* p0[6];  // start-of-step state (cell-sorted)
  double* pn[6];        // (x^{n+1,k}, v^{n+1,k}) of this evaluation
  const int32_t* bin_start;
  const double* E;  // E^{n+1/2,k}, ghosts valid (width GZ)
  const double* B;  // B^n, ghosts valid
  double* J;        // this sort's current (ghosted, zeroed by the caller): 64-bit fixed-point accumulators during the pass
  double inv_unit;  // 1 / unit of those accumulators (a power of two)
  double q, m, mpw;
  double cn_tol;
  int cn_maxit;
  int groups_x;
  unsigned long long* counters;  // [0] Picard iterations, [1] traversed segments
  int* error;
  int warm;  // pn holds the pushed state of the previous evaluation of this step: start Picard from it
};

template <int NC>
struct CapCtxT {
  static constexpr int NX = NC + 5, NY = 6, NZ = 6, VOL = NX * NY * NZ;
  Grid g;
  const double* Et;
  const double* Bt;
  double* Jt;
  int x0, y0, z0;  // global (unwrapped) node index of the tile origin
  const double* E;
  const double* B;
  double* J;
  int* error;

  __device__ __forceinline__ bool inside(const int* lo, int ext) const
  {
    return lo[0] >= x0 && lo[0] + ext <= x0 + NX && lo[1] >= y0 && lo[1] + ext <= y0 + NY && lo[2] >= z0 && lo[2] + ext <= z0 + NZ;
  }
  __device__ __forceinline__ int tile_base(const int* lo) const { return ((lo[2] - z0) * NY + (lo[1] - y0)) * NX + (lo[0] - x0); }
  // element index in a ghosted global vector of the (unwrapped) global node (gx, gy, gz)
  __device__ __forceinline__ int64_t gidx(int gx, int gy, int gz, int c) const
  {
    const int x = wrapi(gx, g.nx), y = wrapi(gy, g.ny);
    int zl;
    if (g.nranks == 1)
      zl = wrapi(gz, g.nz);
    else {
      zl = gz - g.z0;
      if (zl < -GZ) zl += g.nz;
      else if (zl >= g.nzl + GZ) zl -= g.nz;
      if (zl < -GZ || zl >= g.nzl + GZ) {
        *error = 2;  // the particle left the ghost planes of this slab within one step
        zl = min(max(zl, -GZ), g.nzl + GZ - 1);
      }
    }
    return g.vidx(x, y, zl, c);
  }
};
using CapCtx = CapCtxT<CAP_CELLS>;

// ---- shapes -----------------------------------------------------------------------------------
__device__ __forceinline__ double sf1(double x) { return 1.0 - fabs(x); }
__device__ __forceinline__ double sf2(int j, double x)  // sfunc_2[j], implicit_esirkepov.h:40-44
{
  x = fabs(x);
  return j == 1 ? (0.75 - x * x) : 0.5 * ((1.5 - x) * (1.5 - x));
}
__device__ __forceinline__ double spline2(double x)  // interfaces/sort_parameters.cpp:21-30
{
  x = fabs(x);
  if (x <= 0.5) return (0.75 - x * x);
  if (x < 1.5) return 0.5 * (1.5 - x) * (1.5 - x);
  return 0.0;
}

// |(x, y, z)|: the reference's Vector3::length() is std::hypot (utils/vector3.h:160-164); the plain
// square root differs from it by an ulp at most and avoids hypot's rescaling (lengths here are O(1))
__device__ __forceinline__ double len3(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

__device__ __forceinline__ void cells3(const Grid& g, const double* r, double* p)
{
  p[0] = to_cells(r[0], g.dx, g.inv_dx, g.exact_inv & 1);
  p[1] = to_cells(r[1], g.dy, g.inv_dy, g.exact_inv & 2);
  p[2] = to_cells(r[2], g.dz, g.inv_dz, g.exact_inv & 4);
}

// ImplicitEsirkepov::Shape::setup (implicit_esirkepov.cpp:11-60): the 54 weights factorise into
// per-axis tables; start = round(midpoint) - 1
struct CapW {
  int start[3];
  double sh1[3][2];          // 1/6 * S1 at the two points along the component's own axis
  double sn[3][3], s0[3][3];  // S2 at the segment end (n) and start (0)
};

__device__ __forceinline__ void cap_weights(const Grid& g, const double* rn, const double* r0, CapW& w)
{
  double prn[3], pr0[3];
  cells3(g, rn, prn);
  cells3(g, r0, pr0);
  constexpr double sixth = 1.0 / 6.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double prh = 0.5 * (prn[a] + pr0[a]);
    const double gc = round(prh);
    w.start[a] = (int)gc - 1;
    const double gv = gc + 0.5;
#pragma unroll
    for (int i = 0; i < 2; ++i) w.sh1[a][i] = sixth * sf1(gv + (i - 1) - prh);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      w.sn[a][j] = sf2(j, gc + (j - 1) - prn[a]);
      w.s0[a][j] = sf2(j, gc + (j - 1) - pr0[a]);
    }
  }
}

// E gather (FAST: from the tile) -- implicit_esirkepov.cpp:71-90; DEPOSIT: the same loop adds
// alpha * v[c] * weight into J (:97-116)
// The current is accumulated in 64-bit FIXED POINT (unit = a power of two 2^-46 below q n / Np, CapArgs::inv_unit
// is folded into alpha): integer additions commute, so the result does not depend on the order in which
// threads, warps and CTAs arrive -- two runs are bit-identical -- and the additions need no compare-and-swap spin
// (sm_100a has no native 64-bit shared atomic, fp64 or integer).  One entry of the shared tile is two 32-bit
// words updated with native 32-bit atomics: the low word reports its carry through the value it returns.
// lo / hi: shared-memory byte addresses of the entry's two words.  The low words of a tile are one array, the high
// words another one behind it (HI_OFFSET bytes): consecutive entries are consecutive banks for both.
__device__ __forceinline__ void shared_add_fixed(unsigned lo_addr, unsigned hi_addr, long long x)
{
  const unsigned lo = (unsigned)x;
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(lo_addr), "r"(lo) : "memory");
  const int hi = (int)(x >> 32) + ((old + lo) < old ? 1 : 0);
  if (hi != 0) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(hi_addr), "r"(hi) : "memory");
}

// entry e of a tile of n entries kept as n low words followed by n high words
__device__ __forceinline__ long long tile_entry_fixed(const double* tile, int n, int e)
{
  const unsigned* w = reinterpret_cast<const unsigned*>(tile);
  return (long long)(((unsigned long long)w[n + e] << 32) | (unsigned long long)w[e]);
}

__device__ __forceinline__ void global_add_fixed(double* entry, long long x)
{
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(entry)), "l"(x) : "memory");
}

template <bool FAST, bool DEPOSIT, class K>
__device__ __forceinline__ void cap_apply(const K& k, const CapW& w, double* Ep, double alpha, const double* v)
{
  const int base = FAST ? k.tile_base(w.start) : 0;
  const unsigned jaddr = (FAST && DEPOSIT) ? (unsigned)__cvta_generic_to_shared(k.Jt) + 4u * (unsigned)base : 0u;
#pragma unroll
  for (int cx = 0; cx < 3; ++cx) {
    const int cy = (cx + 1) % 3, cz = (cx + 2) % 3;
    double T[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) T[j][kk] = w.sn[cy][j] * (2 * w.sn[cz][kk] + w.s0[cz][kk]) + w.s0[cy][j] * (2 * w.s0[cz][kk] + w.sn[cz][kk]);
    double acc = 0.0;
    const double av = DEPOSIT ? alpha * v[cx] : 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          int d[3];
          d[cx] = i;
          d[cy] = j;
          d[cz] = kk;
          const double wt = w.sh1[cx][i] * T[j][kk];
          if (FAST) {
            const int e = cx * K::VOL + base + (d[2] * K::NY + d[1]) * K::NX + d[0];
            if (DEPOSIT)
              shared_add_fixed(jaddr + 4u * (unsigned)(e - base), jaddr + 4u * (unsigned)(e - base + 3 * K::VOL), __double2ll_rn(av * wt));
            else
              acc += k.Et[e] * wt;
          }
          else {
            const int64_t e = k.gidx(w.start[0] + d[0], w.start[1] + d[1], w.start[2] + d[2], cx);
            if (DEPOSIT)
              global_add_fixed(&k.J[e], __double2ll_rn(av * wt));
            else
              acc += __ldg(&k.E[e]) * wt;
          }
        }
    if (!DEPOSIT) Ep[cx] += acc;
  }
}

// B at the point r with the 2nd-order form factor: Shape::setup(r) (utils/shape.cpp:34-45) +
// Shape::magnetic (shape.h:66-73).  Each 1-D weight vector has three non-zero entries inside the
// reference's 3- or 4-point window: nodal from round(p) - 1, staggered from floor(p) - 1.
template <bool FAST, class K>
__device__ __forceinline__ void cap_gather_B(const K& k, const double* p, const int* lo_n, const int* lo_s, double* Bp)
{
  // inside these windows the middle point is always within half a cell (the 0.75 - x^2 branch of the spline)
  // and the outer two between 0.5 and 1.5 (the other branch, which reaches 0 at 1.5): no branches needed,
  // same values as spline_of_2nd_order
  double wn[3][3], ws[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double xn = fabs(p[a] - (double)(lo_n[a] + j));
      const double xs = fabs(p[a] - ((double)(lo_s[a] + j) + 0.5));
      wn[a][j] = j == 1 ? (0.75 - xn * xn) : 0.5 * (1.5 - xn) * (1.5 - xn);
      ws[a][j] = j == 1 ? (0.75 - xs * xs) : 0.5 * (1.5 - xs) * (1.5 - xs);
    }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // component c is nodal along c and staggered along the other two axes
    int lo[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) lo[a] = a == c ? lo_n[a] : lo_s[a];
    const int base = FAST ? k.tile_base(lo) : 0;
    double acc = 0.0;
#pragma unroll
    for (int kz = 0; kz < 3; ++kz)
#pragma unroll
      for (int jy = 0; jy < 3; ++jy)
#pragma unroll
        for (int ix = 0; ix < 3; ++ix) {
          const double wz = c == 2 ? wn[2][kz] : ws[2][kz];
          const double wy = c == 1 ? wn[1][jy] : ws[1][jy];
          const double wx = c == 0 ? wn[0][ix] : ws[0][ix];
          const double b = FAST ? k.Bt[c * K::VOL + base + (kz * K::NY + jy) * K::NX + ix] : __ldg(&k.B[k.gidx(lo[0] + ix, lo[1] + jy, lo[2] + kz, c)]);
          acc += b * (wz * wy * wx);
        }
    Bp[c] += acc;
  }
}

// out-of-tile paths (fast particles): rare, kept out of line; everything travels by value so that the
// in-tile path keeps its operands in registers
struct V3d {
  double x, y, z;
};

template <class K>
__device__ __noinline__ V3d cap_gather_B_slow(K k, double px, double py, double pz)
{
  const double p[3] = {px, py, pz};
  int lo_n[3], lo_s[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo_n[a] = (int)round(p[a]) - 1;
    lo_s[a] = (int)floor(p[a]) - 1;
  }
  double Bp[3] = {0.0, 0.0, 0.0};
  cap_gather_B<false>(k, p, lo_n, lo_s, Bp);
  return V3d{Bp[0], Bp[1], Bp[2]};
}

template <bool DEPOSIT, class K>
__device__ __noinline__ V3d cap_apply_slow(K k, double r0x, double r0y, double r0z, double rnx, double rny, double rnz, double alpha, double vx,
                                           double vy, double vz)
{
  const double rs0[3] = {r0x, r0y, r0z}, rsn[3] = {rnx, rny, rnz}, v[3] = {vx, vy, vz};
  CapW w;
  cap_weights(k.g, rsn, rs0, w);
  double Ep[3] = {0.0, 0.0, 0.0};
  cap_apply<false, DEPOSIT>(k, w, Ep, alpha, v);
  return V3d{Ep[0], Ep[1], Ep[2]};
}

// ImplicitEsirkepov::interpolate (implicit_esirkepov.cpp:63-91) for one segment: adds into Es, Bs
template <class K>
__device__ __forceinline__ void cap_interpolate(const K& k, const double* rsn, const double* rs0, double* Es, double* Bs)
{
  const double rh[3] = {0.5 * (rsn[0] + rs0[0]), 0.5 * (rsn[1] + rs0[1]), 0.5 * (rsn[2] + rs0[2])};
  double p[3];
  cells3(k.g, rh, p);
  int lo_n[3], lo_s[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo_n[a] = (int)round(p[a]) - 1;
    lo_s[a] = (int)floor(p[a]) - 1;
  }
  if (k.inside(lo_s, 4))
    cap_gather_B<true>(k, p, lo_n, lo_s, Bs);
  else {
    const V3d b = cap_gather_B_slow(k, p[0], p[1], p[2]);
    Bs[0] += b.x;
    Bs[1] += b.y;
    Bs[2] += b.z;
  }
  CapW w;
  cap_weights(k.g, rsn, rs0, w);
  if (k.inside(w.start, 3))
    cap_apply<true, false>(k, w, Es, 0.0, nullptr);
  else {
    const V3d e = cap_apply_slow<false>(k, rs0[0], rs0[1], rs0[2], rsn[0], rsn[1], rsn[2], 0.0, 0.0, 0.0, 0.0);
    Es[0] += e.x;
    Es[1] += e.y;
    Es[2] += e.z;
  }
}

// cell_traversal (cell_traversal.cpp:3-77): calls f(segment start, segment end) for every straight
// piece between crossings of the half-shifted cell faces; returns the number of pieces
template <class F>
__device__ __forceinline__ int for_each_segment(const Grid& g, const double* end, const double* start, F&& f)
{
  double ps[3], pe[3];
  cells3(g, start, ps);
  cells3(g, end, pe);
  int curr[3], last[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    curr[a] = (int)round(ps[a]);
    last[a] = (int)round(pe[a]);
  }
  if (curr[0] == last[0] && curr[1] == last[1] && curr[2] == last[2]) {
    f(start, end);
    return 1;
  }
  const double d3[3] = {g.dx, g.dy, g.dz};
  const double maxv = 1.7976931348623157e308;
  double dir[3], t3[3], dt3[3];
  int sg[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    dir[a] = end[a] - start[a];
    sg[a] = dir[a] > 0 ? 1 : -1;
    const double next = (curr[a] + sg[a] * 0.5) * d3[a];
    const double inv = (dir[a] != 0) ? 1.0 / dir[a] : 0.0;  // one division per axis (the reference divides twice)
    t3[a] = (dir[a] != 0) ? (next - start[a]) * inv : maxv;
    dt3[a] = (dir[a] != 0) ? d3[a] * inv * sg[a] : 0.0;
  }
  double prev[3] = {start[0], start[1], start[2]};
  int n = 0;
  while (!(curr[0] == last[0] && curr[1] == last[1] && curr[2] == last[2]) && n < 255) {
    int a;
    if (t3[0] < t3[1])
      a = (t3[0] < t3[2]) ? 0 : 2;
    else
      a = (t3[1] < t3[2]) ? 1 : 2;
    double t;
    // (no dynamic indexing of the small arrays)
    if (a == 0) { t = t3[0]; curr[0] += sg[0]; t3[0] += dt3[0]; }
    else if (a == 1) { t = t3[1]; curr[1] += sg[1]; t3[1] += dt3[1]; }
    else { t = t3[2]; curr[2] += sg[2]; t3[2] += dt3[2]; }
    const double pt[3] = {start[0] + dir[0] * t, start[1] + dir[1] * t, start[2] + dir[2] * t};
    f(prev, pt);
    prev[0] = pt[0];
    prev[1] = pt[1];
    prev[2] = pt[2];
    ++n;
  }
  f(prev, end);
  return n + 1;
}

// eccapfim::Particles::form_iteration for one particle (particles.cpp:73-176)
__device__ __forceinline__ void cap_push_particle(const CapCtx& k, const CapArgs& a, double* r, double* v, int& it_out, int& seg_out)
{
  const Grid& g = k.g;
  const double dt = g.dt, q = a.q, m = a.m;
  const double maxv = 1.7976931348623157e308;
  // :39-45 with the whole periodic box as the domain (start = 0, end = N on every axis)
  const double lo[3] = {(0 - 0.5) * g.dx, (0 - 0.5) * g.dy, (0 - 0.5) * g.dz};
  const double hi[3] = {(g.nx + 0.5) * g.dx, (g.ny + 0.5) * g.dy, (g.nz + 0.5) * g.dz};
  const double L[3] = {g.Lx, g.Ly, g.Lz};
  double r0[3] = {r[0], r[1], r[2]}, v0[3] = {v[0], v[1], v[2]};  // p0 (tmp); r, v are pn (curr)
  double tau = 0.0, dtau = 0.0;
  it_out = 0;
  seg_out = 0;
  for (; tau < dt; tau += dtau) {
    double vh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) vh[c] = 0.5 * (v[c] + v0[c]);
    dtau = dt - tau;
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // process_bound :49-56
      double tb = maxv;
      if (vh[c] > 0 && fabs(hi[c] - r0[c]) > 1e-7)
        tb = (hi[c] - r0[c]) / vh[c];
      else if (vh[c] < 0 && fabs(lo[c] - r0[c]) > 1e-7)
        tb = (lo[c] - r0[c]) / vh[c];
      dtau = fmin(dtau, tb);
    }
    const double a0 = q * a.mpw * a.inv_unit;  // the current in accumulator units (exact scaling)
    const double alpha = 0.5 * dtau * (q / m);
    double Ep[3], Bp[3];
    int nseg = 0;
    auto set_fields = [&]() {  // :109-127
      Ep[0] = Ep[1] = Ep[2] = Bp[0] = Bp[1] = Bp[2] = 0.0;
      const double d = len3(r[0] - r0[0], r[1] - r0[1], r[2] - r0[2]);
      nseg = for_each_segment(g, r, r0, [&](const double* rs0, const double* rsn) {
        const double ds = len3(rsn[0] - rs0[0], rsn[1] - rs0[1], rsn[2] - rs0[2]);
        const double bs = (d > 0 ? ds / d : 1.0);
        double Es[3] = {0.0, 0.0, 0.0}, Bs[3] = {0.0, 0.0, 0.0};
        cap_interpolate(k, rsn, rs0, Es, Bs);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Ep[c] += Es[c] * bs;
          Bp[c] += Bs[c] * bs;
        }
      });
    };
    auto residue = [&]() {  // :129-131
      double vxb[3];
      cross3(vh, Bp, vxb);
      const double f = dtau * q / m;
      return len3((v[0] - v0[0]) - f * (Ep[0] + vxb[0]), (v[1] - v0[1]) - f * (Ep[1] + vxb[1]), (v[2] - v0[2]) - f * (Ep[2] + vxb[2]));
    };
    set_fields();
    double rn = residue();
    const double rr0 = rn;
    int it = 0;
    for (; rn > a.cn_tol + a.cn_tol * rr0 && it < a.cn_maxit; ++it) {  // :136-148
      double aa[3], b[3], w[3], wxb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        aa[c] = alpha * Ep[c];
        b[c] = alpha * Bp[c];
        w[c] = v0[c] + aa[c];
      }
      cross3(w, b, wxb);
      const double wb = dot3(w, b), den = 1.0 + dot3(b, b);
#pragma unroll
      for (int c = 0; c < 3; ++c) vh[c] = ((w[c] + wxb[c]) + b[c] * wb) / den;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        r[c] = r0[c] + dtau * vh[c];
        v[c] = 2.0 * vh[c] - v0[c];
      }
      set_fields();
      rn = residue();
    }
    it_out += it;
    seg_out += nseg;
    {  // current of this sub-step, :153-163
      const double d = len3(r[0] - r0[0], r[1] - r0[1], r[2] - r0[2]);
      for_each_segment(g, r, r0, [&](const double* rs0, const double* rsn) {
        const double ds = len3(rsn[0] - rs0[0], rsn[1] - rs0[1], rsn[2] - rs0[2]);
        const double bs = (d > 0 ? ds / d : 1.0);
        CapW w;
        cap_weights(g, rsn, rs0, w);
        const double al = a0 * bs * (dtau / dt);
        if (k.inside(w.start, 3))
          cap_apply<true, true>(k, w, nullptr, al, vh);
        else
          cap_apply_slow<true>(k, rs0[0], rs0[1], rs0[2], rsn[0], rsn[1], rsn[2], al, vh[0], vh[1], vh[2]);
      });
    }
    bool reset = false;  // bound_periodic :58-68, :165-171
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (r[c] < 0.0) {
        r[c] = L[c] - (0.0 - r[c]);
        reset = true;
      }
      else if (r[c] > L[c]) {
        r[c] = 0.0 + (r[c] - L[c]);
        reset = true;
      }
    }
    if (reset) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        r0[c] = r[c];
        v0[c] = v[c];
      }
    }
  }
}

__global__ void __launch_bounds__(CAP_THREADS) k_cap_push(Grid g, CapArgs a)
{
  __shared__ double Et[3 * CAP_VOL], Bt[3 * CAP_VOL], Jt[3 * CAP_VOL];
  __shared__ unsigned long long cnt[2];
  const int gx = blockIdx.x % a.groups_x, row = blockIdx.x / a.groups_x;  // row = zl * ny + cy
  const int cy = row % g.ny, zl = row / g.ny;
  const int cx0 = gx * CAP_CELLS, ncell = min(CAP_CELLS, g.nx - cx0);
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0ull;
  for (int e = threadIdx.x; e < 3 * CAP_VOL; e += CAP_THREADS) {
    const int x = e % CAP_NX, y = (e / CAP_NX) % CAP_NY, z = (e / (CAP_NX * CAP_NY)) % CAP_NZ, c = e / CAP_VOL;
    const int64_t o = g.vidx(wrap_near(cx0 - CAP_LO + x, g.nx), wrap_near(cy - CAP_LO + y, g.ny), zl - CAP_LO + z, c);
    Et[e] = __ldg(&a.E[o]);
    Bt[e] = __ldg(&a.B[o]);
    Jt[e] = 0.0;
  }
  const int64_t cell0 = ((int64_t)(zl + 1) * g.ny + cy) * g.nx + cx0;  // bin plane = zl + 1
  const int32_t p0 = a.bin_start[cell0 << 3], p1 = a.bin_start[(cell0 + ncell) << 3];
  __syncthreads();
  CapCtx k{g, Et, Bt, Jt, cx0 - CAP_LO, cy - CAP_LO, g.z0 + zl - CAP_LO, a.E, a.B, a.J, a.error};
  unsigned its = 0, segs = 0;
  for (int32_t i = p0 + threadIdx.x; i < p1; i += CAP_THREADS) {
    double r[3] = {a.p0[0][i], a.p0[1][i], a.p0[2][i]};
    double v[3] = {a.p0[3][i], a.p0[4][i], a.p0[5][i]};
    int it, ns;
    cap_push_particle(k, a, r, v, it, ns);
    a.pn[0][i] = r[0];
    a.pn[1][i] = r[1];
    a.pn[2][i] = r[2];
    a.pn[3][i] = v[0];
    a.pn[4][i] = v[1];
    a.pn[5][i] = v[2];
    its += it;
    segs += ns;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    its += __shfl_down_sync(0xffffffffu, its, o);
    segs += __shfl_down_sync(0xffffffffu, segs, o);
  }
  if ((threadIdx.x & 31) == 0 && (its | segs)) {
    atomicAdd(&cnt[0], (unsigned long long)its);
    atomicAdd(&cnt[1], (unsigned long long)segs);
  }
  __syncthreads();
  if (threadIdx.x < 2 && cnt[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], cnt[threadIdx.x]);
  // flush the current tile: one reduction per touched node instead of 54 per particle segment
  for (int e = threadIdx.x; e < 3 * CAP_VOL; e += CAP_THREADS) {
    const long long val = tile_entry_fixed(Jt, 3 * CAP_VOL, e);
    if (val == 0) continue;
    if (llabs(val) >> 61) *a.error = 5;  // within a factor 4 of the accumulator range
    const int x = e % CAP_NX, y = (e / CAP_NX) % CAP_NY, z = (e / (CAP_NX * CAP_NY)) % CAP_NZ, c = e / CAP_VOL;
    global_add_fixed(&a.J[g.vidx(wrap_near(cx0 - CAP_LO + x, g.nx), wrap_near(cy - CAP_LO + y, g.ny), zl - CAP_LO + z, c)], val);
  }
}

// ---------------------------------------------------------------------------------------------
// k_cap_push_tasks: the same particle pass organised as a CTA-wide task machine.
//
// In k_cap_push a warp pays for the longest path (pieces) and the largest Picard count among its 32
// particles: 97 % of the warps contain a particle that crosses a face, so every field evaluation
// costs two piece evaluations, and the iteration count is the warp maximum.  Here CAP2_OWNERS
// threads each own one particle at a time (taken from a CTA work queue) and advance it through a
// small state machine.  Every round the owners EMIT the pieces of their present path as gather
// tasks into shared memory, all CAP2_THREADS threads PROCESS the tasks densely -- one task per
// lane, whatever particle it belongs to -- and the owners CONSUME the results in piece order, so
// the arithmetic per particle and its order are those of k_cap_push while the work is proportional
// to the pieces actually evaluated.  A converged particle appends its deposit tasks to a second
// queue that is drained in full warps whenever it holds a CTA's worth of tasks; nobody waits for a
// deposit, so the owner moves on to its next particle in the same round.
// ---------------------------------------------------------------------------------------------
constexpr int CAP2_CELLS = 16;
constexpr int CAP2_THREADS = 256;
#ifndef CAP2_OWNERS_V
#define CAP2_OWNERS_V 224
#endif
constexpr int CAP2_OWNERS = CAP2_OWNERS_V;   // ~1.1 pieces per particle => ~246 gather tasks per round: one dense pass
constexpr int CAP2_GCAP = 272;     // gather tasks per round (two buffers: consume one while the next fills)
constexpr int CAP2_GSTRIDE = 7;    // rs0[3], rsn[3], bs  -> Es[3], Bs[3], bs
constexpr int CAP2_DCAP = 320;     // ring of queued deposit tasks
constexpr int CAP2_DSTRIDE = 10;   // rs0[3], rsn[3], al, vh[3]
constexpr int CAP2_MAXSEG = 24;
using CapCtx2 = CapCtxT<CAP2_CELLS>;
constexpr int CAP2_SMEM_DOUBLES = 3 * 3 * CapCtx2::VOL + 2 * CAP2_GCAP * CAP2_GSTRIDE + CAP2_DCAP * CAP2_DSTRIDE;

enum { CS_FETCH = 0, CS_FIELDS = 1, CS_WAIT = 2, CS_DEPOSIT = 3, CS_IDLE = 4 };

// Reservation of n consecutive slots with one atomicAdd (a CAS loop over 224 owners is quadratic).  The
// counter only grows within a round; a reservation that crosses `limit` fails, and its owner neutralises
// the part of its range that lies below the limit (the pass over the queue would otherwise read stale
// slots there).  The round's uniform bookkeeping clamps the counter back to the limit.
__global__ void __launch_bounds__(CAP2_THREADS, 2) k_cap_push_tasks(Grid g, CapArgs a)
{
  extern __shared__ double smem[];
  constexpr int VOL3 = 3 * CapCtx2::VOL;
  double* Et = smem;
  double* Bt = Et + VOL3;
  double* Jt = Bt + VOL3;
  double* gq0 = Jt + VOL3;                           // gather tasks, two buffers
  double* dq = gq0 + 2 * CAP2_GCAP * CAP2_GSTRIDE;   // deposit ring
  __shared__ int q_next, n_gather[2], d_tail;
  __shared__ unsigned long long cnt[2];
  const int tid = threadIdx.x;
  const int gx = blockIdx.x % a.groups_x, row = blockIdx.x / a.groups_x;  // row = zl * ny + cy
  const int cy = row % g.ny, zl = row / g.ny;
  const int cx0 = gx * CAP2_CELLS, ncell = min(CAP2_CELLS, g.nx - cx0);
  const int64_t cell0 = ((int64_t)(zl + 1) * g.ny + cy) * g.nx + cx0;  // bin plane = zl + 1
  const int32_t p0 = a.bin_start[cell0 << 3], p1 = a.bin_start[(cell0 + ncell) << 3];
  if (p0 == p1) return;
  if (tid == 0) {
    q_next = p0;
    n_gather[0] = n_gather[1] = 0;
    d_tail = 0;
    cnt[0] = cnt[1] = 0ull;
  }
  for (int e = tid; e < VOL3; e += CAP2_THREADS) {
    const int x = e % CapCtx2::NX, y = (e / CapCtx2::NX) % CapCtx2::NY, z = (e / (CapCtx2::NX * CapCtx2::NY)) % CapCtx2::NZ, c = e / CapCtx2::VOL;
    const int64_t o = g.vidx(wrap_near(cx0 - CAP_LO + x, g.nx), wrap_near(cy - CAP_LO + y, g.ny), zl - CAP_LO + z, c);
    Et[e] = __ldg(&a.E[o]);
    Bt[e] = __ldg(&a.B[o]);
    Jt[e] = 0.0;
  }
  __syncthreads();
  const CapCtx2 k{g, Et, Bt, Jt, cx0 - CAP_LO, cy - CAP_LO, g.z0 + zl - CAP_LO, a.E, a.B, a.J, a.error};

  const double dt = g.dt, qm = a.q / a.m, a0 = a.q * a.mpw * a.inv_unit;  // currents in accumulator units

  // the particle this thread owns: start state (r0, v0), mean velocity vh of the present Picard iterate,
  // from which r = r0 + dtau vh, v = 2 vh - v0 once the particle has been moved (particles.cpp:143-144)
  int state = tid < CAP2_OWNERS ? CS_FETCH : CS_IDLE;
  int idx = 0, it = 0, base = 0, nseg = 0;
  bool first = true, moved = false;
  double r0[3], v0[3], vh[3];
  double tau = 0.0, dtau = 0.0, rr0 = 0.0;
  unsigned its = 0, segs = 0;
  int d_head = 0;  // the same value in every thread (advanced by the uniform drain decision)
  int cur = 0;     // gather buffer whose results are being consumed; the other one is being filled

  auto position = [&](double* r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = moved ? r0[c] + dtau * vh[c] : r0[c];
  };
  auto start_substep = [&]() {  // particles.cpp:84-101; the domain is the whole periodic box (:39-45)
    const double maxv = 1.7976931348623157e308;
    const double lo[3] = {(0 - 0.5) * g.dx, (0 - 0.5) * g.dy, (0 - 0.5) * g.dz};
    const double hi[3] = {(g.nx + 0.5) * g.dx, (g.ny + 0.5) * g.dy, (g.nz + 0.5) * g.dz};
#pragma unroll
    for (int c = 0; c < 3; ++c) vh[c] = v0[c];  // 0.5 (v + v0) with v = v0 at the start of a sub-step
    dtau = dt - tau;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double tb = maxv;
      if (vh[c] > 0 && fabs(hi[c] - r0[c]) > 1e-7)
        tb = (hi[c] - r0[c]) / vh[c];
      else if (vh[c] < 0 && fabs(lo[c] - r0[c]) > 1e-7)
        tb = (lo[c] - r0[c]) / vh[c];
      dtau = fmin(dtau, tb);
    }
    it = 0;
    first = true;
    moved = false;
  };
  // number of straight pieces of r0 -> r: one more than the faces of the half-shifted lattice crossed
  auto count_pieces = [&](const double* r) {
    double ps[3], pe[3];
    cells3(g, r0, ps);
    cells3(g, r, pe);
    int n = 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) n += abs((int)round(pe[c]) - (int)round(ps[c]));
    if (n > CAP2_MAXSEG) {
      *a.error = 3;  // more than CAP2_MAXSEG - 1 cell faces crossed in one step
      n = CAP2_MAXSEG;
    }
    return n;
  };
  // pieces of r0 -> r into `n` queue slots starting at first_slot (modulo `cap`); slots the walk does not
  // fill are neutralised; returns the pieces written
  auto emit = [&](double* queue, int stride, int cap, int first_slot, int n, const double* r, bool dep) {
    const double d = len3(r[0] - r0[0], r[1] - r0[1], r[2] - r0[2]);
    int kk = 0;
    auto put = [&](const double* rs0, const double* rsn, double wgt) {
      int slot = first_slot + kk;
      if (slot >= cap) slot -= cap;
      double* t = queue + slot * stride;
      t[0] = rs0[0]; t[1] = rs0[1]; t[2] = rs0[2];
      t[3] = rsn[0]; t[4] = rsn[1]; t[5] = rsn[2];
      t[6] = wgt;
      if (dep) {
        t[7] = vh[0]; t[8] = vh[1]; t[9] = vh[2];
      }
      ++kk;
    };
    for_each_segment(g, r, r0, [&](const double* rs0, const double* rsn) {
      if (kk >= n) return;
      const double ds = len3(rsn[0] - rs0[0], rsn[1] - rs0[1], rsn[2] - rs0[2]);
      const double bs = (d > 0 ? ds / d : 1.0);
      put(rs0, rsn, dep ? a0 * bs * (dtau / dt) : bs);
    });
    const int filled = kk;
    while (kk < n) put(r0, r0, 0.0);
    return filled;
  };
  auto deposit_task = [&](const double* tk) {
    const double rs0[3] = {tk[0], tk[1], tk[2]}, rsn[3] = {tk[3], tk[4], tk[5]};
    const double al = tk[6], vv[3] = {tk[7], tk[8], tk[9]};
    CapW w;
    cap_weights(g, rsn, rs0, w);
    if (k.inside(w.start, 3))
      cap_apply<true, true>(k, w, nullptr, al, vv);
    else
      cap_apply_slow<true>(k, rs0[0], rs0[1], rs0[2], rsn[0], rsn[1], rsn[2], al, vv[0], vv[1], vv[2]);
  };

  while (true) {
    double* gq_cur = gq0 + cur * (CAP2_GCAP * CAP2_GSTRIDE);
    double* gq_nxt = gq0 + (cur ^ 1) * (CAP2_GCAP * CAP2_GSTRIDE);
    // ---- owners: consume the results of the last pass, then put up the next request ------------------
    if (state == CS_WAIT) {
      double Ep[3] = {0.0, 0.0, 0.0}, Bp[3] = {0.0, 0.0, 0.0};
      for (int s = 0; s < nseg; ++s) {
        const double* tk = gq_cur + (base + s) * CAP2_GSTRIDE;
        const double bs = tk[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Ep[c] += tk[c] * bs;
          Bp[c] += tk[3 + c] * bs;
        }
      }
      double vxb[3];
      cross3(vh, Bp, vxb);
      const double f = dtau * qm;
      // v - v0 = 2 (vh - v0) once moved, 0 before
      double dv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) dv[c] = moved ? (2.0 * vh[c] - v0[c]) - v0[c] : 0.0;
      const double rn = len3(dv[0] - f * (Ep[0] + vxb[0]), dv[1] - f * (Ep[1] + vxb[1]), dv[2] - f * (Ep[2] + vxb[2]));
      if (first) {
        rr0 = rn;
        first = false;
      }
      if (rn > a.cn_tol + a.cn_tol * rr0 && it < a.cn_maxit) {  // particles.cpp:136-148
        const double alpha = 0.5 * dtau * qm;
        double aa[3], b[3], w[3], wxb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          aa[c] = alpha * Ep[c];
          b[c] = alpha * Bp[c];
          w[c] = v0[c] + aa[c];
        }
        cross3(w, b, wxb);
        const double wb = dot3(w, b), den = 1.0 + dot3(b, b);
#pragma unroll
        for (int c = 0; c < 3; ++c) vh[c] = ((w[c] + wxb[c]) + b[c] * wb) / den;
        moved = true;
        ++it;
        state = CS_FIELDS;
      }
      else {
        its += it;
        segs += nseg;
        state = CS_DEPOSIT;
      }
    }
    if (state == CS_DEPOSIT) {
      // queue the current of this sub-step (:153-163); nobody waits for it, so finish the sub-step at once
      double r[3];
      position(r);
      const int n = count_pieces(r);
      const int mine = atomicAdd(&d_tail, n), dlimit = d_head + CAP2_DCAP;
      if (mine + n > dlimit) {
        for (int s = mine; s < min(mine + n, dlimit); ++s) {  // stays in CS_DEPOSIT; a drain follows this round
          double* t = dq + (s % CAP2_DCAP) * CAP2_DSTRIDE;
          t[0] = t[3] = r0[0]; t[1] = t[4] = r0[1]; t[2] = t[5] = r0[2];
          t[6] = t[7] = t[8] = t[9] = 0.0;
        }
      }
      else {
        emit(dq, CAP2_DSTRIDE, CAP2_DCAP, mine % CAP2_DCAP, n, r, true);
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = moved ? 2.0 * vh[c] - v0[c] : v0[c];
        bool reset = false;  // bound_periodic, particles.cpp:58-68,165-171
        const double L[3] = {g.Lx, g.Ly, g.Lz};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (r[c] < 0.0) {
            r[c] = L[c] - (0.0 - r[c]);
            reset = true;
          }
          else if (r[c] > L[c]) {
            r[c] = 0.0 + (r[c] - L[c]);
            reset = true;
          }
        }
        tau += dtau;
        if (tau < dt) {
          // the sub-step ended on the box edge + 1/2 cell, i.e. outside the box: the wrap always resets the
          // start state (:169-171); anything else would need the unreset (r, v) pair this kernel does not keep
          if (!reset) *a.error = 4;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            r0[c] = r[c];
            v0[c] = v[c];
          }
          start_substep();
          state = CS_FIELDS;
        }
        else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            a.pn[c][idx] = r[c];
            a.pn[3 + c][idx] = v[c];
          }
          state = CS_FETCH;
        }
      }
    }
    if (state == CS_FETCH) {
      idx = atomicAdd(&q_next, 1);
      if (idx < p1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          r0[c] = a.p0[c][idx];
          v0[c] = a.p0[3 + c][idx];
        }
        tau = 0.0;
        start_substep();
        if (a.warm && dtau == dt) {
          // Picard starts from the mean velocity the previous residual evaluation of this step converged to
          // (the reference restarts from v0 every time, particles.cpp:77-78; the fixed point is the same)
#pragma unroll
          for (int c = 0; c < 3; ++c) vh[c] = 0.5 * (a.pn[3 + c][idx] + v0[c]);
          moved = true;
        }
        state = CS_FIELDS;
      }
      else
        state = CS_IDLE;
    }
    if (state == CS_FIELDS) {
      double r[3];
      position(r);
      const int n = count_pieces(r);
      const int mine = atomicAdd(&n_gather[cur ^ 1], n);
      if (mine + n <= CAP2_GCAP) {
        base = mine;
        nseg = emit(gq_nxt, CAP2_GSTRIDE, CAP2_GCAP, base, n, r, false);
        state = CS_WAIT;
      }
      else {
        for (int s = mine; s < min(mine + n, CAP2_GCAP); ++s) {  // stays in CS_FIELDS, tries again next round
          double* t = gq_nxt + s * CAP2_GSTRIDE;
          t[0] = t[3] = r0[0]; t[1] = t[4] = r0[1]; t[2] = t[5] = r0[2];
          t[6] = 0.0;
        }
      }
    }
    const bool busy = __syncthreads_or(state != CS_IDLE) != 0;
    // ---- all threads: one gather task per lane; deposits in full passes (everything once no particle is left)
    {
      const int ng = min(n_gather[cur ^ 1], CAP2_GCAP);
      for (int t = tid; t < ng; t += CAP2_THREADS) {
        double* tk = gq_nxt + t * CAP2_GSTRIDE;
        const double rs0[3] = {tk[0], tk[1], tk[2]}, rsn[3] = {tk[3], tk[4], tk[5]};
        double Es[3] = {0.0, 0.0, 0.0}, Bs[3] = {0.0, 0.0, 0.0};
        cap_interpolate(k, rsn, rs0, Es, Bs);
        tk[0] = Es[0]; tk[1] = Es[1]; tk[2] = Es[2];
        tk[3] = Bs[0]; tk[4] = Bs[1]; tk[5] = Bs[2];
      }
      const int avail = min(d_tail, d_head + CAP2_DCAP) - d_head;  // failed reservations overshoot the limit
      const int take = busy ? (avail / CAP2_THREADS) * CAP2_THREADS : avail;
      // lane l of warp w takes tasks 8 l + w (+ 256 j): the tasks of one warp instruction are 8 apart in the
      // queue, i.e. mostly particles of different half cells, so their shared-memory additions rarely
      // collide (consecutive tasks share a 54-node window and would serialise the CAS loops)
      for (int t = (tid & 31) * (CAP2_THREADS / 32) + (tid >> 5); t < take; t += CAP2_THREADS)
        deposit_task(dq + ((d_head + t) % CAP2_DCAP) * CAP2_DSTRIDE);
      if (tid == 0) {
        n_gather[cur] = 0;  // that buffer was consumed above; it is filled again in the next round
        if (d_tail > d_head + CAP2_DCAP) d_tail = d_head + CAP2_DCAP;  // idempotent under the min() above
      }
      d_head += take;
    }
    if (!busy) break;
    cur ^= 1;
    __syncthreads();
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    its += __shfl_down_sync(0xffffffffu, its, o);
    segs += __shfl_down_sync(0xffffffffu, segs, o);
  }
  if ((tid & 31) == 0 && (its | segs)) {
    atomicAdd(&cnt[0], (unsigned long long)its);
    atomicAdd(&cnt[1], (unsigned long long)segs);
  }
  __syncthreads();
  if (tid < 2 && cnt[tid]) atomicAdd(&a.counters[tid], cnt[tid]);
  for (int e = tid; e < VOL3; e += CAP2_THREADS) {
    const long long val = tile_entry_fixed(Jt, VOL3, e);
    if (val == 0) continue;
    if (llabs(val) >> 61) *a.error = 5;  // within a factor 4 of the accumulator range
    const int x = e % CapCtx2::NX, y = (e / CapCtx2::NX) % CapCtx2::NY, z = (e / (CapCtx2::NX * CapCtx2::NY)) % CapCtx2::NZ, c = e / CapCtx2::VOL;
    global_add_fixed(&a.J[g.vidx(wrap_near(cx0 - CAP_LO + x, g.nx), wrap_near(cy - CAP_LO + y, g.ny), zl - CAP_LO + z, c)], val);
  }
}

// max over the particles of |v_c| as the bit pattern of a non-negative double (integer order == numeric order)
__global__ void __launch_bounds__(256) k_cap_max_speed(int64_t n, const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                                                      unsigned long long* __restrict__ out)
{
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, fmax(fabs(vx[i]), fmax(fabs(vy[i]), fabs(vz[i]))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// fixed-point accumulators -> the current, in place (one rounding per entry)
__global__ void __launch_bounds__(256) k_cap_current_from_fixed(double* __restrict__ J, int64_t n, double unit, int* __restrict__ error)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long acc = reinterpret_cast<const long long*>(J)[i];
  if (llabs(acc) >> 62) *error = 5;
  J[i] = (double)acc * unit;
}

// F = x + dt^2/4 curl^- curl^+ x - rhs0 + dt/2 J        (form_function, simulation.cpp:228-236;
// rhs0 = E^n + dt/2 curl^- B^n is constant during the step)
__global__ void __launch_bounds__(256) k_cap_function(Grid g, const double* __restrict__ xk, const double* __restrict__ rhs0, const double* __restrict__ J,
                                                     double* __restrict__ F)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.ncl) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = (int)(node / g.plane);
  const int xm = x == 0 ? g.nx - 1 : x - 1, xp = x == g.nx - 1 ? 0 : x + 1;
  const int ym = y == 0 ? g.ny - 1 : y - 1, yp = y == g.ny - 1 ? 0 : y + 1;
  auto f = [&](int comp, int ox, int oy, int oz) {
    const int xx = ox < 0 ? xm : (ox > 0 ? xp : x), yy = oy < 0 ? ym : (oy > 0 ? yp : y);
    return __ldg(&xk[g.vidx(xx, yy, zl + oz, comp)]);
  };
  const double inv_d[3] = {1.0 / g.dx, 1.0 / g.dy, 1.0 / g.dz};
  const double h = 0.25 * g.dt * g.dt;
  const int64_t o = g.vidx(x, y, zl, 0);
#pragma unroll
  for (int c = 0; c < 3; ++c) F[o + c] = ((f(c, 0, 0, 0) + h * curlcurl(c, inv_d, f)) - rhs0[o + c]) + (0.5 * g.dt) * J[o + c];
}

static inline int grid_for(int64_t n, int threads = 256)
{
  int64_t b = (n + threads - 1) / threads;
  return (int)(b < 1 ? 1 : b);
}

int cap_alloc(xb_ctx* c)
{
  if (c->cap_x) return 0;
  for (double** v : {&c->cap_x, &c->cap_F, &c->cap_g, &c->cap_rhs0, &c->cap_J}) {
    XB_CUDA(cudaMalloc(v, sizeof(double) * c->g.ntot));
    XB_CUDA(cudaMemsetAsync(*v, 0, sizeof(double) * c->g.ntot, c->stream));
  }
  XB_CUDA(cudaMalloc(&c->cap_counters, 4 * sizeof(unsigned long long)));
  return 0;
}

// form_iteration (simulation.cpp:132-155): J(x) from all sorts, then F(x).  x must be a ghosted vector.
int cap_form_function(xb_ctx* c, double* x, double* F)
{
  const Grid& g = c->g;
  Nonlinear& nl = c->nl;
  XB_CHECK(cap_alloc(c));
  XB_CHECK(halo_fill(c, x, GZ));
  XB_CHECK(vec_zero(c, c->cap_J));
  XB_CUDA(cudaMemsetAsync(c->cap_counters, 0, 4 * sizeof(unsigned long long), c->stream));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (nl.profile) {
    if (nl.events.size() < nl.events_used + 2) {
      nl.events.resize(nl.events_used + 2, nullptr);
      XB_CUDA(cudaEventCreate(&nl.events[nl.events_used]));
      XB_CUDA(cudaEventCreate(&nl.events[nl.events_used + 1]));
    }
    e0 = nl.events[nl.events_used];
    e1 = nl.events[nl.events_used + 1];
    nl.events_used += 2;
    XB_CUDA(cudaEventRecord(e0, c->stream));
  }
  int64_t total = 0;
  for (auto& s : c->sorts) {
    if (!s.sorted) XB_FAIL("eccapfim: particles are not sorted");
    XB_CHECK(vec_zero(c, s.currI));  // Particles::J of this evaluation (clear_sources, particles.cpp:183-189)
    total += s.count;
    // a slab that holds no particle of this sort skips only the particle kernel: the halo add below is a
    // collective with both z neighbours (their currents land in this rank's boundary planes)
    const bool have_particles = s.count > 0;
    CapArgs a;
    for (int k = 0; k < 6; ++k) {
      a.p0[k] = s.p[s.cur][k];
      a.pn[k] = s.p[1 - s.cur][k];
    }
    a.bin_start = s.bin_start;
    a.E = x;
    a.B = c->B;
    a.J = s.currI;
    a.q = s.q;
    a.m = s.m;
    a.mpw = s.n / (double)s.Np;
    a.cn_tol = nl.cn_tol;
    a.cn_maxit = nl.cn_maxit;
    a.counters = c->cap_counters;
    a.error = reinterpret_cast<int*>(c->cap_counters + 2);
    a.warm = (nl.warm_start && nl.pn_valid) ? 1 : 0;
    const double unit = s.cap_unit;  // set once per step by cap_prepare
    a.inv_unit = 1.0 / unit;
    if (!have_particles) {
    }
    else if (c->cap_variant == 1) {  // thread-per-particle kernel, kept as a cross-check
      a.groups_x = (g.nx + CAP_CELLS - 1) / CAP_CELLS;
      const int64_t blocks = (int64_t)a.groups_x * g.ny * g.nzl;
      XB_LAUNCH(c, k_cap_push, (int)blocks, CAP_THREADS, 0, g, a);
    }
    else {
      const size_t smem = sizeof(double) * CAP2_SMEM_DOUBLES;
      if (!c->cap_attr_set) {
        XB_CUDA(cudaFuncSetAttribute(k_cap_push_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->cap_attr_set = true;
      }
      a.groups_x = (g.nx + CAP2_CELLS - 1) / CAP2_CELLS;
      const int64_t blocks = (int64_t)a.groups_x * g.ny * g.nzl;
      XB_LAUNCH(c, k_cap_push_tasks, (int)blocks, CAP2_THREADS, smem, g, a);
    }
    if (have_particles) XB_LAUNCH(c, k_cap_current_from_fixed, grid_for(g.ntot), 256, 0, s.currI, g.ntot, unit, a.error);
    XB_CHECK(halo_reduce(c, s.currI, GZ, GZ));  // DMLocalToGlobal(ADD), particles.cpp:179
    const double one = 1.0;
    const double* vs[1] = {s.currI};
    XB_CHECK(axpy_multi(c, 1, vs, &one, c->cap_J));  // VecAXPY(J, 1, sort->J), simulation.cpp:188
  }
  if (nl.profile) XB_CUDA(cudaEventRecord(e1, c->stream));
  XB_LAUNCH(c, k_cap_function, grid_for(g.ncl), 256, 0, g, x, c->cap_rhs0, c->cap_J, F);
  ++nl.fevals;
  nl.particles_per_eval = total;
  nl.pn_valid = true;
  return 0;
}

int cap_read_counters(xb_ctx* c)
{
  Nonlinear& nl = c->nl;
  unsigned long long h[4];
  XB_CUDA(cudaMemcpyAsync(h, c->cap_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  const int err = (int)(h[2] & 0xffffffffu);
  if (err == 2) XB_FAIL("eccapfim: a particle moved beyond the ghost planes of its slab within one step");
  if (err == 3) XB_FAIL("eccapfim: a particle crossed more cell faces in one step than the path splitter holds (CAP2_MAXSEG)");
  if (err == 4) XB_FAIL("eccapfim: a time-split sub-step ended inside the box (only the split at the box edge is supported)");
  if (err == 5) XB_FAIL("eccapfim: the fixed-point current accumulators are close to their range (thousands of particles per node at twice the largest start-of-step speed)");
  if (err) XB_FAIL("eccapfim: particle pass failed with device error code " + std::to_string(err));
  double sums[2] = {(double)h[0], (double)h[1]};
  double n = (double)nl.particles_per_eval;
  if (c->g.nranks > 1) {
    c->red_host[0] = sums[0];
    c->red_host[1] = sums[1];
    c->red_host[2] = n;
    XB_CUDA(cudaMemcpyAsync(c->red_out, c->red_host, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    XB_CHECK(comm_allreduce_sum(c, c->red_out, 3));
    XB_CUDA(cudaMemcpyAsync(c->red_host, c->red_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    sums[0] = c->red_host[0];
    sums[1] = c->red_host[1];
    n = c->red_host[2];
  }
  nl.avg_cn = n > 0 ? sums[0] / n : 0.0;
  nl.avg_cells = n > 0 ? sums[1] / n : 0.0;
  return 0;
}

// rhs0 = E^n + dt/2 curl^- B^n
int cap_prepare(xb_ctx* c)
{
  XB_CHECK(cap_alloc(c));
  for (auto& s : c->sorts)
    if (!s.sorted) {
      if (c->g.nranks > 1) XB_CHECK(migrate_and_sort(c, s, 0.0));
      else XB_CHECK(particles_sort(c, s, 0.0));
    }
  // Unit of the fixed-point current accumulators of every sort for this step: 2^-49 of (the power of two below
  // |q n / Np|) x (the power of two above twice the largest velocity component at the start of the step).  An
  // accumulator holds 2^12 such products before the range check where the tiles are flushed fires -- in units of
  // particles per node: more than 8000 at the largest speed, typically ten times that.
  {
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(c->red_out);
    const size_t ns = c->sorts.size();
    if (ns > (size_t)RED_MAXV) XB_FAIL("eccapfim: too many sorts");
    XB_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned long long) * ns, c->stream));
    for (size_t i = 0; i < ns; ++i) {
      Species& s = c->sorts[i];
      if (s.count > 0) XB_LAUNCH(c, k_cap_max_speed, RED_BLOCKS, 256, 0, s.count, s.p[s.cur][3], s.p[s.cur][4], s.p[s.cur][5], slot + i);
    }
    XB_CUDA(cudaMemcpyAsync(c->red_host, slot, sizeof(unsigned long long) * ns, cudaMemcpyDeviceToHost, c->stream));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < ns; ++i) {
      Species& s = c->sorts[i];
      double vmax;
      std::memcpy(&vmax, reinterpret_cast<const char*>(c->red_host) + sizeof(double) * i, sizeof(double));
      const double flux = std::fabs(s.q * (s.n / (double)s.Np));
      const int ev = vmax > 0.0 ? std::ilogb(vmax) + 2 : 0;  // 2^ev > 2 vmax
      s.cap_unit = flux > 0.0 ? std::ldexp(1.0, std::ilogb(flux) + ev - 49) : 1.0;
    }
  }
  c->nl.pn_valid = false;  // the first evaluation of a step starts every particle from (r0, v0)
  XB_CHECK(halo_fill(c, c->B, GZ));  // DMGlobalToLocal(B), simulation.cpp:64
  XB_CHECK(vec_copy_owned(c, c->E, c->cap_rhs0));
  XB_CHECK(curl_apply(c, false, c->B, c->cap_rhs0, 0.5 * c->g.dt, true));
  return 0;
}

// symmetric least squares  G gamma = b  through a Jacobi eigen-decomposition; eigenvalues below
// cut * max dropped (G is the Gram matrix of the Anderson differences, possibly near-singular)
static void solve_gram(int m, std::vector<double> G, const std::vector<double>& b, std::vector<double>& gamma, double cut)
{
  std::vector<double> V((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) V[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) (i == j ? diag : off) += G[(size_t)i * m + j] * G[(size_t)i * m + j];
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = G[(size_t)p * m + q];
        if (apq == 0.0) continue;
        const double theta = (G[(size_t)q * m + q] - G[(size_t)p * m + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < m; ++k) {  // columns p, q
          const double gkp = G[(size_t)k * m + p], gkq = G[(size_t)k * m + q];
          G[(size_t)k * m + p] = cs * gkp - sn * gkq;
          G[(size_t)k * m + q] = sn * gkp + cs * gkq;
        }
        for (int k = 0; k < m; ++k) {  // rows p, q
          const double gpk = G[(size_t)p * m + k], gqk = G[(size_t)q * m + k];
          G[(size_t)p * m + k] = cs * gpk - sn * gqk;
          G[(size_t)q * m + k] = sn * gpk + cs * gqk;
        }
        for (int k = 0; k < m; ++k) {
          const double vkp = V[(size_t)k * m + p], vkq = V[(size_t)k * m + q];
          V[(size_t)k * m + p] = cs * vkp - sn * vkq;
          V[(size_t)k * m + q] = sn * vkp + cs * vkq;
        }
      }
  }
  double lmax = 0.0;
  for (int i = 0; i < m; ++i) lmax = std::max(lmax, G[(size_t)i * m + i]);
  gamma.assign(m, 0.0);
  for (int j = 0; j < m; ++j) {
    const double lam = G[(size_t)j * m + j];
    if (!(lam > cut * lmax)) continue;
    double vb = 0.0;
    for (int i = 0; i < m; ++i) vb += V[(size_t)i * m + j] * b[i];
    const double y = vb / lam;
    for (int i = 0; i < m; ++i) gamma[i] += V[(size_t)i * m + j] * y;
  }
}

static double norm_of(xb_ctx* c, const double* v, int* rc)
{
  const double* vs[1] = {v};
  double n2 = 0.0;
  *rc = dots(c, 1, vs, v, &n2);
  return std::sqrt(n2);
}

// g = -P F,  P = ((1 + sigma) I + dt^2/4 curl curl)^-1 = 2 ((2 + 2 sigma) I + dt^2/2 curl curl)^-1
static int cap_precondition(xb_ctx* c, const double* F, double* gout)
{
  Nonlinear& nl = c->nl;
  if (nl.cheb_degree <= 0) return scale_into(c, F, -1.0, gout);
  double sigma = 0.0;
  for (auto& s : c->sorts) sigma += 0.25 * c->g.dt * c->g.dt * s.q * s.q * s.n / s.m;
  XB_CHECK(cheb_solve_shifted(c, nl.cheb_degree, 2.0 + 2.0 * sigma, F, gout));
  return scale_into(c, gout, -2.0, gout);
}

// calc_iteration (simulation.cpp:72-104): solve F(x) = 0 from x = E^n by Anderson acceleration of
// x <- x + g(x), g = -P F.  Ring of `slots` difference pairs (dX_i = x_{i+1} - x_i, dG_i = g_{i+1} - g_i):
// up to slots - 1 complete ("live") pairs plus the pending one whose dG is closed by the next evaluation.
int cap_solve(xb_ctx* c)
{
  Nonlinear& nl = c->nl;
  const int slots = std::max(2, std::min(nl.depth + 1, (int)(c->V.size() / 2)));
  double** dX = c->V.data();
  double** dG = c->V.data() + slots;
  nl.hist.clear();
  nl.iterations = 0;
  nl.fevals = 0;
  nl.reason = 0;
  nl.events_used = 0;
  double *x = c->cap_x, *F = c->cap_F, *gk = c->cap_g;
  XB_CHECK(vec_copy_owned(c, c->E, x));  // initial guess E^{n+1/2,0} = E^n (:58-63)
  XB_CHECK(cap_form_function(c, x, F));
  int rc = 0;
  double fnorm = norm_of(c, F, &rc);
  XB_CHECK(rc);
  nl.hist.push_back(fnorm);
  const double ttol = fnorm * nl.rtol;
  auto converged = [&](double fn) {  // SNESConvergedDefault with snorm = xnorm = 0
    if (!(fn == fn)) return -4;      // SNES_DIVERGED_FNORM_NAN
    if (fn < nl.atol) return 2;      // SNES_CONVERGED_FNORM_ABS
    if (fn <= ttol) return 3;        // SNES_CONVERGED_FNORM_RELATIVE
    return 0;
  };
  nl.reason = converged(fnorm);
  std::vector<double> G((size_t)slots * slots, 0.0), row(slots), gamma;
  int mk = 0, head = 0;  // live pairs; the pending slot
  bool pending = false;
  const double one = 1.0;
  for (int k = 1; !nl.reason && k <= nl.maxit; ++k) {
    XB_CHECK(cap_precondition(c, F, gk));
    if (pending) {
      // close the pending pair: dG[last] held -g_{k-1}
      const int last = (head + slots - 1) % slots;
      const double* vs[1] = {gk};
      XB_CHECK(axpy_multi(c, 1, vs, &one, dG[last]));
      if (mk < slots - 1) ++mk;
      std::vector<const double*> ptr(mk);
      for (int j = 0; j < mk; ++j) ptr[j] = dG[(head + slots - 1 - j) % slots];
      XB_CHECK(dots(c, mk, ptr.data(), dG[last], row.data()));
      for (int j = 0; j < mk; ++j) {
        const int i = (head + slots - 1 - j) % slots;
        G[(size_t)i * slots + last] = G[(size_t)last * slots + i] = row[j];
      }
    }
    double* step = dX[head];  // becomes x_{k+1} - x_k
    XB_CHECK(vec_copy_owned(c, gk, step));
    if (mk > 0) {
      std::vector<const double*> ptr(mk);
      std::vector<int> live(mk);
      for (int j = 0; j < mk; ++j) {
        live[j] = (head + slots - 1 - j) % slots;
        ptr[j] = dG[live[j]];
      }
      std::vector<double> bs(mk), Gs((size_t)mk * mk);
      XB_CHECK(dots(c, mk, ptr.data(), gk, bs.data()));
      for (int i = 0; i < mk; ++i)
        for (int j = 0; j < mk; ++j) Gs[(size_t)i * mk + j] = G[(size_t)live[i] * slots + live[j]];
      solve_gram(mk, Gs, bs, gamma, 1e-14);
      std::vector<const double*> vs;
      std::vector<double> cf;
      for (int j = 0; j < mk; ++j) {
        vs.push_back(dX[live[j]]);
        cf.push_back(-gamma[j]);
        vs.push_back(dG[live[j]]);
        cf.push_back(-gamma[j]);
      }
      XB_CHECK(axpy_multi(c, (int)vs.size(), vs.data(), cf.data(), step));
    }
    XB_CHECK(scale_into(c, gk, -1.0, dG[head]));
    pending = true;
    {
      const double* vs[1] = {step};
      XB_CHECK(axpy_multi(c, 1, vs, &one, x));
    }
    head = (head + 1) % slots;
    XB_CHECK(cap_form_function(c, x, F));
    const double fprev = fnorm;
    fnorm = norm_of(c, F, &rc);
    XB_CHECK(rc);
    nl.iterations = k;
    nl.hist.push_back(fnorm);
    nl.reason = converged(fnorm);
    if (!nl.reason && fnorm > 1e3 * fprev) {  // the accelerated step went astray: restart the history
      mk = 0;
      pending = false;
    }
  }
  if (!nl.reason) nl.reason = -5;  // SNES_DIVERGED_MAX_IT
  nl.fnorm = fnorm;
  XB_CHECK(cap_read_counters(c));
  if (nl.profile) {
    XB_CUDA(cudaStreamSynchronize(c->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < nl.events_used; i += 2) {
      float t = 0.f;
      XB_CUDA(cudaEventElapsedTime(&t, nl.events[i], nl.events[i + 1]));
      ms += t;
    }
    nl.push_ms += ms;
    nl.push_evals += (int64_t)(nl.events_used / 2);
  }
  if (nl.reason < 0) XB_FAIL("SNESSolve has not converged: reason " + std::to_string(nl.reason) + ", |F| = " + std::to_string(fnorm));
  return 0;
}

// after_iteration (simulation.cpp:106-129): E, B update; the particles of the last evaluation become
// the stored state and are re-binned
int cap_finish(xb_ctx* c)
{
  XB_CHECK(final_update(c, c->cap_x));  // E = 2 x - E ; B -= dt curl^+ x
  for (auto& s : c->sorts) {
    if (s.count > 0 && c->track_ids)
      XB_CUDA(cudaMemcpyAsync(s.id[1 - s.cur], s.id[s.cur], sizeof(uint64_t) * s.count, cudaMemcpyDeviceToDevice, c->stream));
    s.cur = 1 - s.cur;
    s.sorted = false;
    if (c->g.nranks > 1) XB_CHECK(migrate_and_sort(c, s, 0.0));
    else XB_CHECK(particles_sort(c, s, 0.0));
  }
  return 0;
}

}  // namespace xb
