// xpic_oracle.cpp -- CPU restatement of xpic's ECSIM / ECSIMCorr step.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg may load it.  The product path
// (xpic_b200/csrc) never links or calls anything in oracle/.
//
// Parity status: pinned to the reference's golden files
//   tests/ecsim/expected/ecsim_ex1/temporal/{energy,energy_conservation}.txt
//   tests/ecsimcorr/expected/ecsimcorr_ex1/temporal/{energy,energy_conservation}.txt
//   tests/ecsim/expected/ecsim_ex1/{E,B}/{050,100}
// (copies under tests/golden/, see tests/golden/README.md) with curl_sign = -1, see
// DESIGN.md "sign arbitration".  The Krylov iteration path of PETSc (un-vendored,
// un-pinned third party: GMRES(30)+ILU(0) defaults) is NOT restated -- "parity unpinned"
// for residual histories; the converged solution is what the goldens pin.
//
// Plain C++17, no dependencies.  Single-threaded by default (deterministic, what the tests use);
// xo_set_threads(n) enables OpenMP over cells / rows for the CPU-baseline timings, with the same
// loop structure the reference parallelises (omp parallel for over cells + atomics on shared arrays,
// src/impls/ecsim/particles.cpp:41-47,137-142).  Every routine cites the reference
// file:line (relative to /root/reference) whose arithmetic it follows.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <list>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double mec2 = 511.0;  // src/constants.h:30

struct Point {  // src/interfaces/point.h:7-35 (+ id, which carries no arithmetic)
  double r[3];
  double p[3];
  uint64_t id;
};

struct Solver {  // src/impls/ecsim/simulation.h:15-18
  double rtol = 1e-7, atol = 1e-7;
  int maxit = 100, restart = 30;
  int iterations = 0, reason = 0;
  double rnorm = 0.0;
};

struct Csr {
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
};

struct Species {
  double q, m, n;
  int Np;
  std::vector<std::list<Point>> storage;  // src/interfaces/particles.h:32
  std::vector<double> currI, currJe;      // per-sort global currents
  double energy = 0, pred_w = 0, corr_w = 0, pred_dK = 0, corr_dK = 0, lambda_dK = 0;
  uint64_t next_id = 0;
  // eccapfim (src/impls/eccapfim/particles.h:31-37)
  std::vector<std::vector<Point>> previous_storage;
  std::vector<double> J;
  int64_t size = 0;
  double avgit = 0, avgcell = 0;
};

// SNES NGMRES state: tolerances src/impls/eccapfim/simulation.h:14-19, the rest PETSc's
// SNESNGMRES defaults (third party, un-vendored: src/snes/impls/ngmres/{snesngmres,ngmresfunc}.c)
struct Ngmres {
  double atol = 1e-7, rtol = 1e-7, stol = 1e-7;
  int maxit = 1000;
  int msize = 30, restart_it = 2;
  double gammaA = 2.0, gammaC = 2.0, epsilonB = 0.1, deltaB = 0.9;
  // NOT in the reference: solve P F(x) = 0 with the fixed linear operator
  // P = ((1 + shift) I + 1/4 dt^2 curl curl)^-1 (tests use it to reach tight tolerances quickly)
  int precond = 0;
  double shift = 0.0;
  double cn_tol = 0.5 * 1e-7;  // per-particle Picard tolerance (particles.cpp:100-101); tests may tighten it
  int iterations = 0, fevals = 0, reason = 0;
  std::vector<double> hist;
};

struct Sim {
  int N[3];
  double d[3], L[3], dt;
  int curl_sign;
  int64_t nc, n3;
  std::vector<double> E, Ep, Ec, B, B0, currI, currJe;
  std::vector<Species> sorts;
  Csr matL;  // pattern: 123 columns per row (fewer when periodic aliasing merges them)
  Csr matM;  // 13 columns per row
  Solver predict, correct;
  std::mt19937 gen;  // src/utils/random_generator.h:20-27 (default seed 5489)
  std::uniform_real_distribution<double> uni{0.0, 1.0};
  double last_j_diff_norm = 0.0;
  std::vector<double> Ehk, J;  // eccapfim: E^{n+1/2,k} and the total current of the last evaluation
  Ngmres snes;

  // DM_BOUNDARY_NONE / GHOSTED along z (src/utils/configuration.cpp:88-108): nodes outside the box do not exist.
  // Their matrix entries and deposits are dropped (Operator::remap_stencil gives -1, MatSetValuesCOO ignores it,
  // src/utils/operators.cpp:12-43), the ghost values the gathers read there are zero, and a particle that leaves
  // through such a face is removed (src/interfaces/particles.cpp:100-103).
  bool open_z = false;
  inline int wrap(int i, int a) const
  {
    int n = N[a];
    i %= n;
    return i < 0 ? i + n : i;
  }
  // src/utils/utils.h:20-24  ((z*Ny + y)*Nx + x)*3 + c, periodic ghosts folded; -1: no such node (open boundary)
  inline int64_t vidx(int x, int y, int z, int c) const
  {
    if (open_z && (z < 0 || z >= N[2])) return -1;
    return (((int64_t)wrap(z, 2) * N[1] + wrap(y, 1)) * N[0] + wrap(x, 0)) * 3 + c;
  }
  inline double get(const double* f, int x, int y, int z, int c) const
  {
    const int64_t i = vidx(x, y, z, c);
    return i < 0 ? 0.0 : f[i];
  }
  inline int64_t cell(int x, int y, int z) const { return ((int64_t)z * N[1] + y) * N[0] + x; }
};

// ---------------------------------------------------------------------------------------------
// CSR helpers
// ---------------------------------------------------------------------------------------------
void csr_from_pattern(Csr& A, int64_t nrows, const std::vector<std::vector<int32_t>>& cols)
{
  A.rowptr.assign(nrows + 1, 0);
  for (int64_t r = 0; r < nrows; ++r) A.rowptr[r + 1] = A.rowptr[r] + (int64_t)cols[r].size();
  A.col.resize(A.rowptr[nrows]);
  A.val.assign(A.rowptr[nrows], 0.0);
  for (int64_t r = 0; r < nrows; ++r) std::copy(cols[r].begin(), cols[r].end(), A.col.begin() + A.rowptr[r]);
}

inline double& csr_at(Csr& A, int64_t row, int32_t col)
{
  auto b = A.col.begin() + A.rowptr[row], e = A.col.begin() + A.rowptr[row + 1];
  auto it = std::lower_bound(b, e, col);
  return A.val[it - A.col.begin()];
}

int g_threads = 1;

void csr_mult(const Csr& A, const double* x, double* y, int64_t nrows)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int64_t r = 0; r < nrows; ++r) {
    double s = 0.0;
    for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) s += A.val[k] * x[A.col[k]];
    y[r] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Yee curls: src/utils/operators.cpp:155-215 (values :158-160, stencils :175-213)
//   positive (forward differences, E sites -> B sites), negative (backward, B -> E sites)
// ---------------------------------------------------------------------------------------------
void curl(const Sim& s, bool positive, const double* f, double* out)
{
  const double sg = (double)s.curl_sign;
  const double ix = 1.0 / s.d[0], iy = 1.0 / s.d[1], iz = 1.0 / s.d[2];
  for (int z = 0; z < s.N[2]; ++z)
    for (int y = 0; y < s.N[1]; ++y)
      for (int x = 0; x < s.N[0]; ++x) {
        double cx, cy, cz;
        if (positive) {
          const int xp = x + 1, yp = y + 1, zp = z + 1;
          cx = (+iy * s.get(f, x, yp, z, 2) - iy * s.get(f, x, y, z, 2)) + (-iz * s.get(f, x, y, zp, 1) + iz * s.get(f, x, y, z, 1));
          cy = (-ix * s.get(f, xp, y, z, 2) + ix * s.get(f, x, y, z, 2)) + (+iz * s.get(f, x, y, zp, 0) - iz * s.get(f, x, y, z, 0));
          cz = (+ix * s.get(f, xp, y, z, 1) - ix * s.get(f, x, y, z, 1)) + (-iy * s.get(f, x, yp, z, 0) + iy * s.get(f, x, y, z, 0));
        }
        else {
          const int xm = x - 1, ym = y - 1, zm = z - 1;
          cx = (+iy * s.get(f, x, y, z, 2) - iy * s.get(f, x, ym, z, 2)) + (-iz * s.get(f, x, y, z, 1) + iz * s.get(f, x, y, zm, 1));
          cy = (-ix * s.get(f, x, y, z, 2) + ix * s.get(f, xm, y, z, 2)) + (+iz * s.get(f, x, y, z, 0) - iz * s.get(f, x, y, zm, 0));
          cz = (+ix * s.get(f, x, y, z, 1) - ix * s.get(f, xm, y, z, 1)) + (-iy * s.get(f, x, y, z, 0) + iy * s.get(f, x, ym, z, 0));
        }
        const int64_t o = s.vidx(x, y, z, 0);
        out[o + 0] = sg * cx;
        out[o + 1] = sg * cy;
        out[o + 2] = sg * cz;
      }
}

// matM = 0.5 dt^2 * curl^- curl^+ + 2 I   (src/impls/ecsim/simulation.cpp:544-552)
// built column by column by applying the two curls to unit vectors of a local neighbourhood;
// curl_sign cancels in the product.
void build_matM(Sim& s)
{
  const int64_t n3 = s.n3;
  std::vector<std::vector<std::pair<int32_t, double>>> rows(n3);
  // analytic stencil of curl^- curl^+: apply both operators symbolically per row.
  // row (x,y,z,c):  sum_k cm[k] * ( sum_l cp[l] * E[...] )
  struct Term {
    int dx, dy, dz, c;
    double v;
  };
  auto curl_terms = [&](bool positive, int c, std::vector<Term>& t) {
    const double ix = 1.0 / s.d[0], iy = 1.0 / s.d[1], iz = 1.0 / s.d[2];
    t.clear();
    const int o = positive ? +1 : -1;
    // positive: (f[+1] - f[0]) / h ; negative: (f[0] - f[-1]) / h
    auto diff = [&](int axis, int comp, double sign, double ih) {
      int a[3] = {0, 0, 0};
      a[axis] = o;
      if (positive) {
        t.push_back({a[0], a[1], a[2], comp, +sign * ih});
        t.push_back({0, 0, 0, comp, -sign * ih});
      }
      else {
        t.push_back({0, 0, 0, comp, +sign * ih});
        t.push_back({a[0], a[1], a[2], comp, -sign * ih});
      }
    };
    if (c == 0) { diff(1, 2, +1, iy); diff(2, 1, -1, iz); }
    if (c == 1) { diff(2, 0, +1, iz); diff(0, 2, -1, ix); }
    if (c == 2) { diff(0, 1, +1, ix); diff(1, 0, -1, iy); }
  };
  std::vector<Term> tm, tp;
  for (int z = 0; z < s.N[2]; ++z)
    for (int y = 0; y < s.N[1]; ++y)
      for (int x = 0; x < s.N[0]; ++x)
        for (int c = 0; c < 3; ++c) {
          const int64_t row = s.vidx(x, y, z, c);
          auto& r = rows[row];
          curl_terms(false, c, tm);
          for (auto& a : tm) {
            curl_terms(true, a.c, tp);
            for (auto& b : tp) {
              // both curls drop their own missing columns: the intermediate B site must exist as well
              if (s.vidx(x + a.dx, y + a.dy, z + a.dz, a.c) < 0) continue;
              const int32_t col = (int32_t)s.vidx(x + a.dx + b.dx, y + a.dy + b.dy, z + a.dz + b.dz, b.c);
              if (col < 0) continue;
              r.push_back({col, 0.5 * s.dt * s.dt * a.v * b.v});
            }
          }
          r.push_back({(int32_t)row, 2.0});
        }
  std::vector<std::vector<int32_t>> cols(n3);
  for (int64_t r = 0; r < n3; ++r) {
    for (auto& e : rows[r]) cols[r].push_back(e.first);
    std::sort(cols[r].begin(), cols[r].end());
    cols[r].erase(std::unique(cols[r].begin(), cols[r].end()), cols[r].end());
  }
  csr_from_pattern(s.matM, n3, cols);
  for (int64_t r = 0; r < n3; ++r)
    for (auto& e : rows[r]) csr_at(s.matM, r, e.first) += e.second;
}

// Sparsity of matL: src/impls/ecsim/simulation.cpp:370-469 (every cell assembled; PETSc keeps
// explicit zeros so the pattern is the full 123 per row).
void window(int c, int (&lo)[3], int (&sz)[3])
{
  for (int a = 0; a < 3; ++a) {
    lo[a] = (a == c) ? -1 : 0;  // :409-411
    sz[a] = (a == c) ? 3 : 2;   // :405-407
  }
}

void build_matL_pattern(Sim& s)
{
  std::vector<std::vector<int32_t>> cols(s.n3);
  int lo1[3], sz1[3], lo2[3], sz2[3];
  for (int z = 0; z < s.N[2]; ++z)
    for (int y = 0; y < s.N[1]; ++y)
      for (int x = 0; x < s.N[0]; ++x)
        for (int c1 = 0; c1 < 3; ++c1) {
          window(c1, lo1, sz1);
          for (int k1 = 0; k1 < sz1[2]; ++k1)
            for (int j1 = 0; j1 < sz1[1]; ++j1)
              for (int i1 = 0; i1 < sz1[0]; ++i1) {
                const int64_t row = s.vidx(x + i1 + lo1[0], y + j1 + lo1[1], z + k1 + lo1[2], c1);
                if (row < 0) continue;
                for (int c2 = 0; c2 < 3; ++c2) {
                  window(c2, lo2, sz2);
                  for (int k2 = 0; k2 < sz2[2]; ++k2)
                    for (int j2 = 0; j2 < sz2[1]; ++j2)
                      for (int i2 = 0; i2 < sz2[0]; ++i2) {
                        const int64_t col = s.vidx(x + i2 + lo2[0], y + j2 + lo2[1], z + k2 + lo2[2], c2);
                        if (col >= 0) cols[row].push_back((int32_t)col);
                      }
                }
              }
        }
  for (auto& c : cols) {
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
  }
  csr_from_pattern(s.matL, s.n3, cols);
}

// ---------------------------------------------------------------------------------------------
// Weights: src/impls/ecsim/particles.cpp:76-105 == src/impls/ecsim/simulation.cpp:17-43
// ---------------------------------------------------------------------------------------------
struct W {
  int in[3], is[3];
  double wn[3][2], ws[3][2];
};

inline void weights(const Sim& s, const double* r, W& w)
{
  for (int a = 0; a < 3; ++a) {
    const double xn = r[a] / s.d[a];
    const double xs = xn - 0.5;
    w.in[a] = (int)std::floor(xn);
    w.is[a] = (int)std::floor(xs);
    w.wn[a][1] = xn - w.in[a];
    w.wn[a][0] = 1 - w.wn[a][1];
    w.ws[a][1] = xs - w.is[a];
    w.ws[a][0] = 1 - w.ws[a][1];
  }
}

// src/impls/ecsim/simulation.cpp:8-62
void interpolate_E_s1(const Sim& s, const std::vector<double>& Eg, const W& w, double* Ep)
{
  Ep[0] = Ep[1] = Ep[2] = 0.0;
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        const double sx = w.wn[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sy = w.wn[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sz = w.ws[2][k] * w.wn[1][j] * w.wn[0][i];
        Ep[0] += Eg[s.vidx(w.is[0] + i, w.in[1] + j, w.in[2] + k, 0)] * sx;
        Ep[1] += Eg[s.vidx(w.in[0] + i, w.is[1] + j, w.in[2] + k, 1)] * sy;
        Ep[2] += Eg[s.vidx(w.in[0] + i, w.in[1] + j, w.is[2] + k, 2)] * sz;
      }
}

// src/impls/ecsim/simulation.cpp:64-118
void interpolate_B_s1(const Sim& s, const std::vector<double>& Bg, const W& w, double* Bp)
{
  Bp[0] = Bp[1] = Bp[2] = 0.0;
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        const double sx = w.ws[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sy = w.ws[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sz = w.wn[2][k] * w.ws[1][j] * w.ws[0][i];
        Bp[0] += Bg[s.vidx(w.in[0] + i, w.is[1] + j, w.is[2] + k, 0)] * sx;
        Bp[1] += Bg[s.vidx(w.is[0] + i, w.in[1] + j, w.is[2] + k, 1)] * sy;
        Bp[2] += Bg[s.vidx(w.is[0] + i, w.is[1] + j, w.in[2] + k, 2)] * sz;
      }
}

inline void cross(const double* a, const double* b, double* o)  // src/utils/vector3.h:212-219
{
  o[0] = +(a[1] * b[2] - a[2] * b[1]);
  o[1] = -(a[0] * b[2] - a[2] * b[0]);
  o[2] = +(a[0] * b[1] - a[1] * b[0]);
}
inline double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// src/algorithms/boris_push.cpp:48-57
void update_vEB(double dt, double qm, const double* Ep, const double* Bp, double* v)
{
  const double alpha = dt * qm;
  double a[3], b[3], w[3], bw[3], bbw[3];
  for (int c = 0; c < 3; ++c) {
    a[c] = +alpha * Ep[c];
    b[c] = -alpha * Bp[c];
    w[c] = v[c] + 0.5 * a[c];
  }
  cross(b, w, bw);
  cross(b, bw, bbw);
  const double den = 1.0 + 0.25 * dot(b, b);
  for (int c = 0; c < 3; ++c) v[c] += a[c] + (bw[c] + 0.5 * bbw[c]) / den;
}

// src/interfaces/point.cpp:18-26
inline void g_bound_periodic(const Sim& s, Point& p, int a)
{
  double& x = p.r[a];
  if (x < 0.0)
    x = s.L[a] - (0.0 - x);
  else if (x > s.L[a])
    x = 0.0 + (x - s.L[a]);
}

// src/interfaces/particles.cpp:79-116 (sequential re-binning, all axes periodic)
void update_cells_seq(const Sim& s, Species& sp)
{
  for (int64_t g = 0; g < s.nc; ++g) {
    auto& cellg = sp.storage[g];
    auto it = cellg.begin();
    while (it != cellg.end()) {
      for (int a = 0; a < 3; ++a)
        if (!(a == 2 && s.open_z)) g_bound_periodic(s, *it, a);  // Particles::correct_coordinates, interfaces/particles.cpp:329-338
      const int vx = (int)std::floor(it->r[0] / s.d[0]);
      const int vy = (int)std::floor(it->r[1] / s.d[1]);
      const int vz = (int)std::floor(it->r[2] / s.d[2]);
      const bool inside = vx >= 0 && vx < s.N[0] && vy >= 0 && vy < s.N[1] && vz >= 0 && vz < s.N[2];
      const int64_t ng = s.cell(vx, vy, vz);
      if (inside && ng == g) {
        ++it;
        continue;
      }
      if (inside) sp.storage[ng].emplace_back(*it);
      it = cellg.erase(it);
    }
  }
}

// src/interfaces/particles.cpp:47-57
bool add_particle(const Sim& s, Species& sp, const Point& pt)
{
  const int vx = (int)std::floor(pt.r[0] / s.d[0]);
  const int vy = (int)std::floor(pt.r[1] / s.d[1]);
  const int vz = (int)std::floor(pt.r[2] / s.d[2]);
  if (!(vx >= 0 && vx < s.N[0] && vy >= 0 && vy < s.N[1] && vz >= 0 && vz < s.N[2])) return false;
  sp.storage[s.cell(vx, vy, vz)].emplace_back(pt);
  return true;
}

// ---------------------------------------------------------------------------------------------
// Moments: src/impls/ecsim/particles.cpp:62-173.  The 1296-entry cell block (coo_v) is filled
// exactly as there (index formula :145-163) and then summed into CSR at the rows / columns
// src/impls/ecsim/simulation.cpp:417-466 generates for it.
// ---------------------------------------------------------------------------------------------
void decompose_ecsim_current(const Sim& s, const Species& sp, const Point& pt, std::vector<double>& currI, double* coo_v)
{
  const double q = sp.q, m = sp.m, mpw = sp.n / (double)sp.Np;
  W w;
  weights(s, pt.r, w);
  const int ox = w.is[0] - w.in[0] + 1, oy = w.is[1] - w.in[1] + 1, oz = w.is[2] - w.in[2] + 1;

  double Bp[3], b[3];
  interpolate_B_s1(s, s.B, w, Bp);
  for (int c = 0; c < 3; ++c) b[c] = Bp[c] * ((0.5 * s.dt) * q / m);
  const double* v = pt.p;
  double vxb[3];
  cross(v, b, vxb);
  const double vb = dot(v, b), b2 = dot(b, b);
  double I_p[3];
  for (int c = 0; c < 3; ++c) I_p[c] = q * mpw / (1. + b2) * (v[c] + vxb[c] + vb * b[c]);
  const double A_p = 0.5 * s.dt * s.dt * mpw * q * q / m / (1 + b2);
  const double matB[3][3] = {
    {1.0 + b[0] * b[0], +b[2] + b[0] * b[1], -b[1] + b[0] * b[2]},
    {-b[2] + b[1] * b[0], 1.0 + b[1] * b[1], +b[0] + b[1] * b[2]},
    {+b[1] + b[2] * b[0], -b[0] + b[2] * b[1], 1.0 + b[2] * b[2]},
  };

  int i[3], j[3];
  double s1[3], s2[3];
  for (int k1 = 0; k1 < 2; ++k1)
    for (int j1 = 0; j1 < 2; ++j1)
      for (int i1 = 0; i1 < 2; ++i1) {
        s1[0] = w.wn[2][k1] * w.wn[1][j1] * w.ws[0][i1];
        s1[1] = w.wn[2][k1] * w.ws[1][j1] * w.wn[0][i1];
        s1[2] = w.ws[2][k1] * w.wn[1][j1] * w.wn[0][i1];
        const int64_t node[3] = {s.vidx(w.is[0] + i1, w.in[1] + j1, w.in[2] + k1, 0), s.vidx(w.in[0] + i1, w.is[1] + j1, w.in[2] + k1, 1),
                                 s.vidx(w.in[0] + i1, w.in[1] + j1, w.is[2] + k1, 2)};
        for (int c = 0; c < 3; ++c) {
          if (node[c] < 0) continue;  // outside an open boundary: dropped
#pragma omp atomic update
          currI[node[c]] += s1[c] * I_p[c];
        }

        i[0] = (k1 * 2 + j1) * 3 + (ox + i1);
        i[1] = (k1 * 3 + (oy + j1)) * 2 + i1;
        i[2] = ((oz + k1) * 2 + j1) * 2 + i1;
        for (int k2 = 0; k2 < 2; ++k2)
          for (int j2 = 0; j2 < 2; ++j2)
            for (int i2 = 0; i2 < 2; ++i2) {
              s2[0] = w.ws[0][i2] * w.wn[1][j2] * w.wn[2][k2];
              s2[1] = w.wn[0][i2] * w.ws[1][j2] * w.wn[2][k2];
              s2[2] = w.wn[0][i2] * w.wn[1][j2] * w.ws[2][k2];
              j[0] = (k2 * 2 + j2) * 3 + (ox + i2);
              j[1] = (k2 * 3 + (oy + j2)) * 2 + i2;
              j[2] = ((oz + k2) * 2 + j2) * 2 + i2;
              for (int c1 = 0; c1 < 3; ++c1)
                for (int c2 = 0; c2 < 3; ++c2) {
                  const int ind = (c1 * 3 + c2) * 144 + (i[c1] * 12 + j[c2]);
                  coo_v[ind] += s1[c1] * s2[c2] * A_p * matB[c1][c2];
                }
            }
      }
}

void add_block_to_csr(Sim& s, int x, int y, int z, const double* coo_v)
{
  int lo1[3], sz1[3], lo2[3], sz2[3];
  for (int c1 = 0; c1 < 3; ++c1) {
    window(c1, lo1, sz1);
    for (int k1 = 0; k1 < sz1[2]; ++k1)
      for (int j1 = 0; j1 < sz1[1]; ++j1)
        for (int i1 = 0; i1 < sz1[0]; ++i1) {
          const int64_t row = s.vidx(x + i1 + lo1[0], y + j1 + lo1[1], z + k1 + lo1[2], c1);
          if (row < 0) continue;
          const int i = (k1 * sz1[1] + j1) * sz1[0] + i1;
          for (int c2 = 0; c2 < 3; ++c2) {
            window(c2, lo2, sz2);
            for (int k2 = 0; k2 < sz2[2]; ++k2)
              for (int j2 = 0; j2 < sz2[1]; ++j2)
                for (int i2 = 0; i2 < sz2[0]; ++i2) {
                  const int32_t col = (int32_t)s.vidx(x + i2 + lo2[0], y + j2 + lo2[1], z + k2 + lo2[2], c2);
                  if (col < 0) continue;
                  const int j = (k2 * sz2[1] + j2) * sz2[0] + i2;
                  const int ind = (c1 * 3 + c2) * 144 + (i * 12 + j);
                  double& dst = csr_at(s.matL, row, col);
#pragma omp atomic update
                  dst += coo_v[ind];
                }
          }
        }
  }
}

// src/impls/ecsim/simulation.cpp:336-368,471-484 + src/impls/ecsim/particles.cpp:33-59
void fill_ecsim_current(Sim& s)
{
  std::fill(s.matL.val.begin(), s.matL.val.end(), 0.0);
  for (auto& sp : s.sorts) {
    // per-sort currI was zeroed in clear_sources
#pragma omp parallel num_threads(g_threads)
    {
      std::vector<double> coo_v(1296);
#pragma omp for schedule(dynamic, 16)
      for (int64_t gcell = 0; gcell < s.nc; ++gcell) {
        const int x = (int)(gcell % s.N[0]), y = (int)((gcell / s.N[0]) % s.N[1]), z = (int)(gcell / ((int64_t)s.N[0] * s.N[1]));
        const auto& cell = sp.storage[gcell];
        if (cell.empty()) continue;
        std::fill(coo_v.begin(), coo_v.end(), 0.0);
        for (const auto& pt : cell) decompose_ecsim_current(s, sp, pt, sp.currI, coo_v.data());
        add_block_to_csr(s, x, y, z, coo_v.data());
      }
    }
    for (int64_t i = 0; i < s.n3; ++i) s.currI[i] += sp.currI[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Esirkepov: src/utils/shape.cpp:12-79, src/interfaces/sort_parameters.cpp:21-30,
//            src/algorithms/esirkepov_decomposition.cpp:20-103
// ---------------------------------------------------------------------------------------------
inline double spline2(double x)
{
  x = std::abs(x);
  if (x <= 0.5) return (0.75 - x * x);
  if (0.5 < x && x < 1.5) return 0.5 * (1.5 - x) * (1.5 - x);
  return 0.0;
}

void esirkepov(const Sim& s, const double* old_r, const double* new_r, double alpha, std::vector<double>& J)
{
  constexpr double radius = 1.5;
  constexpr int shw = 4;
  int start[3], size[3];
  double po[3], pn[3];
  for (int a = 0; a < 3; ++a) {
    po[a] = old_r[a] / s.d[a];
    pn[a] = new_r[a] / s.d[a];
    start[a] = (int)std::round(std::min(po[a], pn[a]) - radius);
    size[a] = (int)std::floor(std::max(po[a], pn[a]) + radius) + 1 - start[a];
  }
  double So[3][shw], Sn[3][shw];
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < size[a]; ++i) {
      const double g = (double)(start[a] + i);
      So[a][i] = spline2(po[a] - g);
      Sn[a][i] = spline2(pn[a] - g);
    }
  double tjx[shw * shw] = {0}, tjy[shw * shw] = {0}, tjz[shw * shw] = {0};
  const double qx = alpha * s.d[0], qy = alpha * s.d[1], qz = alpha * s.d[2];
  for (int z = 0; z < size[2]; ++z)
    for (int y = 0; y < size[1]; ++y)
      for (int x = 0; x < size[0]; ++x) {
        const double nX = Sn[0][x], oX = So[0][x], nY = Sn[1][y], oY = So[1][y], nZ = Sn[2][z], oZ = So[2][z];
        const double wx = -qx * (nX - oX) * (nY * (2.0 * nZ + oZ) + oY * (2.0 * oZ + nZ));
        const double wy = -qy * (nY - oY) * (nX * (2.0 * nZ + oZ) + oX * (2.0 * oZ + nZ));
        const double wz = -qz * (nZ - oZ) * (nY * (2.0 * nX + oX) + oY * (2.0 * oX + nX));
        double& jx = tjx[z * shw + y];
        double& jy = tjy[z * shw + x];
        double& jz = tjz[y * shw + x];
        jx = ((double)(x > 0) * jx) + wx;
        jy = ((double)(y > 0) * jy) + wy;
        jz = ((double)(z > 0) * jz) + wz;
        const int64_t o = s.vidx(start[0] + x, start[1] + y, start[2] + z, 0);
        J[o + 0] += jx;
        J[o + 1] += jy;
        J[o + 2] += jz;
      }
}

// ---------------------------------------------------------------------------------------------
// Restarted GMRES (stand-in for PETSc's KSP; see header note).  Zero initial guess, modified
// Gram-Schmidt, true-residual stop test  ||r|| <= max(rtol*||b||, atol).
// ---------------------------------------------------------------------------------------------
template <class Op>
void gmres(Op&& apply, int64_t n, const double* b, double* x, Solver& sv)
{
  const int m = sv.restart;
  std::vector<std::vector<double>> V(m + 1, std::vector<double>(n));
  std::vector<double> H((m + 1) * m), cs(m), sn(m), g(m + 1), w(n), r(n), y(m);
  auto nrm = [&](const double* a) {
    double s = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * a[i];
    return std::sqrt(s);
  };
  std::fill(x, x + n, 0.0);
  const double bnorm = nrm(b);
  const double tol = std::max(sv.rtol * bnorm, sv.atol);
  sv.iterations = 0;
  sv.reason = 0;
  std::copy(b, b + n, r.begin());
  double rnorm = bnorm;
  while (true) {
    if (rnorm <= tol) { sv.reason = (rnorm <= sv.atol) ? 3 : 2; break; }
    if (sv.iterations >= sv.maxit) { sv.reason = -3; break; }
    for (int64_t i = 0; i < n; ++i) V[0][i] = r[i] / rnorm;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = rnorm;
    int k = 0;
    for (; k < m && sv.iterations < sv.maxit; ++k) {
      apply(V[k].data(), w.data());
      for (int i = 0; i <= k; ++i) {
        double h = 0;
        const double* vi = V[i].data();
        double* wp = w.data();
#pragma omp parallel for num_threads(g_threads) reduction(+ : h) schedule(static)
        for (int64_t l = 0; l < n; ++l) h += wp[l] * vi[l];
        H[i * m + k] = h;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int64_t l = 0; l < n; ++l) wp[l] -= h * vi[l];
      }
      const double hn = nrm(w.data());
      H[(k + 1) * m + k] = hn;
      for (int64_t l = 0; l < n; ++l) V[k + 1][l] = w[l] / hn;
      for (int i = 0; i < k; ++i) {
        const double t = cs[i] * H[i * m + k] + sn[i] * H[(i + 1) * m + k];
        H[(i + 1) * m + k] = -sn[i] * H[i * m + k] + cs[i] * H[(i + 1) * m + k];
        H[i * m + k] = t;
      }
      const double a = H[k * m + k], bb = H[(k + 1) * m + k], rr = std::hypot(a, bb);
      cs[k] = a / rr;
      sn[k] = bb / rr;
      H[k * m + k] = rr;
      H[(k + 1) * m + k] = 0;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      ++sv.iterations;
      rnorm = std::abs(g[k + 1]);
      if (rnorm <= tol) { ++k; break; }
    }
    for (int i = k - 1; i >= 0; --i) {
      double t = g[i];
      for (int j = i + 1; j < k; ++j) t -= H[i * m + j] * y[j];
      y[i] = t / H[i * m + i];
    }
    for (int i = 0; i < k; ++i)
      for (int64_t l = 0; l < n; ++l) x[l] += y[i] * V[i][l];
    apply(x, w.data());
    for (int64_t l = 0; l < n; ++l) r[l] = b[l] - w[l];
    rnorm = nrm(r.data());
  }
  sv.rnorm = rnorm;
}

// src/impls/ecsim/simulation.cpp:255-278
void advance_fields(Sim& s, Solver& sv, bool with_L, const std::vector<double>& curr, std::vector<double>& out)
{
  std::vector<double> rhs(s.n3), Bm(s.n3), cb(s.n3), t1(s.n3);
  for (int64_t i = 0; i < s.n3; ++i) Bm[i] = s.B[i] - s.B0[i];
  curl(s, false, Bm.data(), cb.data());
  for (int64_t i = 0; i < s.n3; ++i) rhs[i] = (2 * s.E[i] + (-s.dt) * curr[i]) + s.dt * cb[i];
  auto apply = [&](const double* x, double* y) {
    csr_mult(s.matM, x, y, s.n3);
    if (with_L) {
      csr_mult(s.matL, x, t1.data(), s.n3);
      for (int64_t i = 0; i < s.n3; ++i) y[i] += t1[i];
    }
  };
  gmres(apply, s.n3, rhs.data(), out.data(), sv);
}

double kinetic_energy(const Species& sp)  // src/impls/ecsimcorr/particles.cpp:134-150
{
  const double mpw = sp.n / sp.Np;
  double e = 0.0;
  for (auto& cell : sp.storage)
    for (auto& pt : cell) e += 0.5 * (sp.m * dot(pt.p, pt.p)) * mpw;  // diagnostics/energy.cpp:188-191
  return e;
}

// ---------------------------------------------------------------------------------------------
// The steps
// ---------------------------------------------------------------------------------------------
void clear_sources(Sim& s)  // ecsim/simulation.cpp:157-172, ecsim/particles.cpp:194-200
{
  std::fill(s.currI.begin(), s.currI.end(), 0.0);
  std::fill(s.currJe.begin(), s.currJe.end(), 0.0);
  for (auto& sp : s.sorts) {
    std::fill(sp.currI.begin(), sp.currI.end(), 0.0);
    std::fill(sp.currJe.begin(), sp.currJe.end(), 0.0);
  }
}

void final_update_fields(Sim& s)  // ecsim/simulation.cpp:241-253
{
  std::vector<double> ce(s.n3);
  curl(s, true, s.Ep.data(), ce.data());
  for (int64_t i = 0; i < s.n3; ++i) {
    s.E[i] = 2 * s.Ep[i] + (-1) * s.E[i];
    s.B[i] = s.B[i] + (-s.dt) * ce[i];
  }
}

void step_ecsim(Sim& s)  // ecsim/simulation.cpp:145-155
{
  clear_sources(s);
  for (auto& sp : s.sorts) {  // first_push :174-189, particles.cpp:21-31
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16)
    for (int64_t gc = 0; gc < s.nc; ++gc)
      for (auto& pt : sp.storage[gc])
        for (int c = 0; c < 3; ++c) pt.r[c] += pt.p[c] * s.dt;
    update_cells_seq(s, sp);  // serial in the reference as well (interfaces/particles.cpp:79-116)
  }
  fill_ecsim_current(s);
  advance_fields(s, s.predict, true, s.currI, s.Ep);
  for (auto& sp : s.sorts) {  // second_push :212-239, particles.cpp:175-192
    const double qm = sp.q / sp.m;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16)
    for (int64_t gc = 0; gc < s.nc; ++gc)
      for (auto& pt : sp.storage[gc]) {
        W w;
        weights(s, pt.r, w);
        double Ep[3], Bp[3];
        interpolate_E_s1(s, s.Ep, w, Ep);
        interpolate_B_s1(s, s.B, w, Bp);
        update_vEB(s.dt, qm, Ep, Bp, pt.p);
      }
    update_cells_seq(s, sp);
  }
  final_update_fields(s);
}

void step_ecsimcorr(Sim& s)  // ecsimcorr/simulation.cpp:21-32
{
  clear_sources(s);
  for (auto& sp : s.sorts) sp.energy = kinetic_energy(sp);  // :44-45
  for (auto& sp : s.sorts) {                                // ecsimcorr/particles.cpp:27-50
    const double alpha = sp.q * sp.n / sp.Np / (6.0 * s.dt);
    for (auto& cell : sp.storage)
      for (auto& pt : cell) {
        double old_r[3] = {pt.r[0], pt.r[1], pt.r[2]};
        for (int c = 0; c < 3; ++c) pt.r[c] += pt.p[c] * (0.5 * s.dt);
        esirkepov(s, old_r, pt.r, alpha, sp.currJe);
      }
    update_cells_seq(s, sp);
  }
  fill_ecsim_current(s);
  advance_fields(s, s.predict, true, s.currI, s.Ep);
  for (auto& sp : s.sorts) {  // ecsimcorr/particles.cpp:52-91
    const double qm = sp.q / sp.m, qn_Np = sp.q * sp.n / sp.Np;
    const double alpha = qn_Np / (6.0 * s.dt);
    sp.pred_w = 0.0;
    for (auto& cell : sp.storage)
      for (auto& pt : cell) {
        double old_r[3] = {pt.r[0], pt.r[1], pt.r[2]};
        double old_v[3] = {pt.p[0], pt.p[1], pt.p[2]};
        W w;
        weights(s, pt.r, w);
        double Ep[3], Bp[3];
        interpolate_E_s1(s, s.Ep, w, Ep);
        interpolate_B_s1(s, s.B, w, Bp);
        update_vEB(s.dt, qm, Ep, Bp, pt.p);
        for (int c = 0; c < 3; ++c) pt.r[c] += pt.p[c] * (0.5 * s.dt);
        esirkepov(s, old_r, pt.r, alpha, sp.currJe);
        double vs[3] = {old_v[0] + pt.p[0], old_v[1] + pt.p[1], old_v[2] + pt.p[2]};
        sp.pred_w += qn_Np * 0.5 * dot(vs, Ep);
      }
    for (int64_t i = 0; i < s.n3; ++i) s.currJe[i] += sp.currJe[i];
    update_cells_seq(s, sp);
  }
  advance_fields(s, s.correct, false, s.currJe, s.Ec);  // correct_fields :52-63
  for (auto& sp : s.sorts) {                             // ecsimcorr/particles.cpp:93-126
    double cw = 0.0;
    for (int64_t i = 0; i < s.n3; ++i) cw += sp.currJe[i] * s.Ec[i];
    sp.corr_w = cw;
    const double K0 = sp.energy;
    const double K = kinetic_energy(sp);
    const double lambda2 = 1.0 + s.dt * (sp.corr_w - sp.pred_w) / K;
    const double lambda = std::sqrt(lambda2);
    for (auto& cell : sp.storage)
      for (auto& pt : cell)
        for (int c = 0; c < 3; ++c) pt.p[c] *= lambda;
    sp.lambda_dK = (lambda2 - 1.0) * K;
    sp.pred_dK = K - K0;
    sp.corr_dK = lambda2 * K - K0;
    sp.energy = lambda2 * K;
  }
  {  // ecsimcorr/simulation.cpp:74-82 (logged only)
    std::vector<double> t(s.n3);
    csr_mult(s.matL, s.Ec.data(), t.data(), s.n3);
    double nn = 0;
    for (int64_t i = 0; i < s.n3; ++i) {
      s.currI[i] += t[i];
      const double u = -s.currI[i] + s.currJe[i];
      nn += u * u;
    }
    s.last_j_diff_norm = std::sqrt(nn);
  }
  std::swap(s.Ep, s.Ec);  // :84
  final_update_fields(s);
}

// =============================================================================================
// eccapfim: fully implicit, energy- and charge-conserving scheme (BASELINE config 5)
//   src/impls/eccapfim/simulation.cpp:36-241, particles.cpp:30-181, cell_traversal.cpp:3-77,
//   src/algorithms/implicit_esirkepov.cpp:11-117, src/utils/shape.cpp:3-107
// Golden: tests/eccapfim/expected/eccapfim_ex1 (copy under tests/golden/eccapfim_ex1); its
// convergence_history.txt records PETSc's NGMRES residual history, so the nonlinear solver below
// restates SNESSolve_NGMRES (PETSc is un-vendored and un-pinned; defaults of release 3.22) and is
// checked against that history.
// ---------------------------------------------------------------------------------------------
using V3 = std::array<double, 3>;

// src/impls/eccapfim/cell_traversal.cpp:3-77 (Amanatides & Woo on the half-shifted lattice)
void cell_traversal(const Sim& s, const V3& end, const V3& start, std::vector<V3>& points)
{
  points.clear();
  int curr[3], last[3];
  for (int a = 0; a < 3; ++a) {
    curr[a] = (int)std::round(start[a] / s.d[a]);
    last[a] = (int)std::round(end[a] / s.d[a]);
  }
  if (curr[0] == last[0] && curr[1] == last[1] && curr[2] == last[2]) {
    points.push_back(start);
    points.push_back(end);
    return;
  }
  static const double maxv = std::numeric_limits<double>::max();
  double dir[3], next[3], t3[3], dt3[3];
  int sg[3];
  for (int a = 0; a < 3; ++a) {
    dir[a] = end[a] - start[a];
    sg[a] = dir[a] > 0 ? 1 : -1;
    next[a] = (curr[a] + sg[a] * 0.5) * s.d[a];
    t3[a] = (dir[a] != 0) ? (next[a] - start[a]) / dir[a] : maxv;
    dt3[a] = (dir[a] != 0) ? s.d[a] / dir[a] * sg[a] : 0.0;
  }
  points.push_back(start);
  int guard = 0;
  while (!(curr[0] == last[0] && curr[1] == last[1] && curr[2] == last[2]) && ++guard < 256) {
    int a;
    if (t3[0] < t3[1])
      a = (t3[0] < t3[2]) ? 0 : 2;
    else
      a = (t3[1] < t3[2]) ? 1 : 2;
    const double t = t3[a];
    curr[a] += sg[a];
    t3[a] += dt3[a];
    points.push_back({start[0] + dir[0] * t, start[1] + dir[1] * t, start[2] + dir[2] * t});
  }
  points.push_back(end);
}

// ImplicitEsirkepov::Shape (src/algorithms/implicit_esirkepov.h:18-53, .cpp:11-60)
struct CapShape {
  int start[3];
  double cache[54];
};

inline double sfunc_1(double x) { return 1.0 - std::abs(x); }
inline double sfunc_21(double x) { x = std::abs(x); return (0.75 - x * x); }
inline double sfunc_22(double x) { x = std::abs(x); return 0.5 * ((1.5 - x) * (1.5 - x)); }
inline double sfunc_2(int j, double x) { return j == 1 ? sfunc_21(x) : sfunc_22(x); }

void cap_shape_setup(const Sim& s, const V3& rn, const V3& r0, CapShape& sh)
{
  double prn[3], pr0[3], prh[3], gc[3], gv[3];
  for (int a = 0; a < 3; ++a) {
    prn[a] = rn[a] / s.d[a];
    pr0[a] = r0[a] / s.d[a];
    prh[a] = 0.5 * (prn[a] + pr0[a]);
    gc[a] = std::round(prh[a]);
    sh.start[a] = (int)gc[a] - 1;
    gv[a] = gc[a] + 0.5;
  }
  static constexpr double sixth = 1.0 / 6.0;
  int m = 0;
  for (int cx = 0; cx < 3; ++cx) {
    const int cy = (cx + 1) % 3, cz = (cx + 2) % 3;
    for (int i = 0; i < 2; ++i) {
      const double shx = sixth * sfunc_1(gv[cx] + (i - 1) - prh[cx]);
      for (int j = 0; j < 3; ++j) {
        const double sny = sfunc_2(j, gc[cy] + (j - 1) - prn[cy]);
        const double s0y = sfunc_2(j, gc[cy] + (j - 1) - pr0[cy]);
        for (int k = 0; k < 3; ++k) {
          const double snz = sfunc_2(k, gc[cz] + (k - 1) - prn[cz]);
          const double s0z = sfunc_2(k, gc[cz] + (k - 1) - pr0[cz]);
          sh.cache[m++] = shx * (sny * (2 * snz + s0z) + s0y * (2 * s0z + snz));
        }
      }
    }
  }
}

// Shape::setup(r) with the global 2nd-order form factor (src/utils/shape.cpp:34-45,81-107,
// src/constants.h:4) + SimpleInterpolation::process for B (simple_interpolation.cpp:8-38,
// Shape::magnetic shape.h:66-73)
void interpolate_B_s2(const Sim& s, const std::vector<double>& Bg, const V3& r, double* Bp)
{
  constexpr double radius = 1.5;
  int start[3], size[3];
  double pr[3];
  for (int a = 0; a < 3; ++a) {
    pr[a] = r[a] / s.d[a];
    start[a] = (int)std::round(pr[a] - radius);
    size[a] = (int)std::floor(pr[a] + radius) + 1 - start[a];
  }
  const int n = size[0] * size[1] * size[2];
  for (int i = 0; i < n; ++i) {
    const int ix = i % size[0], iy = (i / size[0]) % size[1], iz = (i / size[0]) / size[1];
    double gx = (double)(start[0] + ix), gy = (double)(start[1] + iy), gz = (double)(start[2] + iz);
    const double nox = spline2(pr[0] - gx), noy = spline2(pr[1] - gy), noz = spline2(pr[2] - gz);
    gx += 0.5; gy += 0.5; gz += 0.5;
    const double shx = spline2(pr[0] - gx), shy = spline2(pr[1] - gy), shz = spline2(pr[2] - gz);
    const int64_t o = s.vidx(start[0] + ix, start[1] + iy, start[2] + iz, 0);
    Bp[0] += Bg[o + 0] * (shz * shy * nox);
    Bp[1] += Bg[o + 1] * (shz * noy * shx);
    Bp[2] += Bg[o + 2] * (noz * shy * shx);
  }
}

// ImplicitEsirkepov::interpolate (implicit_esirkepov.cpp:63-91): adds into Ep, Bp
void cap_interpolate(const Sim& s, const std::vector<double>& Eg, const std::vector<double>& Bg, double* Ep, double* Bp, const V3& rn, const V3& r0)
{
  const V3 rh = {0.5 * (rn[0] + r0[0]), 0.5 * (rn[1] + r0[1]), 0.5 * (rn[2] + r0[2])};
  interpolate_B_s2(s, Bg, rh, Bp);
  CapShape sh;
  cap_shape_setup(s, rn, r0, sh);
  int i[3], m = 0;
  for (int cx = 0; cx < 3; ++cx) {
    const int cy = (cx + 1) % 3, cz = (cx + 2) % 3;
    for (i[cx] = 0; i[cx] < 2; i[cx]++)
      for (i[cy] = 0; i[cy] < 3; i[cy]++)
        for (i[cz] = 0; i[cz] < 3; i[cz]++)
          Ep[cx] += Eg[s.vidx(sh.start[0] + i[0], sh.start[1] + i[1], sh.start[2] + i[2], cx)] * sh.cache[m++];
  }
}

// ImplicitEsirkepov::decompose (implicit_esirkepov.cpp:93-117)
void cap_decompose(const Sim& s, std::vector<double>& Jg, double alpha, const double* v, const V3& rn, const V3& r0)
{
  CapShape sh;
  cap_shape_setup(s, rn, r0, sh);
  int i[3], m = 0;
  for (int cx = 0; cx < 3; ++cx) {
    const int cy = (cx + 1) % 3, cz = (cx + 2) % 3;
    for (i[cx] = 0; i[cx] < 2; i[cx]++)
      for (i[cy] = 0; i[cy] < 3; i[cy]++)
        for (i[cz] = 0; i[cz] < 3; i[cz]++) {
          double& dst = Jg[s.vidx(sh.start[0] + i[0], sh.start[1] + i[1], sh.start[2] + i[2], cx)];
          const double add = alpha * v[cx] * sh.cache[m++];
#pragma omp atomic update
          dst += add;
        }
  }
}

// The closed form of one Crank-Nicolson velocity solve for given mean fields: particles.cpp:137-144 ==
// CrankNicolsonPush::process, src/algorithms/crank_nicolson_push.cpp:53-62
//   a = alpha E, b = alpha B, w = v0 + a, vh = (w + w x b + b (w . b)) / (1 + b^2),  alpha = dtau q / (2 m)
inline void cn_mean_velocity(double alpha, const double* v0, const double* Ep, const double* Bp, double* vh)
{
  double a[3], b[3], w[3], wxb[3];
  for (int c = 0; c < 3; ++c) {
    a[c] = alpha * Ep[c];
    b[c] = alpha * Bp[c];
    w[c] = v0[c] + a[c];
  }
  cross(w, b, wxb);
  const double wb = dot(w, b), den = 1.0 + dot(b, b);
  for (int c = 0; c < 3; ++c) vh[c] = ((w[c] + wxb[c]) + b[c] * wb) / den;
}

inline double len3(const V3& a, const V3& b) { return std::hypot(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }  // vector3.h:160-164

// eccapfim::Particles::form_iteration for one particle (particles.cpp:73-176).  `curr` enters as
// the start-of-step state and leaves as (x^{n+1,k}, v^{n+1,k}); returns Picard iterations and
// traversed segments through it / ncell.
void cap_push_particle(const Sim& s, const Species& sp, const std::vector<double>& Eg, const std::vector<double>& Bg, std::vector<double>& Jg, Point& curr,
                       double& it_sum, double& cell_sum, std::vector<V3>& coords)
{
  const double q = sp.q, m = sp.m, mpw = sp.n / sp.Np, dt = s.dt;
  const double maxv = std::numeric_limits<double>::max();
  // :39-45 single box: start = 0, end = N
  const double lo[3] = {(0 - 0.5) * s.d[0], (0 - 0.5) * s.d[1], (0 - 0.5) * s.d[2]};
  const double hi[3] = {(s.N[0] + 0.5) * s.d[0], (s.N[1] + 0.5) * s.d[1], (s.N[2] + 0.5) * s.d[2]};
  auto process_bound = [&](double vh, double x, double xb, double xe) {  // :49-56
    if (vh > 0 && std::abs(xe - x) > 1e-7) return (xe - x) / vh;
    else if (vh < 0 && std::abs(xb - x) > 1e-7) return (xb - x) / vh;
    else return maxv;
  };
  Point tmp = curr;
  Point& pn = curr;
  Point& p0 = tmp;
  double tau = 0, dtau = 0;
  for (; tau < dt; tau += dtau) {
    double vh[3];
    for (int c = 0; c < 3; ++c) vh[c] = 0.5 * (pn.p[c] + p0.p[c]);
    const double dtx = process_bound(vh[0], p0.r[0], lo[0], hi[0]);
    const double dty = process_bound(vh[1], p0.r[1], lo[1], hi[1]);
    const double dtz = process_bound(vh[2], p0.r[2], lo[2], hi[2]);
    dtau = std::min({dt - tau, dtx, dty, dtz});
    const double a0 = q * mpw;
    const double alpha = 0.5 * dtau * (q / m);
    const int cn_maxit = 30;
    const double cn_atol = s.snes.cn_tol, cn_rtol = s.snes.cn_tol;  // 0.5 * eccapfim::atol (:100-101)
    int it = 0;
    double Ep[3], Bp[3];
    auto set_fields = [&] {  // :109-127
      Ep[0] = Ep[1] = Ep[2] = Bp[0] = Bp[1] = Bp[2] = 0.0;
      const V3 rn = {pn.r[0], pn.r[1], pn.r[2]}, r0 = {p0.r[0], p0.r[1], p0.r[2]};
      const double d = len3(rn, r0);
      cell_traversal(s, rn, r0, coords);
      for (size_t k = 1; k < coords.size(); ++k) {
        const double ds = len3(coords[k], coords[k - 1]);
        const double bs = (d > 0 ? ds / d : 1.0);
        double Es[3] = {0, 0, 0}, Bs[3] = {0, 0, 0};
        cap_interpolate(s, Eg, Bg, Es, Bs, coords[k], coords[k - 1]);
        for (int c = 0; c < 3; ++c) {
          Ep[c] += Es[c] * bs;
          Bp[c] += Bs[c] * bs;
        }
      }
    };
    auto get_residue = [&] {  // :129-131
      double vxb[3];
      cross(vh, Bp, vxb);
      const double f = dtau * q / m;
      return std::hypot((pn.p[0] - p0.p[0]) - f * (Ep[0] + vxb[0]), (pn.p[1] - p0.p[1]) - f * (Ep[1] + vxb[1]), (pn.p[2] - p0.p[2]) - f * (Ep[2] + vxb[2]));
    };
    set_fields();
    double rn_, r0_;
    rn_ = r0_ = get_residue();
    for (; rn_ > cn_atol + cn_rtol * r0_ && it < cn_maxit; it++) {  // :136-148
      cn_mean_velocity(alpha, p0.p, Ep, Bp, vh);
      for (int c = 0; c < 3; ++c) {
        pn.r[c] = p0.r[c] + dtau * vh[c];
        pn.p[c] = 2.0 * vh[c] - p0.p[c];
      }
      set_fields();
      rn_ = get_residue();
    }
    it_sum += (double)it;
    cell_sum += (double)(coords.size() - 1);
    {  // :153-163
      const V3 rn = {pn.r[0], pn.r[1], pn.r[2]}, r0 = {p0.r[0], p0.r[1], p0.r[2]};
      const double d = len3(rn, r0);
      cell_traversal(s, rn, r0, coords);
      for (size_t k = 1; k < coords.size(); ++k) {
        const double ds = len3(coords[k], coords[k - 1]);
        const double bs = (d > 0 ? ds / d : 1.0);
        cap_decompose(s, Jg, a0 * bs * (dtau / dt), vh, coords[k], coords[k - 1]);
      }
    }
    bool reset = false;  // :58-68,165-171
    for (int a = 0; a < 3; ++a) {
      double& x = pn.r[a];
      if (x < 0.0) { x = s.L[a] - (0.0 - x); reset = true; }
      else if (x > s.L[a]) { x = 0.0 + (x - s.L[a]); reset = true; }
    }
    if (reset) p0 = pn;
  }
}

// eccapfim::Simulation::form_iteration: clear_sources, form_current, form_function
// (simulation.cpp:132-241): vf = F(vx)
void cap_form_function(Sim& s, const double* vx, double* vf)
{
  std::copy(vx, vx + s.n3, s.Ehk.begin());
  std::fill(s.J.begin(), s.J.end(), 0.0);
  for (auto& sp : s.sorts) {
    std::fill(sp.J.begin(), sp.J.end(), 0.0);
    double it_sum = 0, cell_sum = 0;
#pragma omp parallel num_threads(g_threads) reduction(+ : it_sum, cell_sum)
    {
      std::vector<V3> coords;
#pragma omp for schedule(dynamic, 16)
      for (int64_t g = 0; g < s.nc; ++g) {
        int i = 0;
        for (auto& curr : sp.storage[g]) {
          curr = sp.previous_storage[g][i++];
          cap_push_particle(s, sp, s.Ehk, s.B, sp.J, curr, it_sum, cell_sum, coords);
        }
      }
    }
    sp.avgit = sp.size ? it_sum / sp.size : 0.0;
    sp.avgcell = sp.size ? cell_sum / sp.size : 0.0;
    for (int64_t i = 0; i < s.n3; ++i) s.J[i] += sp.J[i];
  }
  // form_function :228-236 with matM = +1/4 dt^2 C-C+, rotB = -1/2 dt C- (:349-351)
  std::vector<double> t1(s.n3), t2(s.n3), cb(s.n3);
  curl(s, true, s.Ehk.data(), t1.data());
  curl(s, false, t1.data(), t2.data());
  curl(s, false, s.B.data(), cb.data());
  const double dt = s.dt;
  for (int64_t i = 0; i < s.n3; ++i) {
    double f = s.Ehk[i];
    f += (0.25 * dt * dt) * t2[i];
    f += -1 * s.E[i];
    f += (0.5 * dt) * s.J[i];
    f += (-0.5 * dt) * cb[i];
    vf[i] = f;
  }
  ++s.snes.fevals;
}

// P r, P = ((1 + shift) I + 1/4 dt^2 C-C+)^-1 by conjugate gradients (oracle-only helper)
void cap_precondition(Sim& s, const double* r, double* z)
{
  const int64_t n = s.n3;
  const double sh = 1.0 + s.snes.shift, c = 0.25 * s.dt * s.dt;
  std::vector<double> res(r, r + n), p(r, r + n), t1(n), Ap(n);
  std::fill(z, z + n, 0.0);
  double rr = 0;
  for (int64_t i = 0; i < n; ++i) rr += res[i] * res[i];
  const double stop = rr * 1e-30;
  for (int it = 0; it < 2000 && rr > stop && rr > 0; ++it) {
    curl(s, true, p.data(), t1.data());
    curl(s, false, t1.data(), Ap.data());
    double pAp = 0;
    for (int64_t i = 0; i < n; ++i) {
      Ap[i] = sh * p[i] + c * Ap[i];
      pAp += p[i] * Ap[i];
    }
    const double al = rr / pAp;
    double rn = 0;
    for (int64_t i = 0; i < n; ++i) {
      z[i] += al * p[i];
      res[i] -= al * Ap[i];
      rn += res[i] * res[i];
    }
    const double be = rn / rr;
    rr = rn;
    for (int64_t i = 0; i < n; ++i) p[i] = res[i] + be * p[i];
  }
}

// minimum-norm least squares  min |H a - b|  through a one-sided Jacobi SVD, singular values below
// eps * s_max dropped (LAPACK gelss with rcond = -1, as SNESNGMRESFormCombinedSolution_Private calls it)
void lstsq_svd(int l, std::vector<double> H /* l x l, H[i*l+j] */, std::vector<double>& b)
{
  std::vector<double> V(l * l, 0.0);
  for (int i = 0; i < l; ++i) V[i * l + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < l - 1; ++p)
      for (int q = p + 1; q < l; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < l; ++i) {
          al += H[i * l + p] * H[i * l + p];
          be += H[i * l + q] * H[i * l + q];
          ga += H[i * l + p] * H[i * l + q];
        }
        if (ga == 0.0 || std::abs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
        off = std::max(off, std::abs(ga) / std::sqrt(al * be));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < l; ++i) {
          const double hp = H[i * l + p], hq = H[i * l + q];
          H[i * l + p] = cs * hp - sn * hq;
          H[i * l + q] = sn * hp + cs * hq;
          const double vp = V[i * l + p], vq = V[i * l + q];
          V[i * l + p] = cs * vp - sn * vq;
          V[i * l + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  std::vector<double> sig(l), y(l, 0.0);
  double smax = 0;
  for (int j = 0; j < l; ++j) {
    double nn = 0;
    for (int i = 0; i < l; ++i) nn += H[i * l + j] * H[i * l + j];
    sig[j] = std::sqrt(nn);
    smax = std::max(smax, sig[j]);
  }
  const double thr = std::numeric_limits<double>::epsilon() * smax;
  for (int j = 0; j < l; ++j) {
    if (sig[j] <= thr || sig[j] == 0.0) continue;
    double ub = 0;  // u_j . b with u_j = H[:, j] / sig_j
    for (int i = 0; i < l; ++i) ub += H[i * l + j] * b[i];
    y[j] = ub / (sig[j] * sig[j]);
  }
  for (int i = 0; i < l; ++i) {
    double x = 0;
    for (int j = 0; j < l; ++j) x += V[i * l + j] * y[j];
    b[i] = x;
  }
}

// SNESSolve_NGMRES with PETSc's defaults (select/restart type "difference", basic line search,
// no nonlinear preconditioner): per iteration X^M = X - F(X), F^M = F(X^M), the combined X^A from
// the stored (X_i, F_i) and F^A = F(X^A) -- two function evaluations per iteration, as the golden
// FEvals = 2 ItNum + 1 shows.  With snes.precond the function is G = P F (not in the reference);
// the stop test and the history always use the true |F|.
void cap_ngmres(Sim& s, std::vector<double>& X)
{
  Ngmres& ng = s.snes;
  const int64_t n = s.n3;
  const int ms = ng.msize;
  ng.hist.clear();
  ng.iterations = 0;
  ng.fevals = 0;
  ng.reason = 0;
  std::vector<double> F(n), XA(n), FA(n), XM(n), FM(n), D(n), raw(n);
  std::vector<std::vector<double>> Fdot(ms, std::vector<double>(n)), Xdot(ms, std::vector<double>(n));
  std::vector<double> fnorms(ms), Q(ms * ms, 0.0), beta(ms), xi(ms);
  auto norm = [&](const std::vector<double>& a) {
    double t = 0;
    for (int64_t i = 0; i < n; ++i) t += a[i] * a[i];
    return std::sqrt(t);
  };
  auto vdot = [&](const std::vector<double>& a, const std::vector<double>& b) {
    double t = 0;
    for (int64_t i = 0; i < n; ++i) t += a[i] * b[i];
    return t;
  };
  // function value (possibly preconditioned) and the true residual norm
  auto eval = [&](const std::vector<double>& x, std::vector<double>& f, double& true_norm) {
    if (!ng.precond) {
      cap_form_function(s, x.data(), f.data());
      true_norm = norm(f);
    }
    else {
      cap_form_function(s, x.data(), raw.data());
      true_norm = norm(raw);
      cap_precondition(s, raw.data(), f.data());
    }
  };
  auto update_subspace = [&](int ivec, const std::vector<double>& f, double fn, const std::vector<double>& x) {
    Fdot[ivec] = f;
    Xdot[ivec] = x;
    fnorms[ivec] = fn;
  };
  double tF, tFM, tFA;
  eval(X, F, tF);
  double fnorm = norm(F), fminnorm = fnorm;
  ng.hist.push_back(tF);
  const double ttol = tF * ng.rtol;
  auto converged = [&](double tn) {  // SNESConvergedDefault (snorm = xnorm = 0 as NGMRES passes them)
    if (tn < ng.atol) return 2;      // SNES_CONVERGED_FNORM_ABS
    if (tn <= ttol) return 3;        // SNES_CONVERGED_FNORM_RELATIVE
    return 0;
  };
  if ((ng.reason = converged(tF))) return;
  update_subspace(0, F, fnorm, X);
  int k_restart = 1, l = 1, ivec = 0, restart_count = 0;
  for (int k = 1; k < ng.maxit + 1; ++k) {
    // x^M: basic line search, full step along -F
    for (int64_t i = 0; i < n; ++i) XM[i] = X[i] - F[i];
    eval(XM, FM, tFM);
    const double fMnorm = norm(FM);
    // combined solution (ngmresfunc.c: SNESNGMRESFormCombinedSolution_Private)
    const double nu = fMnorm * fMnorm;
    for (int i = 0; i < l; ++i) {
      xi[i] = vdot(FM, Fdot[i]);
      beta[i] = vdot(Fdot[ivec], Fdot[i]);
    }
    for (int i = 0; i < l; ++i) Q[i * ms + ivec] = Q[ivec * ms + i] = beta[i];
    for (int i = 0; i < l; ++i) beta[i] = nu - xi[i];
    if (l == 1) {
      const double h00 = Q[0] - xi[0] - xi[0] + nu;
      beta[0] = h00 != 0.0 ? beta[0] / h00 : 0.0;
    }
    else {
      std::vector<double> H(l * l), rhs(beta.begin(), beta.begin() + l);
      for (int j = 0; j < l; ++j)
        for (int i = 0; i < l; ++i) H[i * l + j] = Q[i * ms + j] - xi[i] - xi[j] + nu;
      lstsq_svd(l, H, rhs);
      std::copy(rhs.begin(), rhs.end(), beta.begin());
    }
    double alph_total = 0;
    for (int i = 0; i < l; ++i) alph_total += beta[i];
    for (int64_t j = 0; j < n; ++j) XA[j] = (1.0 - alph_total) * XM[j];
    for (int i = 0; i < l; ++i)
      for (int64_t j = 0; j < n; ++j) XA[j] += beta[i] * Xdot[i][j];
    eval(XA, FA, tFA);
    if (fminnorm > fMnorm) fminnorm = fMnorm;
    // SNESNGMRESNorms_Private
    const double fAnorm = norm(FA);
    for (int64_t j = 0; j < n; ++j) D[j] = XA[j] - XM[j];
    const double dnorm = norm(D);
    double dminnorm = -1.0;
    for (int i = 0; i < l; ++i) {
      for (int64_t j = 0; j < n; ++j) D[j] = Xdot[i][j] - XA[j];
      const double dc = norm(D);
      if (dc < dminnorm || dminnorm < 0.0) dminnorm = dc;
    }
    // SNESNGMRESSelect_Private, select_type difference
    bool selectA = true;
    if (fAnorm >= ng.gammaA * fminnorm) selectA = false;
    if (ng.epsilonB * dnorm < dminnorm || std::sqrt(fnorm) < ng.deltaB * std::sqrt(fminnorm)) {}
    else selectA = false;
    if (selectA) {
      fnorm = fAnorm; tF = tFA;
      F = FA; X = XA;
    }
    else {
      fnorm = fMnorm; tF = tFM;
      F = FM; X = XM;
    }
    // SNESNGMRESSelectRestart_Private
    bool selectRestart = false;
    if ((ng.epsilonB * dnorm > dminnorm) && (std::sqrt(fAnorm) > ng.deltaB * std::sqrt(fminnorm)) && l > 0) selectRestart = true;
    if (std::sqrt(fAnorm) > ng.gammaC * std::sqrt(fminnorm)) selectRestart = true;
    if (selectRestart) restart_count++;
    else restart_count = 0;
    ivec = k_restart % ms;
    if (restart_count >= ng.restart_it) {
      restart_count = 0;
      k_restart = 1;
      l = 1;
      ivec = 0;
      update_subspace(0, FM, fMnorm, XM);
    }
    else {
      if (l < ms) l++;
      k_restart++;
      if (fminnorm > fnorm) fminnorm = fnorm;
      update_subspace(ivec, F, fnorm, X);
    }
    if (getenv("XO_NGMRES_DEBUG"))
      fprintf(stderr, "k %d l %d fM %.4e fA %.4e fmin %.4e dnorm %.3e dmin %.3e selA %d restart %d cnt %d asum %.3e\n", k, l, fMnorm, fAnorm, fminnorm, dnorm, dminnorm, (int)selectA,
              (int)selectRestart, restart_count, alph_total);
    ng.iterations = k;
    ng.hist.push_back(tF);
    if ((ng.reason = converged(tF))) return;
  }
  ng.reason = -5;  // SNES_DIVERGED_MAX_IT
}

void step_eccapfim(Sim& s)  // eccapfim/simulation.cpp:36-130
{
  // init_iteration :46-70, Particles::prepare_storage particles.cpp:191-208
  for (auto& sp : s.sorts) {
    sp.size = 0;
    for (int64_t g = 0; g < s.nc; ++g) {
      auto& cur = sp.storage[g];
      if (cur.empty()) continue;
      sp.previous_storage[g].assign(cur.begin(), cur.end());
      sp.size += (int64_t)cur.size();
    }
  }
  std::vector<double> sol(s.E);
  cap_ngmres(s, sol);  // calc_iteration :72-104
  // after_iteration :106-129
  std::vector<double> ce(s.n3);
  curl(s, true, sol.data(), ce.data());
  for (int64_t i = 0; i < s.n3; ++i) {
    s.E[i] = 2 * sol[i] + (-1) * s.E[i];
    s.B[i] = s.B[i] + (-s.dt) * ce[i];
  }
  s.Ehk = sol;
  for (auto& sp : s.sorts) update_cells_seq(s, sp);
}

std::vector<double>* field_by_id(Sim& s, int which, int sid)
{
  switch (which) {
    case 0: return &s.E;
    case 1: return &s.B;
    case 2: return &s.B0;
    case 3: return &s.Ep;
    case 4: return &s.Ec;
    case 5: return &s.currI;
    case 6: return &s.currJe;
    case 7: return &s.sorts.at(sid).currI;
    case 8: return &s.sorts.at(sid).currJe;
    case 9: return &s.J;
    case 10: return &s.sorts.at(sid).J;
    case 11: return &s.Ehk;
  }
  return nullptr;
}

}  // namespace

// =============================================================================================
// C interface (ctypes)
// =============================================================================================
extern "C" {

// OpenMP threads for the baseline timings (1 = deterministic serial execution, the default)
void xo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int xo_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

void* xo_create(int nx, int ny, int nz, double dx, double dy, double dz, double dt, int curl_sign)
{
  Sim* s = new Sim();
  s->N[0] = nx; s->N[1] = ny; s->N[2] = nz;
  s->d[0] = dx; s->d[1] = dy; s->d[2] = dz;
  for (int a = 0; a < 3; ++a) s->L[a] = s->N[a] * s->d[a];  // utils/world.cpp:97-100 (geom = n * d)
  s->dt = dt;
  s->curl_sign = curl_sign;
  s->nc = (int64_t)nx * ny * nz;
  s->n3 = 3 * s->nc;
  for (auto* v : {&s->E, &s->Ep, &s->Ec, &s->B, &s->B0, &s->currI, &s->currJe, &s->Ehk, &s->J}) v->assign(s->n3, 0.0);
  build_matM(*s);
  build_matL_pattern(*s);
  return s;
}

void xo_destroy(void* h) { delete (Sim*)h; }

// "da_boundary_z": DM_BOUNDARY_NONE / DM_BOUNDARY_GHOSTED instead of PERIODIC; call before particles are added
void xo_set_open_z(void* h, int open)
{
  Sim& s = *(Sim*)h;
  s.open_z = open != 0;
  build_matM(s);
  build_matL_pattern(s);
}

int xo_add_species(void* h, double q, double m, double n, int Np)
{
  Sim& s = *(Sim*)h;
  s.sorts.emplace_back();
  Species& sp = s.sorts.back();
  sp.q = q; sp.m = m; sp.n = n; sp.Np = Np;
  sp.storage.resize(s.nc);
  sp.currI.assign(s.n3, 0.0);
  sp.currJe.assign(s.n3, 0.0);
  sp.J.assign(s.n3, 0.0);
  sp.previous_storage.resize(s.nc);
  return (int)s.sorts.size() - 1;
}

// commands/set_particles.cpp:19-43 + utils/particles_load.cpp:11-18,52-76 +
// commands/builders/particles_builder.cpp:16-27 (count < 0 -> the builder's expression)
long xo_set_particles_maxwell(void* h, int sid, double Tx, double Ty, double Tz, int tov, long count)
{
  Sim& s = *(Sim*)h;
  Species& sp = s.sorts.at(sid);
  if (count < 0) {
    const double frac = sp.Np / (s.d[0] * s.d[1] * s.d[2]);
    count = (long)((s.L[0] * s.L[1] * s.L[2]) * frac);
  }
  auto r01 = [&]() { return s.uni(s.gen); };
  auto tm = [&](double T) { return std::sqrt(-2.0 * (T * sp.m / mec2) * std::log(r01())); };
  long added = 0;
  for (long p = 0; p < count; ++p) {
    Point pt;
    pt.r[0] = 0.0 + r01() * (s.L[0] - 0.0);
    pt.r[1] = 0.0 + r01() * (s.L[1] - 0.0);
    pt.r[2] = 0.0 + r01() * (s.L[2] - 0.0);
    const double T[3] = {Tx, Ty, Tz};
    for (int c = 0; c < 3; ++c) {
      const double sn = std::sin(2.0 * M_PI * r01());  // sin factor drawn first
      pt.p[c] = 0.0 + sn * tm(T[c]);
    }
    if (tov) {
      const double den = std::sqrt(sp.m * sp.m + dot(pt.p, pt.p));
      for (int c = 0; c < 3; ++c) pt.p[c] /= den;
    }
    pt.id = sp.next_id++;
    if (add_particle(s, sp, pt)) ++added;
  }
  return added;
}

long xo_set_particles(void* h, int sid, const double* aos6, long count)
{
  Sim& s = *(Sim*)h;
  Species& sp = s.sorts.at(sid);
  long added = 0;
  for (long p = 0; p < count; ++p) {
    Point pt;
    for (int c = 0; c < 3; ++c) {
      pt.r[c] = aos6[6 * p + c];
      pt.p[c] = aos6[6 * p + 3 + c];
    }
    pt.id = sp.next_id++;
    if (add_particle(s, sp, pt)) ++added;
  }
  return added;
}

long xo_particle_count(void* h, int sid)
{
  Sim& s = *(Sim*)h;
  long n = 0;
  for (auto& c : s.sorts.at(sid).storage) n += (long)c.size();
  return n;
}

// cell-major order (the reference's storage traversal order)
void xo_get_particles(void* h, int sid, double* aos6, uint64_t* ids)
{
  Sim& s = *(Sim*)h;
  long k = 0;
  for (auto& cell : s.sorts.at(sid).storage)
    for (auto& pt : cell) {
      for (int c = 0; c < 3; ++c) {
        aos6[6 * k + c] = pt.r[c];
        aos6[6 * k + 3 + c] = pt.p[c];
      }
      if (ids) ids[k] = pt.id;
      ++k;
    }
}

void xo_step(void* h, int scheme)
{
  Sim& s = *(Sim*)h;
  if (scheme == 0) step_ecsim(s);
  else if (scheme == 1) step_ecsimcorr(s);
  else step_eccapfim(s);
}

void xo_get_field(void* h, int which, int sid, double* out)
{
  Sim& s = *(Sim*)h;
  auto* f = field_by_id(s, which, sid);
  std::copy(f->begin(), f->end(), out);
}

void xo_set_field(void* h, int which, int sid, const double* in)
{
  Sim& s = *(Sim*)h;
  auto* f = field_by_id(s, which, sid);
  std::copy(in, in + s.n3, f->begin());
}

void xo_solver_set(void* h, int which, double rtol, double atol, int maxit, int restart)
{
  Sim& s = *(Sim*)h;
  Solver& sv = which == 0 ? s.predict : s.correct;
  sv.rtol = rtol; sv.atol = atol; sv.maxit = maxit; sv.restart = restart;
}

void xo_solver_info(void* h, int which, int* iterations, double* rnorm, int* reason)
{
  Sim& s = *(Sim*)h;
  Solver& sv = which == 0 ? s.predict : s.correct;
  *iterations = sv.iterations; *rnorm = sv.rnorm; *reason = sv.reason;
}

// which: 0 energy (sum 1/2 m v^2 n/Np now), 1 pred_w, 2 corr_w, 3 pred_dK, 4 corr_dK, 5 lambda_dK,
//        6 stored `energy` member, 7 ||currJe - (currI + L Ec)||
double xo_scalar(void* h, int sid, int which)
{
  Sim& s = *(Sim*)h;
  Species& sp = s.sorts.at(sid);
  switch (which) {
    case 0: {  // diagnostics/energy.cpp:61-88:  K = 0.5 m (n/Np) * sum v^2
      double w = 0;
      for (auto& cell : sp.storage)
        for (auto& pt : cell) w += dot(pt.p, pt.p);
      return 0.5 * sp.m * (sp.n / (double)sp.Np) * w;
    }
    case 1: return sp.pred_w;
    case 2: return sp.corr_w;
    case 3: return sp.pred_dK;
    case 4: return sp.corr_dK;
    case 5: return sp.lambda_dK;
    case 6: return sp.energy;
    case 7: return s.last_j_diff_norm;
  }
  return 0.0;
}

// eccapfim nonlinear solver controls / results
void xo_snes_set(void* h, double atol, double rtol, double stol, int maxit, int precond, double shift)
{
  Ngmres& ng = ((Sim*)h)->snes;
  ng.atol = atol; ng.rtol = rtol; ng.stol = stol; ng.maxit = maxit; ng.precond = precond; ng.shift = shift;
}

void xo_snes_set_particle_tol(void* h, double tol) { ((Sim*)h)->snes.cn_tol = tol; }

void xo_snes_info(void* h, int* iterations, int* fevals, int* reason, double* avgit, double* avgcell)
{
  Sim& s = *(Sim*)h;
  *iterations = s.snes.iterations; *fevals = s.snes.fevals; *reason = s.snes.reason;
  *avgit = s.sorts.empty() ? 0.0 : s.sorts[0].avgit;
  *avgcell = s.sorts.empty() ? 0.0 : s.sorts[0].avgcell;
}

int xo_snes_history(void* h, double* out, int cap)
{
  Ngmres& ng = ((Sim*)h)->snes;
  const int n = (int)std::min<size_t>(ng.hist.size(), (size_t)cap);
  std::copy(ng.hist.begin(), ng.hist.begin() + n, out);
  return (int)ng.hist.size();
}

// F(x) of the eccapfim step at the present (E^n, B^n, particles); the particles afterwards hold the
// pushed state of this evaluation and `previous_storage` the start-of-step state
void xo_eccapfim_function(void* h, const double* x, double* f, int prepare)
{
  Sim& s = *(Sim*)h;
  if (prepare)
    for (auto& sp : s.sorts) {
      sp.size = 0;
      for (int64_t g = 0; g < s.nc; ++g) {
        sp.previous_storage[g].assign(sp.storage[g].begin(), sp.storage[g].end());
        sp.size += (int64_t)sp.storage[g].size();
      }
    }
  cap_form_function(s, x, f);
}

// one Crank-Nicolson step in fields that do not depend on the path (CrankNicolsonPush::process with a
// constant set_fields callback converges in its first iteration): r += dt vh, v = 2 vh - v
void xo_crank_nicolson_uniform(double dt, double qm, const double* Ep, const double* Bp, double* r, double* v)
{
  double vh[3];
  cn_mean_velocity(0.5 * dt * qm, v, Ep, Bp, vh);
  for (int c = 0; c < 3; ++c) {
    r[c] = r[c] + dt * vh[c];
    v[c] = 2.0 * vh[c] - v[c];
  }
}

// the stopping points of cell_traversal(end, start); returns their number (<= cap written)
int xo_cell_traversal(void* h, const double* end, const double* start, double* out, int cap)
{
  Sim& s = *(Sim*)h;
  std::vector<V3> pts;
  cell_traversal(s, {end[0], end[1], end[2]}, {start[0], start[1], start[2]}, pts);
  for (int i = 0; i < (int)pts.size() && i < cap; ++i)
    for (int c = 0; c < 3; ++c) out[3 * i + c] = pts[i][c];
  return (int)pts.size();
}

// Stand-alone pieces (used to test individual kernels) --------------------------------------
void xo_deposit(void* h)  // clear + moments at the current particle positions with current B
{
  Sim& s = *(Sim*)h;
  clear_sources(s);
  fill_ecsim_current(s);
}

long xo_csr_nnz(void* h, int which) { Sim& s = *(Sim*)h; return (long)(which == 0 ? s.matL : s.matM).val.size(); }

void xo_csr_export(void* h, int which, int64_t* rowptr, int32_t* col, double* val)
{
  Sim& s = *(Sim*)h;
  const Csr& A = which == 0 ? s.matL : s.matM;
  std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
  std::copy(A.col.begin(), A.col.end(), col);
  std::copy(A.val.begin(), A.val.end(), val);
}

// y = (which&1 ? L x : 0) + (which&2 ? M x : 0)
void xo_spmv(void* h, int which, const double* x, double* y)
{
  Sim& s = *(Sim*)h;
  std::vector<double> t(s.n3, 0.0);
  std::fill(y, y + s.n3, 0.0);
  if (which & 1) { csr_mult(s.matL, x, t.data(), s.n3); for (int64_t i = 0; i < s.n3; ++i) y[i] += t[i]; }
  if (which & 2) { csr_mult(s.matM, x, t.data(), s.n3); for (int64_t i = 0; i < s.n3; ++i) y[i] += t[i]; }
}

void xo_curl(void* h, int positive, const double* f, double* out) { curl(*(Sim*)h, positive != 0, f, out); }

void xo_interpolate(void* h, const double* r, double* Ep, double* Bp)
{
  Sim& s = *(Sim*)h;
  W w;
  weights(s, r, w);
  interpolate_E_s1(s, s.Ep, w, Ep);
  interpolate_B_s1(s, s.B, w, Bp);
}

void xo_boris_update_vEB(double dt, double qm, const double* Ep, const double* Bp, double* v) { update_vEB(dt, qm, Ep, Bp, v); }

void xo_esirkepov(void* h, const double* old_r, const double* new_r, double alpha, double* J)
{
  Sim& s = *(Sim*)h;
  std::vector<double> j(s.n3, 0.0);
  esirkepov(s, old_r, new_r, alpha, j);
  std::copy(j.begin(), j.end(), J);
}

}  // extern "C"
