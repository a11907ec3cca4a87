// common.cuh -- context, layout conventions and small device helpers shared by every kernel file.
//
// Layout conventions (DESIGN.md "Data layout in HBM"):
//   * grid vectors: [zl + GZ][y][x][c] fp64, zl in [-GZ, nzl + GZ); x/y are never decomposed and
//     wrap by index arithmetic inside the kernels; the GZ ghost planes per side mirror DMDA's
//     local vectors (stencil width 4 in the reference, utils/world.h:25; 3 is what the kernels need).
//   * operator L: coef[k][node], k < NCOEF = 369 fixed (c1, c2, offset) slots per owned node.
//   * particles: SoA x,y,z,vx,vy,vz (+ optional id), sorted by bin = (cell << 3 | octant), cells over
//     nzl + 2 planes (plane 0 and nzl + 1 collect particles that left the slab).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/xpic_b200.h"

namespace xb {

constexpr int GZ = 3;  // ghost planes per side of every grid vector

void set_error(const std::string& msg);

#define XB_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      xb::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                    std::to_string(__LINE__));                                                          \
      return 1;                                                                                         \
    }                                                                                                   \
  } while (0)

#define XB_CHECK(expr)    \
  do {                    \
    int rc_ = (expr);     \
    if (rc_) return rc_;  \
  } while (0)

#define XB_FAIL(msg)      \
  do {                    \
    xb::set_error(msg);   \
    return 1;             \
  } while (0)

struct Grid {
  int nx, ny, nz;  // global cells
  int z0, nzl;     // owned planes [z0, z0 + nzl)
  int rank, nranks;
  double dx, dy, dz, dt;
  double inv_dx, inv_dy, inv_dz;
  int exact_inv;  // bit a set: d[a] is a power of two, r * inv_d == r / d exactly
  double Lx, Ly, Lz;
  int curl_sign;
  int open_z;  // da_boundary_z is not periodic: no nodes below plane 0 / above plane nz - 1 (xb_grid.boundary)
  int64_t plane;  // nx * ny
  int64_t ncl;    // owned cells = plane * nzl
  int64_t nown;   // 3 * ncl
  int64_t ntot;   // 3 * plane * (nzl + 2 GZ)
  int64_t own0;   // offset of the first owned element = 3 * plane * GZ

  __host__ __device__ inline int64_t vidx(int x, int y, int zl, int c) const
  {
    return ((((int64_t)(zl + GZ)) * ny + y) * nx + x) * 3 + c;
  }
};

// geometry of a StepPresets command (src/utils/geometries.h): kind < 0: none
struct Geometry {
  int kind = -1;
  double p[6] = {0, 0, 0, 0, 0, 0};  // box: min[3], max[3]; cylinder: center[3], radius, height
};

__host__ __device__ inline bool within_geometry(const Geometry& ge, double x, double y, double z)
{
  if (ge.kind == XB_GEOMETRY_BOX)  // WithinBox, src/utils/geometries.cpp:3-9
    return (ge.p[0] <= x && x < ge.p[3]) && (ge.p[1] <= y && y < ge.p[4]) && (ge.p[2] <= z && z < ge.p[5]);
  const double dx = x - ge.p[0], dy = y - ge.p[1], dz = z - ge.p[2];  // WithinCylinder, :12-19
  return (fabs(dz) < 0.5 * ge.p[4]) && ((dx * dx + dy * dy) <= ge.p[3] * ge.p[3]);
}

struct Solver {
  double rtol = 1e-7, atol = 1e-7;  // src/impls/ecsim/simulation.h:15-16
  int maxit = 100, restart = 30;    // :18, PETSc GMRES default restart
  int precond = 0;
  int iterations = 0, reason = 0;
  double rnorm = 0.0;
};

// eccapfim's nonlinear solve: tolerances of src/impls/eccapfim/simulation.h:14-19, per-particle
// Picard tolerance 0.5 * atol and 30 iterations (src/impls/eccapfim/particles.cpp:99-101)
struct Nonlinear {
  double atol = 1e-7, rtol = 1e-7, stol = 1e-7;
  int maxit = 1000;
  int depth = 10;         // Anderson history
  int cheb_degree = 12;   // Chebyshev degree of the residual preconditioner (0: none)
  double cn_tol = 0.5e-7;
  int cn_maxit = 30;
  bool warm_start = true;  // Picard iterations of evaluation k + 1 start from the velocities of evaluation k
  bool pn_valid = false;
  int iterations = 0, fevals = 0, reason = 0;
  double fnorm = 0.0, avg_cn = 0.0, avg_cells = 0.0;
  std::vector<double> hist;
  int64_t particles_per_eval = 0;
  bool profile = false;   // CUDA-event timing of every particle pass
  std::vector<cudaEvent_t> events;
  size_t events_used = 0;
  double push_ms = 0.0;
  int64_t push_evals = 0;
};

struct MigrateBuffers {  // multi-rank only (migrate.cu)
  double* send[2][7] = {{nullptr}};   // [0] to the rank below, [1] to the rank above; 6 SoA arrays + ids
  double* recv[2][7] = {{nullptr}};   // [0] from below, [1] from above
  double* ghost[2][6] = {{nullptr}};  // copies of the neighbours' boundary-plane particles
  double* ghost_rec[2] = {nullptr, nullptr};  // their field records, SoA [12][ghost_cap] (cross-check pipeline only)
  int32_t* ghost_bins[2] = {nullptr, nullptr};      // their bin tables, rebased to the copy
  int32_t* ghost_bins_raw[2] = {nullptr, nullptr};  // as received (offsets into the owner's arrays)
  int32_t* recv_key[2] = {nullptr, nullptr};
  int64_t cap = 0, ghost_cap = 0;
  int64_t nghost[2] = {0, 0};
  unsigned long long* counts_dev = nullptr;  // [0], [1] leavers down / up of this sort; [2] particles lost so far
  unsigned long long* table_dev = nullptr;   // all-gathered rows of counts and limits, one per rank
  unsigned long long* table_host = nullptr;  // pinned
  cudaEvent_t sorted = nullptr, ghosts_here = nullptr;
};

struct Species {
  double q, m, n;
  int Np;
  int64_t count = 0, capacity = 0;
  int cur = 0;                    // which of the two SoA buffers is live
  double* p[2][6] = {{nullptr}};  // x,y,z,vx,vy,vz
  double* p_alloc[2][6] = {{nullptr}};  // what cudaMalloc returned (p is skewed, species_alloc)
  uint64_t* id[2] = {nullptr, nullptr};
  int32_t* key = nullptr;        // bin of every particle (capacity)
  double* rec = nullptr;         // field record of every particle, SoA [12][capacity]: cross-check pipeline only, allocated on first use (deposit.cu)
  int32_t* bin_start = nullptr;  // nbins + 1, valid after sort
  double* currI = nullptr;       // per-sort currents (ghosted grid vectors)
  double* currJe = nullptr;
  double* rho[2] = {nullptr, nullptr};  // charge density of the last two collections (diagnostics.cu), component 0 of a grid vector
  int rho_cur = 0;
  uint64_t next_id = 0;
  bool sorted = false;
  MigrateBuffers* mig = nullptr;
  // ecsimcorr::Particles scalars (src/impls/ecsimcorr/particles.h:33-38)
  double energy = 0, pred_w = 0, corr_w = 0, pred_dK = 0, corr_dK = 0, lambda_dK = 0;
  double cap_unit = 1.0;  // eccapfim: unit of the fixed-point current accumulators of this step (a power of two)
};

struct Comm;  // comm.cu (NCCL over NVLink, loaded lazily)

struct StageClock {
  double seconds[XB_STAGE_COUNT] = {0};
  int64_t calls[XB_STAGE_COUNT] = {0};
};

}  // namespace xb

struct xb_ctx {
  xb::Grid g;
  int device = 0;
  int sm_count = 0;
  int* work_counter = nullptr;  // k_cell_moments_ws: next group of cells to hand out
  int ws_backoff_ns = 64;  // k_cell_moments_ws: pause of a waiting warp between two polls of its mbarrier (xb_set_option 5)
  bool track_ids = false;
  bool deterministic = false;  // canonical particle order inside every bin even without ids (costs one more pass)
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // exchanges that run beside the main stream's kernels (ghost particles, halo planes)
  cudaStream_t host_stream = nullptr;  // host<->device copies of xb_step_host that overlap with the particle stages
  cudaEvent_t copy_done = nullptr;
  cudaEvent_t b_ready = nullptr;       // xb_step_host: B^n has arrived (awaited before the moment deposition)
  bool b_pending = false;
  cudaEvent_t halo_ready = nullptr, halo_done = nullptr;  // halo_begin / halo_end
  cudaEvent_t blocks_ready = nullptr, blocks_here = nullptr;  // exchange of the boundary planes' cell blocks (deposit.cu)
  bool halo_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // CUDA-event pairs (start, stop) recorded around every launch group of a kernel family inside the step
  // (xb_family_profile); family 3 is the operator SpMV (xb_spmv_profile)
  bool family_profile = false;
  std::vector<cudaEvent_t> prof_events[XB_FAMILY_COUNT];
  size_t prof_used[XB_FAMILY_COUNT] = {0};
  // named grid vectors (ghosted)
  double *E = nullptr, *B = nullptr, *B0 = nullptr, *Ep = nullptr, *Ec = nullptr, *currI = nullptr, *currJe = nullptr;
  double *rhs = nullptr, *tmp = nullptr, *tmp2 = nullptr;
  // operator L, stencil layout
  double* coef = nullptr;   // blocked [tile][k][t], see stencil.cuh
  int64_t coef_elems = 0;
  bool coef_valid = false;
  int deposit_variant = 0;  // 0: fused DMMA kernel, variant tiles (default); cross-checks: 4 same with the fold in shared memory, 3 no role split, 2 round-1 DMMA pipeline, 1 scalar FMA
  bool fused_attr_set = false, deposit_attr_set = false, esirkepov_attr_set = false, cap_attr_set = false;  // per context = per device
  int cap_variant = 0;  // eccapfim particle pass: 0 CTA task machine (default), 1 thread per particle (cross-check)
  int esirkepov_variant = 0;  // 0: DMMA cell blocks + gather (default), 1: per-particle global reductions (cross-check)
  // deposit staging (cell blocks)
  double* stage = nullptr;
  int64_t stage_cells = 0;
  bool stage_tiles = false;  // what the last deposit_cells left there: variant tiles (STAGE_CELL doubles per cell) or folded cell blocks (BLOCK_ALL)
  int batch_planes = 0;  // 0: the staging area holds the whole slab; P > 0: batches of P planes through P + 2 staging planes
  // Krylov workspace
  std::vector<double*> V;  // restart + 1 basis vectors (ghosted)
  double* Z = nullptr;     // preconditioned direction (ghosted)
  double *cheb_r = nullptr, *cheb_d = nullptr, *cheb_Md = nullptr;  // Chebyshev work vectors (ghosted)
  double* ksp_u = nullptr;  // sum y_i V_i before right preconditioning
  double* red_partial = nullptr;  // [RED_BLOCKS][RED_MAXV]
  double* red_out = nullptr;      // device results
  double* red_host = nullptr;     // pinned
  double* hcoef_dev = nullptr;    // small coefficient vectors for multi-axpy
  // sort workspace
  int32_t* hist = nullptr;
  int32_t* cursor = nullptr;
  int32_t* scan_tmp = nullptr;
  int64_t nbins = 0;
  // host staging
  double* pinned = nullptr;
  size_t pinned_bytes = 0;

  xb::Geometry remove;  // RemoveParticles: applied by the next re-binning of the sort it was set for (kind >= 0)
  unsigned long long* removed_dev = nullptr;  // { particles, kinetic energy as double bits } of that re-binning
  std::vector<xb::Species> sorts;
  xb::Solver solver[2];
  xb::StageClock clock;
  xb::Comm* comm = nullptr;
  int64_t launches = 0;
  double j_diff_norm = 0.0;
  // eccapfim (eccapfim.cu), allocated on first use
  xb::Nonlinear nl;
  double *cap_x = nullptr, *cap_F = nullptr, *cap_g = nullptr, *cap_rhs0 = nullptr, *cap_J = nullptr;
  unsigned long long* cap_counters = nullptr;
};

namespace xb {

constexpr int RED_BLOCKS = 1184;  // 148 SMs x 8
constexpr int RED_THREADS = 256;
constexpr int RED_MAXV = 32;

// ---- fields.cu -------------------------------------------------------------------------------
int halo_fill(xb_ctx* c, double* v, int width);                 // DMGlobalToLocal(INSERT)
int halo_begin(xb_ctx* c, double* v, int width);                // DMGlobalToLocalBegin: the exchange runs beside the main stream
int halo_end(xb_ctx* c);                                        // DMGlobalToLocalEnd
int halo_reduce(xb_ctx* c, double* v, int width_lo, int width_hi);  // DMLocalToGlobal(ADD)
int vec_zero(xb_ctx* c, double* v);                              // whole ghosted vector
int vec_copy_owned(xb_ctx* c, const double* src, double* dst);
int curl_apply(xb_ctx* c, bool positive, const double* f, double* out, double scale, bool accumulate);
int build_rhs(xb_ctx* c, const double* curr, double* rhs);       // 2E - dt curr + dt curl^-(B - B0)
int final_update(xb_ctx* c, const double* Ehalf);                // E = 2 Eh - E ; B -= dt curl^+ Eh
int final_update_into(xb_ctx* c, const double* Ehalf, double* Eout, double* Bout);
int dots(xb_ctx* c, int nv, const double* const* vs, const double* w, double* host_out);  // host_out[i] = vs[i].w
int field_sums(xb_ctx* c, const double* v, double* out4);  // component sums and sum of squares, all ranks
int axpy_multi(xb_ctx* c, int nv, const double* const* vs, const double* coef_host, double* w);  // w += sum coef_i vs_i
int axpy_multi_scaled(xb_ctx* c, int nv, const double* const* vs, const double* coef_host, double* w, double alpha);  // w = alpha (w + sum coef_i vs_i)
int scale_into(xb_ctx* c, const double* w, double alpha, double* out);                            // out = alpha w
int axpby(xb_ctx* c, double a, const double* x, double b, double* y);                             // y = a x + b y
int upload_owned(xb_ctx* c, const double* host, double* dev);
int download_owned(xb_ctx* c, const double* dev, double* host);

// ---- spmv.cu ---------------------------------------------------------------------------------
int spmv(xb_ctx* c, int op, double* x_ghosted, double* y);  // fills x's halo (width 2) itself

// ---- krylov.cu -------------------------------------------------------------------------------
int gmres(xb_ctx* c, int which, int op, const double* b, double* x);
int krylov_prepare(xb_ctx* c);  // allocate the Krylov workspace up front (not inside the first timed solve)

// ---- particles.cu ----------------------------------------------------------------------------
int species_alloc(xb_ctx* c, Species& s, int64_t capacity);
void species_free(Species& s);
void migrate_free(Species& s);  // migrate.cu
int particles_sort(xb_ctx* c, Species& s, double dt_move);   // r += v dt_move, wrap, re-bin
int fields_damping(xb_ctx* c, const Geometry& ge, double coefficient, double* damped_energy);  // fields.cu
int push_second(xb_ctx* c, Species& s, const double* Eh, const double* B);
int push_second_work(xb_ctx* c, Species& s, const double* Eh, const double* B, double* pred_w);
int kinetic_energy(xb_ctx* c, Species& s, double* sum_v2, double* K);
int scale_velocities(xb_ctx* c, Species& s, double lambda);
int particle_moments(xb_ctx* c, Species& s, double* out5);
int particles_generate(xb_ctx* c, Species& s, int64_t total, const double* T, uint64_t seed, int tov, int64_t* added);

// ---- deposit.cu ------------------------------------------------------------------------------
int deposit_moments(xb_ctx* c);  // currI (+ per sort) and coef from all sorts
int coef_convert(xb_ctx* c, double* plain_dev, bool to_blocked);

// ---- esirkepov.cu ----------------------------------------------------------------------------
int push_first_corr(xb_ctx* c, Species& s);
int push_second_corr(xb_ctx* c, Species& s, const double* Eh, const double* B);

// ---- esirkepov_mma.cu (atomic-free tensor-core form, the default) ----------------------------
int push_first_corr_mma(xb_ctx* c, Species& s);
int push_second_corr_mma(xb_ctx* c, Species& s, const double* Eh, const double* B);

// ---- eccapfim.cu -----------------------------------------------------------------------------
int cap_alloc(xb_ctx* c);                                // the scheme's grid vectors, on first use
int cap_prepare(xb_ctx* c);                              // init_iteration: sort if needed, rhs0
int cap_form_function(xb_ctx* c, double* x, double* F);  // form_iteration: F(x), x ghosted
int cap_solve(xb_ctx* c);                                // calc_iteration
int cap_finish(xb_ctx* c);                               // after_iteration
int cap_read_counters(xb_ctx* c);                        // averages of the last evaluation -> c->nl

// ---- diagnostics.cu --------------------------------------------------------------------------
int charge_density(xb_ctx* c, Species& s);                               // ParticlesChargeDensity::collect
int charge_conservation(xb_ctx* c, int which_current, double* norms);   // ChargeConservation::add_columns
int momentum(xb_ctx* c, Species& s, double* out6);                      // MomentumConservation::calculate
int distribution_moment(xb_ctx* c, Species& s, int moment, const int32_t* start, const int32_t* size);  // DistributionMoment::collect -> c->tmp2 (+ c->tmp)
int moment_size(int moment);
void velocity_region(const double dv[2], const double vmin[2], const double vmax[2], int32_t* vstart, int32_t* vsize);
int velocity_distribution(xb_ctx* c, Species& s, int projector, const Geometry& ge, const double dv[2], const double vmin[2], const double vmax[2],
                          double* host_out);

// ---- api.cu: in-step kernel-family timing ------------------------------------------------------
int prof_begin(xb_ctx* c, int family);
int prof_end(xb_ctx* c, int family);

// ---- launch bookkeeping ----------------------------------------------------------------------
#define XB_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                          \
    XB_CUDA(cudaGetLastError());                                                \
  } while (0)

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int wrapi(int i, int n)
{
  i %= n;
  return i < 0 ? i + n : i;
}

}  // namespace xb
