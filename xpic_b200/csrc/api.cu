// api.cu -- the extern "C" surface declared in include/xpic_b200.h and the step orchestration
// (ecsim::Simulation::timestep_implementation, src/impls/ecsim/simulation.cpp:145-155, and
// ecsimcorr::Simulation::timestep_implementation, src/impls/ecsimcorr/simulation.cpp:21-32).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "comm.cuh"
#include "common.cuh"
#include "stencil.cuh"

namespace xb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

int migrate_and_sort(xb_ctx* c, Species& s, double dt_move);  // migrate.cu

int prof_begin(xb_ctx* c, int family)
{
  if (!c->family_profile) return 0;
  auto& ev = c->prof_events[family];
  if (c->prof_used[family] + 2 > ev.size())
    for (int i = 0; i < 256; ++i) {
      cudaEvent_t e;
      XB_CUDA(cudaEventCreate(&e));
      ev.push_back(e);
    }
  XB_CUDA(cudaEventRecord(ev[c->prof_used[family]], c->stream));
  return 0;
}

int prof_end(xb_ctx* c, int family)
{
  if (!c->family_profile) return 0;
  XB_CUDA(cudaEventRecord(c->prof_events[family][c->prof_used[family] + 1], c->stream));
  c->prof_used[family] += 2;
  return 0;
}

static int sort_species(xb_ctx* c, Species& s, double dt_move)
{
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT));
  if (c->g.nranks > 1)
    XB_CHECK(migrate_and_sort(c, s, dt_move));
  else
    XB_CHECK(particles_sort(c, s, dt_move));
  return prof_end(c, XB_FAMILY_SORT);
}

struct StageTimer {
  xb_ctx* c;
  int stage;
  StageTimer(xb_ctx* c_, int st) : c(c_), stage(st) { cudaEventRecord(c->ev0, c->stream); }
  int finish()
  {
    XB_CUDA(cudaEventRecord(c->ev1, c->stream));
    XB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    XB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->clock.seconds[stage] += 1e-3 * ms;
    c->clock.calls[stage] += 1;
    return 0;
  }
};

static int ensure_sorted(xb_ctx* c)
{
  for (auto& s : c->sorts)
    if (!s.sorted) XB_CHECK(sort_species(c, s, 0.0));
  return 0;
}

static int zero_owned_and_ghosts(xb_ctx* c, double* v) { return vec_zero(c, v); }

// ---- stages ---------------------------------------------------------------------------------
static int stage_clear(xb_ctx* c, int scheme)
{
  XB_CHECK(zero_owned_and_ghosts(c, c->currI));
  for (auto& s : c->sorts) XB_CHECK(zero_owned_and_ghosts(c, s.currI));
  if (scheme == XB_ECSIMCORR) {
    XB_CHECK(zero_owned_and_ghosts(c, c->currJe));
    for (auto& s : c->sorts) {
      XB_CHECK(zero_owned_and_ghosts(c, s.currJe));
      XB_CHECK(kinetic_energy(c, s, nullptr, &s.energy));  // ecsimcorr/simulation.cpp:44-45
    }
  }
  return 0;
}

static int stage_first_push(xb_ctx* c, int scheme)
{
  for (auto& s : c->sorts) {
    if (scheme == XB_ECSIM) {
      XB_CHECK(sort_species(c, s, c->g.dt));  // r += v dt, wrap, re-bin (ecsim/particles.cpp:21-31)
    }
    else {
      if (c->esirkepov_variant == 1)
        XB_CHECK(push_first_corr(c, s));
      else
        XB_CHECK(push_first_corr_mma(c, s));
      XB_CHECK(sort_species(c, s, 0.0));
    }
  }
  if (c->b_pending) {  // xb_step_host: B^n travelled underneath the re-binning
    XB_CUDA(cudaStreamWaitEvent(c->stream, c->b_ready, 0));
    c->b_pending = false;
  }
  XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS));
  XB_CHECK(deposit_moments(c));
  return prof_end(c, XB_FAMILY_MOMENTS);
}

static int solve_checked(xb_ctx* c, int which, int op, const double* curr, double* out)
{
  XB_CHECK(build_rhs(c, curr, c->rhs));
  XB_CHECK(gmres(c, which, op, c->rhs, out));
  if (c->solver[which].reason < 0)
    XB_FAIL(std::string("KSPSolve(") + (which == 0 ? "predict" : "correct") + ") did not converge: reason " +
            std::to_string(c->solver[which].reason) + ", iterations " + std::to_string(c->solver[which].iterations));
  return 0;
}

static int stage_second_push(xb_ctx* c, int scheme)
{
  XB_CHECK(halo_fill(c, c->Ep, 1));
  XB_CHECK(halo_fill(c, c->B, 1));
  for (auto& s : c->sorts) {
    if (scheme == XB_ECSIM) {
      XB_CHECK(prof_begin(c, XB_FAMILY_PUSH2));
      XB_CHECK(push_second(c, s, c->Ep, c->B));
      XB_CHECK(prof_end(c, XB_FAMILY_PUSH2));
      // positions did not change: the reference's second update_cells is a no-op here
    }
    else {
      if (c->esirkepov_variant == 1)
        XB_CHECK(push_second_corr(c, s, c->Ep, c->B));
      else
        XB_CHECK(push_second_corr_mma(c, s, c->Ep, c->B));
      XB_CHECK(halo_reduce(c, s.currJe, GZ, GZ));
      const double one = 1.0;
      const double* vs[1] = {s.currJe};
      XB_CHECK(axpy_multi(c, 1, vs, &one, c->currJe));  // ecsimcorr/particles.cpp:89
      XB_CHECK(sort_species(c, s, 0.0));
    }
  }
  return 0;
}

static int stage_final(xb_ctx* c, int scheme)
{
  if (scheme == XB_ECSIMCORR) {
    for (auto& s : c->sorts) {  // ecsimcorr/particles.cpp:93-126
      const double* vs[1] = {s.currJe};
      XB_CHECK(dots(c, 1, vs, c->Ec, &s.corr_w));
      const double K0 = s.energy;
      double K = 0.0;
      XB_CHECK(kinetic_energy(c, s, nullptr, &K));
      const double lambda2 = 1.0 + c->g.dt * (s.corr_w - s.pred_w) / K;
      const double lambda = std::sqrt(lambda2);
      XB_CHECK(scale_velocities(c, s, lambda));
      s.lambda_dK = (lambda2 - 1.0) * K;
      s.pred_dK = K - K0;
      s.corr_dK = lambda2 * K - K0;
      s.energy = lambda2 * K;
    }
    {  // ||currJe - (currI + L Ec)||, logged by the reference (ecsimcorr/simulation.cpp:74-82)
      XB_CHECK(spmv(c, XB_OP_L, c->Ec, c->tmp));
      const double one = 1.0, mone = -1.0;
      const double* a1[1] = {c->tmp};
      XB_CHECK(axpy_multi(c, 1, a1, &one, c->currI));
      XB_CHECK(vec_copy_owned(c, c->currJe, c->tmp));
      const double* a2[1] = {c->currI};
      XB_CHECK(axpy_multi(c, 1, a2, &mone, c->tmp));
      const double* a3[1] = {c->tmp};
      double n2 = 0.0;
      XB_CHECK(dots(c, 1, a3, c->tmp, &n2));
      c->j_diff_norm = std::sqrt(n2);
    }
    std::swap(c->Ep, c->Ec);  // VecSwap(Ep, Ec), :84
  }
  XB_CHECK(final_update(c, c->Ep));
  return 0;
}

static int run_stage(xb_ctx* c, int scheme, int stage)
{
  StageTimer t(c, stage);
  if (scheme == XB_ECCAPFIM) {  // eccapfim/simulation.cpp:36-44: init_iteration, calc_iteration, after_iteration
    switch (stage) {
      case XB_STAGE_CLEAR_SOURCES: XB_CHECK(cap_prepare(c)); break;
      case XB_STAGE_ADVANCE_FIELDS: XB_CHECK(cap_solve(c)); break;
      case XB_STAGE_FINAL_UPDATE: XB_CHECK(cap_finish(c)); break;
      default: break;
    }
    return t.finish();
  }
  switch (stage) {
    case XB_STAGE_CLEAR_SOURCES: XB_CHECK(stage_clear(c, scheme)); break;
    case XB_STAGE_FIRST_PUSH: XB_CHECK(stage_first_push(c, scheme)); break;
    case XB_STAGE_ADVANCE_FIELDS: XB_CHECK(solve_checked(c, XB_SOLVER_PREDICT, XB_OP_A, c->currI, c->Ep)); break;
    case XB_STAGE_SECOND_PUSH: XB_CHECK(stage_second_push(c, scheme)); break;
    case XB_STAGE_CORRECT_FIELDS:
      if (scheme == XB_ECSIMCORR) XB_CHECK(solve_checked(c, XB_SOLVER_CORRECT, XB_OP_M, c->currJe, c->Ec));
      break;
    case XB_STAGE_FINAL_UPDATE: XB_CHECK(stage_final(c, scheme)); break;
    default: XB_FAIL("unknown stage");
  }
  return t.finish();
}

static double* named_vector(xb_ctx* c, int which, int sid)
{
  switch (which) {
    case XB_E: return c->E;
    case XB_B: return c->B;
    case XB_B0: return c->B0;
    case XB_EP: return c->Ep;
    case XB_EC: return c->Ec;
    case XB_CURRI: return c->currI;
    case XB_CURRJE: return c->currJe;
    case XB_CURRI_SORT: return sid >= 0 && sid < (int)c->sorts.size() ? c->sorts[sid].currI : nullptr;
    case XB_CURRJE_SORT: return sid >= 0 && sid < (int)c->sorts.size() ? c->sorts[sid].currJe : nullptr;
    case XB_J: return c->cap_J;
    case XB_J_SORT: return sid >= 0 && sid < (int)c->sorts.size() ? c->sorts[sid].currI : nullptr;  // eccapfim keeps Particles::J there
    case XB_EHK: return c->cap_x;
  }
  return nullptr;
}

}  // namespace xb

using namespace xb;

#define XB_API_BEGIN(ctx)                                   \
  if (!(ctx)) {                                             \
    xb::set_error("null context");                          \
    return 1;                                               \
  }                                                         \
  if (cudaSetDevice((ctx)->device) != cudaSuccess) {        \
    xb::set_error("cudaSetDevice failed");                  \
    return 1;                                               \
  }

extern "C" {

const char* xb_last_error(void) { return g_error.c_str(); }
int xb_version(void) { return 100; }
int xb_operator_ncoef(void) { return NCOEF; }

int xb_operator_coef_info(int k, int* c1, int* c2, int* dx, int* dy, int* dz)
{
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const int base = pair_base(a, b), n = pair_size(a, b);
      if (k >= base && k < base + n) {
        const DRange rx = drange(a, b, 0), ry = drange(a, b, 1), rz = drange(a, b, 2);
        const int r = k - base;
        *c1 = a;
        *c2 = b;
        *dx = rx.lo + r % rx.n;
        *dy = ry.lo + (r / rx.n) % ry.n;
        *dz = rz.lo + r / (rx.n * ry.n);
        return 0;
      }
    }
  XB_FAIL("coefficient index out of range");
}

int xb_comm_unique_id(void* out128) { return comm_unique_id(out128); }

int xb_destroy(xb_ctx* c);
}

static int create_impl(xb_ctx* c, const xb_grid* gr, const void* uid)
{
  c->device = gr->device;
  c->track_ids = gr->track_ids != 0;
  Grid& g = c->g;
  g.nx = gr->n[0]; g.ny = gr->n[1]; g.nz = gr->n[2];
  g.dx = gr->d[0]; g.dy = gr->d[1]; g.dz = gr->d[2];
  g.dt = gr->dt;
  g.inv_dx = 1.0 / g.dx; g.inv_dy = 1.0 / g.dy; g.inv_dz = 1.0 / g.dz;
  g.exact_inv = 0;
  for (int a = 0; a < 3; ++a) {
    int e = 0;
    if (std::frexp(gr->d[a], &e) == 0.5) g.exact_inv |= 1 << a;
  }
  g.Lx = g.nx * g.dx; g.Ly = g.ny * g.dy; g.Lz = g.nz * g.dz;  // utils/world.cpp:97-100
  g.curl_sign = gr->curl_sign < 0 ? -1 : +1;
  if (gr->boundary[0] != XB_BOUNDARY_PERIODIC || gr->boundary[1] != XB_BOUNDARY_PERIODIC)
    XB_FAIL("xb_create: da_boundary_x / da_boundary_y other than DM_BOUNDARY_PERIODIC are not covered by this build (z may be open)");
  g.open_z = gr->boundary[2] != XB_BOUNDARY_PERIODIC ? 1 : 0;
  g.rank = gr->rank; g.nranks = gr->nranks;
  const int base = g.nz / g.nranks, rem = g.nz % g.nranks;
  g.nzl = base + (g.rank < rem ? 1 : 0);
  g.z0 = g.rank * base + std::min(g.rank, rem);
  if (g.nzl < 1) XB_FAIL("xb_create: more ranks than z planes");
  if (g.nranks > 1 && g.nzl < GZ) XB_FAIL("xb_create: slab thinner than the ghost width");
  g.plane = (int64_t)g.nx * g.ny;
  g.ncl = g.plane * g.nzl;
  g.nown = 3 * g.ncl;
  g.ntot = 3 * g.plane * (g.nzl + 2 * GZ);
  g.own0 = 3 * g.plane * GZ;
  if ((g.nzl + 2) * g.plane * 8 >= (int64_t)0x7fffffff) XB_FAIL("xb_create: slab too large for 32-bit bin keys");
  if (g.ntot >= (int64_t)0x7fffffff) XB_FAIL("xb_create: slab too large for 32-bit field offsets");

  XB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {  // the copy stream carries host copies and, in multi-rank runs, the exchanges that run beside the main stream's
     // kernels (ghost particles, halo planes): highest priority, so that its kernels get the first SM that frees up
    int lo = 0, hi = 0;
    XB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    XB_CUDA(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi));
  }
  XB_CUDA(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
  XB_CUDA(cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
  XB_CUDA(cudaEventCreateWithFlags(&c->b_ready, cudaEventDisableTiming));
  XB_CUDA(cudaEventCreate(&c->ev0));
  XB_CUDA(cudaEventCreate(&c->ev1));
  XB_CUDA(cudaEventCreate(&c->ev2));
  XB_CUDA(cudaEventCreate(&c->ev3));
  for (double** v : {&c->E, &c->B, &c->B0, &c->Ep, &c->Ec, &c->currI, &c->currJe, &c->rhs, &c->tmp, &c->tmp2}) {
    XB_CUDA(cudaMalloc(v, sizeof(double) * g.ntot));
    XB_CUDA(cudaMemset(*v, 0, sizeof(double) * g.ntot));
  }
  c->coef_elems = make_tilemap(g.nx, g.ny, g.nzl).ntiles() * (int64_t)(NCOEF * TILE_NODES);
  XB_CUDA(cudaMalloc(&c->coef, sizeof(double) * c->coef_elems));
  XB_CUDA(cudaMemset(c->coef, 0, sizeof(double) * c->coef_elems));
  c->stage_cells = g.nranks == 1 ? g.ncl : (g.nzl + 2) * g.plane;
  {
    // 15.9 KB of variant-tile staging per cell (the folded cell blocks of the cross-check variants need 10.6 KB): the whole slab when that stays inside the budget (72 GB of the 180 GB,
    // XPIC_STAGE_GB overrides), else batches of P planes through P + 2 staging planes (deposit_moments)
    const char* env = std::getenv("XPIC_STAGE_GB");
    const double budget = (env && *env ? std::atof(env) : 72.0) * 1e9;
    const double per_plane = (double)g.plane * STAGE_CELL * sizeof(double);
    if ((double)c->stage_cells * STAGE_CELL * sizeof(double) > budget) {
      const int P = (int)(budget / per_plane) - 2;
      if (P < 1) XB_FAIL("xb_create: one plane of cell blocks does not fit the staging budget (XPIC_STAGE_GB)");
      if ((g.plane % CELL_GROUP) != 0) XB_FAIL("xb_create: batched staging needs nx * ny to be a multiple of 4");
      c->batch_planes = P < g.nzl ? P : g.nzl;
      c->stage_cells = (int64_t)(c->batch_planes + 2) * g.plane;
    }
  }
  const int64_t groups = (c->stage_cells + CELL_GROUP - 1) / CELL_GROUP;
  XB_CUDA(cudaMalloc(&c->stage, sizeof(double) * groups * CELL_GROUP * STAGE_CELL));
  XB_CUDA(cudaMemset(c->stage, 0, sizeof(double) * groups * CELL_GROUP * STAGE_CELL));
  c->nbins = (int64_t)(g.nzl + 2) * g.plane * 8;
  XB_CUDA(cudaMalloc(&c->hist, sizeof(int32_t) * c->nbins));
  XB_CUDA(cudaMalloc(&c->cursor, sizeof(int32_t) * c->nbins));
  XB_CUDA(cudaMalloc(&c->scan_tmp, sizeof(int32_t) * (c->nbins / 4096 + 2)));
  XB_CUDA(cudaMalloc(&c->red_partial, sizeof(double) * RED_BLOCKS * RED_MAXV));
  XB_CUDA(cudaMalloc(&c->red_out, sizeof(double) * RED_MAXV));
  XB_CUDA(cudaMallocHost(&c->red_host, sizeof(double) * RED_MAXV));
  if (g.nranks > 1) XB_CHECK(comm_init(c, uid));
  XB_CHECK(krylov_prepare(c));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}


extern "C" {

int xb_create(const xb_grid* gr, const void* uid, xb_ctx** out)
{
  if (!gr || !out) XB_FAIL("xb_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    XB_FAIL("xb_create: no CUDA device available (this library has no CPU fallback)");
  if (gr->device < 0 || gr->device >= ndev) XB_FAIL("xb_create: bad device ordinal");
  if (gr->nranks < 1 || gr->rank < 0 || gr->rank >= gr->nranks) XB_FAIL("xb_create: bad rank / nranks");
  for (int a = 0; a < 3; ++a)
    if (gr->n[a] < 1 || !(gr->d[a] > 0.0)) XB_FAIL("xb_create: bad geometry");
  XB_CUDA(cudaSetDevice(gr->device));
  xb_ctx* c = new xb_ctx();
  if (create_impl(c, gr, uid)) {  // any failure after `new` releases what was allocated so far (the message survives)
    xb_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

int xb_destroy(xb_ctx* c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  comm_free(c);
  for (auto& s : c->sorts) species_free(s);
  for (double* v : {c->E, c->B, c->B0, c->Ep, c->Ec, c->currI, c->currJe, c->rhs, c->tmp, c->tmp2, c->coef, c->stage, c->Z, c->cheb_r, c->cheb_d,
                    c->cheb_Md, c->ksp_u, c->red_partial, c->red_out, c->cap_x, c->cap_F, c->cap_g, c->cap_rhs0, c->cap_J})
    cudaFree(v);
  cudaFree(c->cap_counters);
  cudaFree(c->work_counter);
  cudaFree(c->removed_dev);
  for (auto e : c->nl.events)
    if (e) cudaEventDestroy(e);
  for (double* v : c->V) cudaFree(v);
  cudaFree(c->hist);
  cudaFree(c->cursor);
  cudaFree(c->scan_tmp);
  if (c->red_host) cudaFreeHost(c->red_host);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->ev3) cudaEventDestroy(c->ev3);
  for (auto& v : c->prof_events)
    for (auto e : v) cudaEventDestroy(e);
  if (c->copy_done) cudaEventDestroy(c->copy_done);
  if (c->b_ready) cudaEventDestroy(c->b_ready);
  if (c->host_stream) cudaStreamDestroy(c->host_stream);
  if (c->blocks_ready) cudaEventDestroy(c->blocks_ready);
  if (c->blocks_here) cudaEventDestroy(c->blocks_here);
  if (c->halo_ready) cudaEventDestroy(c->halo_ready);
  if (c->halo_done) cudaEventDestroy(c->halo_done);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int xb_species_add(xb_ctx* c, double q, double m, double n, int32_t Np, int64_t capacity, int32_t* sid)
{
  XB_API_BEGIN(c);
  if (capacity < 1 || capacity >= (int64_t)0x7fffffff) XB_FAIL("xb_species_add: capacity must be in [1, 2^31)");
  if (Np < 1 || !(m > 0.0)) XB_FAIL("xb_species_add: bad parameters");
  c->sorts.emplace_back();
  Species& s = c->sorts.back();
  s.q = q; s.m = m; s.n = n; s.Np = Np;
  if (species_alloc(c, s, capacity)) return 1;
  if (sid) *sid = (int32_t)c->sorts.size() - 1;
  return 0;
}

int xb_particles_append(xb_ctx* c, int32_t sid, const double* aos6, const uint64_t* ids, int64_t count, int64_t* added)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  Species& s = c->sorts[sid];
  const Grid& g = c->g;
  std::vector<double> soa[6];
  std::vector<uint64_t> kid;
  for (auto& v : soa) v.reserve(count);
  for (int64_t p = 0; p < count; ++p) {
    const double* pt = aos6 + 6 * p;
    // interfaces/particles.cpp:49-56: FLOOR_STEP(r, d) - start must lie inside the local box
    const int vx = (int)std::floor(pt[0] / g.dx), vy = (int)std::floor(pt[1] / g.dy), vz = (int)std::floor(pt[2] / g.dz) - g.z0;
    const uint64_t this_id = ids ? ids[p] : s.next_id + (uint64_t)p;
    if (vx < 0 || vx >= g.nx || vy < 0 || vy >= g.ny || vz < 0 || vz >= g.nzl) continue;
    for (int k = 0; k < 6; ++k) soa[k].push_back(pt[k]);
    kid.push_back(this_id);
  }
  if (!ids) s.next_id += (uint64_t)count;
  const int64_t nadd = (int64_t)kid.size();
  if (s.count + nadd > s.capacity) XB_FAIL("xb_particles_append: species capacity exceeded");
  for (int k = 0; k < 6; ++k)
    XB_CUDA(cudaMemcpy(s.p[s.cur][k] + s.count, soa[k].data(), sizeof(double) * nadd, cudaMemcpyHostToDevice));
  if (c->track_ids) XB_CUDA(cudaMemcpy(s.id[s.cur] + s.count, kid.data(), sizeof(uint64_t) * nadd, cudaMemcpyHostToDevice));
  s.count += nadd;
  s.sorted = false;
  if (added) *added = nadd;
  return 0;
}

int xb_particles_maxwellian(xb_ctx* c, int32_t sid, int64_t total, const double T[3], uint64_t seed, int32_t tov, int64_t* added)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  return particles_generate(c, c->sorts[sid], total, T, seed, tov, added);
}

int xb_particles_count(xb_ctx* c, int32_t sid, int64_t* count)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  *count = c->sorts[sid].count;
  return 0;
}

int xb_particles_download(xb_ctx* c, int32_t sid, double* aos6, uint64_t* ids, int64_t capacity, int64_t* count)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  Species& s = c->sorts[sid];
  XB_CHECK(ensure_sorted(c));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  if (capacity < s.count) XB_FAIL("xb_particles_download: buffer too small");
  std::vector<double> tmp(s.count);
  for (int k = 0; k < 6; ++k) {
    XB_CUDA(cudaMemcpy(tmp.data(), s.p[s.cur][k], sizeof(double) * s.count, cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < s.count; ++p) aos6[6 * p + k] = tmp[p];
  }
  if (ids) {
    if (!c->track_ids) XB_FAIL("xb_particles_download: ids requested but the context was created with track_ids = 0");
    XB_CUDA(cudaMemcpy(ids, s.id[s.cur], sizeof(uint64_t) * s.count, cudaMemcpyDeviceToHost));
  }
  if (count) *count = s.count;
  return 0;
}

int xb_field_upload(xb_ctx* c, int32_t which, int32_t sid, const double* host)
{
  XB_API_BEGIN(c);
  double* v = named_vector(c, which, sid);
  if (!v) XB_FAIL("xb_field_upload: unknown vector");
  XB_CHECK(upload_owned(c, host, v));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_field_download(xb_ctx* c, int32_t which, int32_t sid, double* host)
{
  XB_API_BEGIN(c);
  double* v = named_vector(c, which, sid);
  if (!v) XB_FAIL("xb_field_download: unknown vector");
  XB_CHECK(download_owned(c, v, host));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_solver_set(xb_ctx* c, int32_t which, double rtol, double atol, int32_t maxit, int32_t restart, int32_t precond)
{
  XB_API_BEGIN(c);
  if (which < 0 || which > 1) XB_FAIL("bad solver slot");
  if (restart < 1 || restart > RED_MAXV - 1) XB_FAIL("restart must be in [1, 31]");
  Solver& s = c->solver[which];
  s.rtol = rtol; s.atol = atol; s.maxit = maxit; s.restart = restart; s.precond = precond;
  return krylov_prepare(c);
}

int xb_solver_info(xb_ctx* c, int32_t which, int32_t* iterations, double* rnorm, int32_t* reason)
{
  XB_API_BEGIN(c);
  if (which < 0 || which > 1) XB_FAIL("bad solver slot");
  const Solver& s = c->solver[which];
  if (iterations) *iterations = s.iterations;
  if (rnorm) *rnorm = s.rnorm;
  if (reason) *reason = s.reason;
  return 0;
}

int xb_stage(xb_ctx* c, int32_t scheme, int32_t stage)
{
  XB_API_BEGIN(c);
  if (scheme != XB_ECSIM && scheme != XB_ECSIMCORR && scheme != XB_ECCAPFIM) XB_FAIL("unknown scheme");
  if (stage == XB_STAGE_CLEAR_SOURCES) XB_CHECK(ensure_sorted(c));
  return run_stage(c, scheme, stage);
}

int xb_step(xb_ctx* c, int32_t scheme)
{
  XB_API_BEGIN(c);
  if (scheme != XB_ECSIM && scheme != XB_ECSIMCORR && scheme != XB_ECCAPFIM) XB_FAIL("unknown scheme");
  if (c->g.open_z && scheme != XB_ECSIM) XB_FAIL("an open z boundary is covered for scheme ecsim only");
  XB_CHECK(ensure_sorted(c));
  for (int st = 0; st < XB_STAGE_COUNT; ++st) XB_CHECK(run_stage(c, scheme, st));
  return 0;
}

int xb_step_host(xb_ctx* c, int32_t scheme, double* E, double* B, const double* B0, double* kinetic)
{
  XB_API_BEGIN(c);
  if (scheme != XB_ECSIM && scheme != XB_ECSIMCORR && scheme != XB_ECCAPFIM) XB_FAIL("unknown scheme");
  const Grid& g = c->g;
  // B^n is needed from the moment deposition on (after the re-binning), E^n and B0 only by the field solve: for
  // ecsim / ecsimcorr the uploads travel on their own stream underneath the push, re-binning and moment kernels
  const bool overlap = scheme != XB_ECCAPFIM;
  if (overlap) {
    XB_CUDA(cudaEventRecord(c->copy_done, c->stream));  // the previous step's readers of E, B, B0 are done
    XB_CUDA(cudaStreamWaitEvent(c->host_stream, c->copy_done, 0));
    XB_CUDA(cudaMemcpyAsync(c->B + g.own0, B, sizeof(double) * g.nown, cudaMemcpyHostToDevice, c->host_stream));
    XB_CUDA(cudaEventRecord(c->b_ready, c->host_stream));
    c->b_pending = true;
    XB_CUDA(cudaMemcpyAsync(c->E + g.own0, E, sizeof(double) * g.nown, cudaMemcpyHostToDevice, c->host_stream));
    if (B0) XB_CUDA(cudaMemcpyAsync(c->B0 + g.own0, B0, sizeof(double) * g.nown, cudaMemcpyHostToDevice, c->host_stream));
    XB_CUDA(cudaEventRecord(c->copy_done, c->host_stream));
  }
  else {
    XB_CHECK(upload_owned(c, B, c->B));
    XB_CHECK(upload_owned(c, E, c->E));
    if (B0) XB_CHECK(upload_owned(c, B0, c->B0));
  }
  int rc = ensure_sorted(c);
  bool sent_early = false;
  for (int st = 0; st < XB_STAGE_COUNT && !rc; ++st) {
    if (overlap && st == XB_STAGE_ADVANCE_FIELDS) XB_CUDA(cudaStreamWaitEvent(c->stream, c->copy_done, 0));
    rc = run_stage(c, scheme, st);
    if (!rc && scheme == XB_ECSIM && st == XB_STAGE_ADVANCE_FIELDS) {
      // ecsim: E^{n+1} and B^{n+1} follow from E^{n+1/2} alone (ecsim/simulation.cpp:247-248).  They are formed in scratch
      // vectors right after the solve and go home underneath the second push, which still reads B^n; the final stage
      // forms them again in place (same kernel, same bits).
      rc = final_update_into(c, c->Ep, c->tmp, c->tmp2);
      if (!rc) {
        XB_CUDA(cudaEventRecord(c->copy_done, c->stream));
        XB_CUDA(cudaStreamWaitEvent(c->host_stream, c->copy_done, 0));
        XB_CUDA(cudaMemcpyAsync(E, c->tmp + g.own0, sizeof(double) * g.nown, cudaMemcpyDeviceToHost, c->host_stream));
        XB_CUDA(cudaMemcpyAsync(B, c->tmp2 + g.own0, sizeof(double) * g.nown, cudaMemcpyDeviceToHost, c->host_stream));
        sent_early = true;
      }
    }
  }
  if (c->b_pending) {  // an error before the deposition: nothing may be left waiting for the next call
    c->b_pending = false;
    cudaStreamSynchronize(c->host_stream);
  }
  if (rc) return rc;
  // (other schemes) E goes home on the second stream while the kinetic energies are reduced and B follows on the first
  if (!sent_early) {
    XB_CUDA(cudaEventRecord(c->copy_done, c->stream));
    XB_CUDA(cudaStreamWaitEvent(c->host_stream, c->copy_done, 0));
    XB_CUDA(cudaMemcpyAsync(E, c->E + g.own0, sizeof(double) * g.nown, cudaMemcpyDeviceToHost, c->host_stream));
  }
  if (kinetic)
    for (size_t i = 0; i < c->sorts.size(); ++i) XB_CHECK(kinetic_energy(c, c->sorts[i], nullptr, &kinetic[i]));
  if (!sent_early) XB_CHECK(download_owned(c, c->B, B));
  XB_CUDA(cudaStreamSynchronize(c->host_stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_run_steps(xb_ctx* c, int32_t scheme, int32_t k, double* ms)
{
  XB_API_BEGIN(c);
  XB_CHECK(ensure_sorted(c));
  XB_CUDA(cudaEventRecord(c->ev2, c->stream));
  for (int i = 0; i < k; ++i) XB_CHECK(xb_step(c, scheme));
  XB_CUDA(cudaEventRecord(c->ev3, c->stream));
  XB_CUDA(cudaEventSynchronize(c->ev3));
  float t = 0.f;
  XB_CUDA(cudaEventElapsedTime(&t, c->ev2, c->ev3));
  if (ms) *ms = t;
  return 0;
}

int xb_run_steps_host(xb_ctx* c, int32_t scheme, int32_t k, double* E, double* B, const double* B0, double* kinetic, double* ms)
{
  XB_API_BEGIN(c);
  XB_CHECK(ensure_sorted(c));
  XB_CUDA(cudaEventRecord(c->ev2, c->stream));
  for (int i = 0; i < k; ++i) XB_CHECK(xb_step_host(c, scheme, E, B, B0, kinetic));
  XB_CUDA(cudaEventRecord(c->ev3, c->stream));
  XB_CUDA(cudaEventSynchronize(c->ev3));
  float t = 0.f;
  XB_CUDA(cudaEventElapsedTime(&t, c->ev2, c->ev3));
  if (ms) *ms = t;
  return 0;
}

int xb_family_profile(xb_ctx* c, int32_t enable)
{
  XB_API_BEGIN(c);
  c->family_profile = enable != 0;
  for (auto& u : c->prof_used) u = 0;
  return 0;
}

int xb_family_profile_read(xb_ctx* c, int32_t family, int64_t* launches, double* total_ms)
{
  XB_API_BEGIN(c);
  if (family < 0 || family >= XB_FAMILY_COUNT) XB_FAIL("xb_family_profile_read: unknown family");
  XB_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  const auto& ev = c->prof_events[family];
  for (size_t i = 0; i + 1 < c->prof_used[family]; i += 2) {
    float t = 0.f;
    XB_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
    tot += t;
  }
  if (launches) *launches = (int64_t)(c->prof_used[family] / 2);
  if (total_ms) *total_ms = tot;
  return 0;
}

int xb_spmv_profile(xb_ctx* c, int32_t enable) { return xb_family_profile(c, enable); }
int xb_spmv_profile_read(xb_ctx* c, int64_t* launches, double* total_ms) { return xb_family_profile_read(c, XB_FAMILY_SPMV, launches, total_ms); }

int xb_field_sums(xb_ctx* c, int32_t which, int32_t sid, double out[4])
{
  XB_API_BEGIN(c);
  const double* v = named_vector(c, which, sid);
  if (!v || !out) XB_FAIL("xb_field_sums: unknown vector");
  return field_sums(c, v, out);
}

int xb_field_energy(xb_ctx* c, int32_t which, int32_t sid, double* out)
{
  XB_API_BEGIN(c);
  const double* v = named_vector(c, which, sid);
  if (!v || !out) XB_FAIL("xb_field_energy: unknown vector");
  const double* vs[1] = {v};
  double n2 = 0.0;
  XB_CHECK(dots(c, 1, vs, v, &n2));
  *out = 0.5 * n2;
  return 0;
}

int xb_scalar(xb_ctx* c, int32_t sid, int32_t which, double* out)
{
  XB_API_BEGIN(c);
  if (which == XB_J_DIFF_NORM) {
    *out = c->j_diff_norm;
    return 0;
  }
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  Species& s = c->sorts[sid];
  switch (which) {
    case XB_KINETIC: return kinetic_energy(c, s, nullptr, out);
    case XB_PRED_W: *out = s.pred_w; return 0;
    case XB_CORR_W: *out = s.corr_w; return 0;
    case XB_PRED_DK: *out = s.pred_dK; return 0;
    case XB_CORR_DK: *out = s.corr_dK; return 0;
    case XB_LAMBDA_DK: *out = s.lambda_dK; return 0;
    case XB_ENERGY_MEMBER: *out = s.energy; return 0;
  }
  XB_FAIL("unknown scalar");
}

int xb_particle_moments(xb_ctx* c, int32_t sid, double out[5])
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  return particle_moments(c, c->sorts[sid], out);
}

int xb_timing(xb_ctx* c, int32_t stage, double* seconds, int64_t* calls)
{
  XB_API_BEGIN(c);
  if (stage < 0 || stage >= XB_STAGE_COUNT) XB_FAIL("bad stage");
  if (seconds) *seconds = c->clock.seconds[stage];
  if (calls) *calls = c->clock.calls[stage];
  return 0;
}

int xb_timing_reset(xb_ctx* c)
{
  XB_API_BEGIN(c);
  c->clock = StageClock();
  return 0;
}

int xb_launch_count(xb_ctx* c, int64_t* launches)
{
  XB_API_BEGIN(c);
  *launches = c->launches;
  return 0;
}

int xb_spmv(xb_ctx* c, int32_t op, const double* x, double* y)
{
  XB_API_BEGIN(c);
  XB_CHECK(upload_owned(c, x, c->tmp2));
  XB_CHECK(spmv(c, op, c->tmp2, c->tmp));
  XB_CHECK(download_owned(c, c->tmp, y));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_spmv_bench(xb_ctx* c, int32_t op, int32_t reps, double* ms_per_spmv)
{
  XB_API_BEGIN(c);
  // deterministic pseudo-random x: reuse whatever is in E plus a ramp is not needed -- fill tmp2 on host
  std::vector<double> h(c->g.nown);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  for (auto& v : h) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    v = (double)(st >> 11) / 9007199254740992.0 - 0.5;
  }
  XB_CHECK(upload_owned(c, h.data(), c->tmp2));
  for (int i = 0; i < 3; ++i) XB_CHECK(spmv(c, op, c->tmp2, c->tmp));
  XB_CUDA(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < reps; ++i) XB_CHECK(spmv(c, op, c->tmp2, c->tmp));
  XB_CUDA(cudaEventRecord(c->ev1, c->stream));
  XB_CUDA(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  XB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  *ms_per_spmv = ms / reps;
  return 0;
}

int xb_operator_download(xb_ctx* c, double* coef)
{
  XB_API_BEGIN(c);
  double* plain = nullptr;
  XB_CUDA(cudaMalloc(&plain, sizeof(double) * NCOEF * c->g.ncl));
  int rc = coef_convert(c, plain, false);
  if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = 1;
  if (!rc && cudaMemcpy(coef, plain, sizeof(double) * NCOEF * c->g.ncl, cudaMemcpyDeviceToHost) != cudaSuccess) rc = 1;
  cudaFree(plain);
  if (rc) XB_FAIL("xb_operator_download failed");
  return 0;
}

int xb_operator_upload(xb_ctx* c, const double* coef)
{
  XB_API_BEGIN(c);
  double* plain = nullptr;
  XB_CUDA(cudaMalloc(&plain, sizeof(double) * NCOEF * c->g.ncl));
  int rc = cudaMemcpy(plain, coef, sizeof(double) * NCOEF * c->g.ncl, cudaMemcpyHostToDevice) != cudaSuccess;
  if (!rc) rc = coef_convert(c, plain, true);
  if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = 1;
  cudaFree(plain);
  if (rc) XB_FAIL("xb_operator_upload failed");
  c->coef_valid = true;
  return 0;
}

int xb_set_option(xb_ctx* c, int32_t what, int32_t value)
{
  XB_API_BEGIN(c);
  if (what == 0) {
    c->deposit_variant = value;
    return 0;
  }
  if (what == 1) {
    c->esirkepov_variant = value;
    return 0;
  }
  if (what == 2) {
    c->deterministic = value != 0;
    return 0;
  }
  if (what == 3) {
    c->cap_variant = value;
    return 0;
  }
  if (what == 4) {
    c->nl.warm_start = value != 0;
    return 0;
  }
  if (what == 5) {
    c->ws_backoff_ns = value < 0 ? 0 : value;
    return 0;
  }
  XB_FAIL("xb_set_option: unknown option");
}

int xb_nonlinear_set(xb_ctx* c, double atol, double rtol, double stol, int32_t maxit, int32_t depth, int32_t cheb_degree, double particle_tol,
                     int32_t particle_maxit)
{
  XB_API_BEGIN(c);
  if (maxit < 1 || depth < 1 || cheb_degree < 0 || particle_maxit < 0) XB_FAIL("xb_nonlinear_set: bad argument");
  Nonlinear& nl = c->nl;
  nl.atol = atol; nl.rtol = rtol; nl.stol = stol; nl.maxit = maxit; nl.depth = depth; nl.cheb_degree = cheb_degree;
  nl.cn_tol = particle_tol; nl.cn_maxit = particle_maxit;
  return 0;
}

int xb_nonlinear_info(xb_ctx* c, int32_t* iterations, int32_t* fevals, int32_t* reason, double* fnorm, double* avg_cn, double* avg_cells)
{
  XB_API_BEGIN(c);
  const Nonlinear& nl = c->nl;
  if (iterations) *iterations = nl.iterations;
  if (fevals) *fevals = nl.fevals;
  if (reason) *reason = nl.reason;
  if (fnorm) *fnorm = nl.fnorm;
  if (avg_cn) *avg_cn = nl.avg_cn;
  if (avg_cells) *avg_cells = nl.avg_cells;
  return 0;
}

int xb_nonlinear_history(xb_ctx* c, double* out, int32_t capacity, int32_t* length)
{
  XB_API_BEGIN(c);
  const Nonlinear& nl = c->nl;
  if (length) *length = (int32_t)nl.hist.size();
  for (int i = 0; out && i < capacity && i < (int)nl.hist.size(); ++i) out[i] = nl.hist[i];
  return 0;
}

int xb_nonlinear_profile(xb_ctx* c, int32_t enable, int64_t* evaluations, double* total_ms)
{
  XB_API_BEGIN(c);
  Nonlinear& nl = c->nl;
  if (evaluations) *evaluations = nl.push_evals;
  if (total_ms) *total_ms = nl.push_ms;
  if (enable >= 0) {
    nl.profile = enable != 0;
    nl.push_ms = 0.0;
    nl.push_evals = 0;
  }
  return 0;
}

int xb_eccapfim_function(xb_ctx* c, const double* x, double* f)
{
  XB_API_BEGIN(c);
  XB_CHECK(ensure_sorted(c));
  XB_CHECK(cap_prepare(c));
  XB_CHECK(upload_owned(c, x, c->cap_x));
  XB_CHECK(cap_form_function(c, c->cap_x, c->cap_F));
  XB_CHECK(cap_read_counters(c));
  XB_CHECK(download_owned(c, c->cap_F, f));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_charge_density(xb_ctx* c, int32_t sid, double* rho)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  Species& s = c->sorts[sid];
  XB_CHECK(charge_density(c, s));
  if (rho) {
    // component 0 of the owned part of the grid vector
    std::vector<double> tmp((size_t)c->g.nown);
    XB_CHECK(download_owned(c, s.rho[s.rho_cur], tmp.data()));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < c->g.ncl; ++i) rho[i] = tmp[3 * i];
  }
  return 0;
}

int xb_distribution_moment_region(xb_ctx* c, int32_t sid, int32_t moment, const int32_t start[3], const int32_t size[3], double* out)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  if (!out) XB_FAIL("xb_distribution_moment: null output");
  XB_CHECK(distribution_moment(c, c->sorts[sid], moment, start, size));
  const int ms = moment_size(moment);
  std::vector<double> lo((size_t)c->g.nown), hi;
  XB_CHECK(download_owned(c, c->tmp2, lo.data()));
  if (ms > 3) {
    hi.resize((size_t)c->g.nown);
    XB_CHECK(download_owned(c, c->tmp, hi.data()));
  }
  XB_CUDA(cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < c->g.ncl; ++i)
    for (int j = 0; j < ms; ++j) out[(size_t)i * ms + j] = j < 3 ? lo[3 * i + j] : hi[3 * i + j - 3];
  return 0;
}

int xb_distribution_moment(xb_ctx* c, int32_t sid, int32_t moment, double* out) { return xb_distribution_moment_region(c, sid, moment, nullptr, nullptr, out); }

int xb_velocity_distribution_size(const double dv[2], const double vmin[2], const double vmax[2], int32_t* start, int32_t* size)
{
  if (!dv || !vmin || !vmax || !start || !size) XB_FAIL("xb_velocity_distribution_size: null argument");
  if (!(dv[0] > 0.0) || !(dv[1] > 0.0)) XB_FAIL("xb_velocity_distribution_size: dv must be positive");
  velocity_region(dv, vmin, vmax, start, size);
  return 0;
}

int xb_velocity_distribution(xb_ctx* c, int32_t sid, int32_t projector, int32_t geometry, const double p[6], const double dv[2], const double vmin[2],
                             const double vmax[2], double* out)
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  if (projector < XB_PROJECTOR_VX_VY || projector > XB_PROJECTOR_VR_VPHI) XB_FAIL("xb_velocity_distribution: unknown projector");
  if ((geometry != XB_GEOMETRY_BOX && geometry != XB_GEOMETRY_CYLINDER) || !p) XB_FAIL("xb_velocity_distribution: unknown geometry");
  if (!dv || !vmin || !vmax || !out) XB_FAIL("xb_velocity_distribution: null argument");
  if (!(dv[0] > 0.0) || !(dv[1] > 0.0)) XB_FAIL("xb_velocity_distribution: dv must be positive");
  Geometry ge;
  ge.kind = geometry;
  for (int k = 0; k < 6; ++k) ge.p[k] = p[k];
  return velocity_distribution(c, c->sorts[sid], projector, ge, dv, vmin, vmax, out);
}

int xb_momentum(xb_ctx* c, int32_t sid, double out[6])
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  if (!out) XB_FAIL("xb_momentum: null output");
  return momentum(c, c->sorts[sid], out);
}

int xb_charge_conservation(xb_ctx* c, int32_t which_current, double* norms)
{
  XB_API_BEGIN(c);
  if (which_current != 0 && which_current != 1) XB_FAIL("xb_charge_conservation: which_current must be 0 (currJe) or 1 (J)");
  if (!norms) XB_FAIL("xb_charge_conservation: null output");
  return charge_conservation(c, which_current, norms);
}

int xb_fields_damping(xb_ctx* c, int32_t geometry, const double p[6], double coefficient, double* damped_energy)
{
  XB_API_BEGIN(c);
  if ((geometry != XB_GEOMETRY_BOX && geometry != XB_GEOMETRY_CYLINDER) || !p) XB_FAIL("xb_fields_damping: unknown geometry");
  Geometry ge;
  ge.kind = geometry;
  for (int k = 0; k < 6; ++k) ge.p[k] = p[k];
  return fields_damping(c, ge, coefficient, damped_energy);
}

int xb_particles_remove(xb_ctx* c, int32_t sid, int32_t geometry, const double p[6], double out[2])
{
  XB_API_BEGIN(c);
  if (sid < 0 || sid >= (int)c->sorts.size()) XB_FAIL("bad species id");
  if ((geometry != XB_GEOMETRY_BOX && geometry != XB_GEOMETRY_CYLINDER) || !p) XB_FAIL("xb_particles_remove: unknown geometry");
  // the removal rides on a re-binning without a move: the key pass tests the cell of every particle, the scatter drops the marked ones
  if (!c->removed_dev) XB_CUDA(cudaMalloc(&c->removed_dev, 2 * sizeof(unsigned long long)));
  XB_CUDA(cudaMemsetAsync(c->removed_dev, 0, 2 * sizeof(unsigned long long), c->stream));
  c->remove.kind = geometry;
  for (int k = 0; k < 6; ++k) c->remove.p[k] = p[k];
  const int rc = sort_species(c, c->sorts[sid], 0.0);
  c->remove.kind = -1;
  if (rc) return rc;
  unsigned long long h[2];
  XB_CUDA(cudaMemcpyAsync(h, c->removed_dev, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  double res[2];
  res[0] = (double)h[0];
  memcpy(&res[1], &h[1], sizeof(double));
  if (c->g.nranks > 1) {
    c->red_host[0] = res[0];
    c->red_host[1] = res[1];
    XB_CUDA(cudaMemcpyAsync(c->red_out, c->red_host, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    XB_CHECK(comm_allreduce_sum(c, c->red_out, 2));
    XB_CUDA(cudaMemcpyAsync(c->red_host, c->red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    res[0] = c->red_host[0];
    res[1] = c->red_host[1];
  }
  if (out) {
    out[0] = res[0];
    out[1] = res[1];
  }
  return 0;
}

int xb_deposit(xb_ctx* c)
{
  XB_API_BEGIN(c);
  XB_CHECK(ensure_sorted(c));
  XB_CHECK(stage_clear(c, XB_ECSIM));
  XB_CHECK(deposit_moments(c));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_solve(xb_ctx* c, int32_t which, int32_t op, const double* b, double* x)
{
  XB_API_BEGIN(c);
  if (which < 0 || which > 1) XB_FAIL("bad solver slot");
  XB_CHECK(upload_owned(c, b, c->rhs));
  XB_CHECK(gmres(c, which, op, c->rhs, c->tmp2));
  XB_CHECK(download_owned(c, c->tmp2, x));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_curl(xb_ctx* c, int32_t positive, const double* f, double* out)
{
  XB_API_BEGIN(c);
  XB_CHECK(upload_owned(c, f, c->tmp2));
  XB_CHECK(halo_fill(c, c->tmp2, 1));
  XB_CHECK(curl_apply(c, positive != 0, c->tmp2, c->tmp, 1.0, false));
  XB_CHECK(download_owned(c, c->tmp, out));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int xb_kernel_bench(xb_ctx* c, int32_t what, int32_t reps, double* ms)
{
  XB_API_BEGIN(c);
  XB_CHECK(ensure_sorted(c));
  auto run = [&](int) -> int {
    switch (what) {
      case 0:
        for (auto& s : c->sorts) XB_CHECK(sort_species(c, s, 0.0));
        return 0;
      case 1: return deposit_moments(c);
      case 2:
        XB_CHECK(halo_fill(c, c->Ep, 1));
        XB_CHECK(halo_fill(c, c->B, 1));
        for (auto& s : c->sorts) XB_CHECK(push_second(c, s, c->Ep, c->B));
        return 0;
      case 3: XB_CHECK(build_rhs(c, c->currI, c->rhs)); return gmres(c, XB_SOLVER_PREDICT, XB_OP_A, c->rhs, c->tmp2);
    }
    xb::set_error("xb_kernel_bench: unknown selector");
    return 1;
  };
  XB_CHECK(run(0));
  XB_CUDA(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < reps; ++i) XB_CHECK(run(i));
  XB_CUDA(cudaEventRecord(c->ev1, c->stream));
  XB_CUDA(cudaEventSynchronize(c->ev1));
  float t = 0.f;
  XB_CUDA(cudaEventElapsedTime(&t, c->ev0, c->ev1));
  *ms = t / reps;
  return 0;
}

}  // extern "C"
