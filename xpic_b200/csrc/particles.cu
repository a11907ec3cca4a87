// particles.cu -- cell-sorted SoA particle store: move + periodic wrap + re-binning (counting
// sort by (cell, half-cell octant)), field gather + Boris velocity update, kinetic energy.
//
// Reference code replaced:
//   BorisPush::update_r            src/algorithms/boris_push.cpp:19-22   (fused into the key pass)
//   Particles::update_cells_seq    src/interfaces/particles.cpp:79-116   (wrap + re-bin)
//   g_bound_periodic               src/interfaces/point.cpp:18-26
//   interpolate_E_s1 / B_s1        src/impls/ecsim/simulation.cpp:8-118
//   BorisPush::update_vEB          src/algorithms/boris_push.cpp:48-57
//   ecsim::Particles::second_push  src/impls/ecsim/particles.cpp:175-192
//   Energy::calculate_kinetic      src/diagnostics/energy.cpp:61-107
// All kernels are HBM-bound streams over the SoA arrays (72 B / particle for a push).
#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"

namespace xb {

static inline int grid_for(int64_t n, int threads = 256)
{
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 0x7fffffff) b = 0x7fffffff;
  return (int)b;
}

constexpr int64_t PARTICLE_SKEW = 33 * 32;  // doubles: 33 x 256 bytes

int species_alloc(xb_ctx* c, Species& s, int64_t capacity)
{
  s.capacity = capacity;
  for (int b = 0; b < 2; ++b) {
    for (int k = 0; k < 6; ++k) {
      // The kernels stream the same index of up to thirteen arrays at once.  cudaMalloc hands out bases that are equal
      // modulo 2 MB: every array gets its own offset (an odd multiple of 256 B per array) so that the streams do not
      // walk the DRAM channels in lockstep.
      const int64_t skew = (int64_t)(b * 6 + k) * PARTICLE_SKEW;
      XB_CUDA(cudaMalloc(&s.p_alloc[b][k], sizeof(double) * (capacity + 12 * PARTICLE_SKEW)));
      s.p[b][k] = s.p_alloc[b][k] + skew;
    }
    if (c->track_ids) XB_CUDA(cudaMalloc(&s.id[b], sizeof(uint64_t) * capacity));
  }
  XB_CUDA(cudaMalloc(&s.key, sizeof(int32_t) * capacity));
  XB_CUDA(cudaMalloc(&s.bin_start, sizeof(int32_t) * (c->nbins + 1)));
  XB_CUDA(cudaMemset(s.bin_start, 0, sizeof(int32_t) * (c->nbins + 1)));
  XB_CUDA(cudaMalloc(&s.currI, sizeof(double) * c->g.ntot));
  XB_CUDA(cudaMalloc(&s.currJe, sizeof(double) * c->g.ntot));
  XB_CUDA(cudaMemset(s.currI, 0, sizeof(double) * c->g.ntot));
  XB_CUDA(cudaMemset(s.currJe, 0, sizeof(double) * c->g.ntot));
  return 0;
}

void species_free(Species& s)
{
  for (int b = 0; b < 2; ++b) {
    for (int k = 0; k < 6; ++k) cudaFree(s.p_alloc[b][k]);
    cudaFree(s.id[b]);
  }
  cudaFree(s.key);
  cudaFree(s.rec);
  cudaFree(s.bin_start);
  cudaFree(s.currI);
  cudaFree(s.currJe);
  cudaFree(s.rho[0]);
  cudaFree(s.rho[1]);
  migrate_free(s);
}

// ---------------------------------------------------------------------------------------------
// pass 1: r += v * dtm, periodic wrap (point.cpp:18-26), bin key, histogram
// ---------------------------------------------------------------------------------------------
// single-rank key pass: the moved position is NOT stored -- the scatter pass recomputes it with the
// same moved_coord() while it copies the particle (saves 24 B / particle of HBM writes)
__global__ void k_move_key(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                           const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz, double dtm,
                           int32_t* __restrict__ key, int32_t* __restrict__ hist, Geometry rm, double m_mpw, unsigned long long* __restrict__ tally)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = moved_coord(x[i], dtm != 0.0 ? vx[i] : 0.0, dtm, g.Lx);
  const double py = moved_coord(y[i], dtm != 0.0 ? vy[i] : 0.0, dtm, g.Ly);
  const double pz = moved_z(g, z[i], dtm != 0.0 ? vz[i] : 0.0, dtm);
  if (left_the_box(g, pz) || removed_by_command(g, rm, px, py, pz, vx[i], vy[i], vz[i], m_mpw, tally)) {
    key[i] = -1;  // not scattered: the particle is gone
    return;
  }
  const int32_t k = particle_key(g, px, py, pz, slab_plane(g, pz));
  key[i] = k;
  atomicAdd(&hist[k], 1);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of the histogram (int32), three small kernels
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                                            int32_t* __restrict__ tile_sums)
{
  __shared__ int32_t warp_tot[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int32_t w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int32_t run = incl - sum + (wid > 0 ? warp_tot[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

__global__ void k_scan_sums(int32_t* __restrict__ tile_sums, int ntiles, int32_t* __restrict__ total_out)
{
  // one block; sequential over chunks of blockDim
  __shared__ int32_t carry;
  __shared__ int32_t wt[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int32_t v = i < ntiles ? tile_sums[i] : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wt[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int32_t w = lane < (blockDim.x >> 5) ? wt[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wt[lane] = w;
    }
    __syncthreads();
    const int32_t excl = carry + incl - v + (wid > 0 ? wt[wid - 1] : 0);
    if (i < ntiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void k_scan_add(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ tile_sums, int32_t* __restrict__ cursor)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t v = out[i] + tile_sums[i / SCAN_TILE];
  out[i] = v;
  cursor[i] = v;
}

// pass 2: scatter into the other SoA buffer.  The input is nearly sorted, so the lanes of a warp
// mostly share a handful of bins: lanes with equal keys elect a leader that reserves the whole
// run with one atomic.  The kernel is bound by the latency of those returning atomics (ncu: 65 stall
// cycles per issued instruction on the long scoreboard), so every thread carries SCATTER_ITEMS
// particles whose reservations are all in flight before the first one is used.
constexpr int SCATTER_ITEMS = 4;
constexpr int SCATTER_THREADS = 256;

__global__ void __launch_bounds__(SCATTER_THREADS) k_scatter(int64_t n, const int32_t* __restrict__ key, int32_t* __restrict__ cursor,
                                                            const double* __restrict__ s0, const double* __restrict__ s1,
                                                            const double* __restrict__ s2, const double* __restrict__ s3,
                                                            const double* __restrict__ s4, const double* __restrict__ s5,
                                                            const uint64_t* __restrict__ sid, double* __restrict__ d0, double* __restrict__ d1,
                                                            double* __restrict__ d2, double* __restrict__ d3, double* __restrict__ d4,
                                                            double* __restrict__ d5, uint64_t* __restrict__ did, double dtm, Grid g)
{
  const int lane = threadIdx.x & 31;
  const int64_t first = (int64_t)blockIdx.x * (SCATTER_THREADS * SCATTER_ITEMS) + threadIdx.x;
  int32_t k[SCATTER_ITEMS], base[SCATTER_ITEMS];
  int head_lane[SCATTER_ITEMS];
#pragma unroll
  for (int j = 0; j < SCATTER_ITEMS; ++j) {
    const int64_t i = first + (int64_t)j * SCATTER_THREADS;
    k[j] = i < n ? key[i] : -1;  // key -1: the particle left the slab (migrate.cu) or the box (open boundary)
  }
  // The input is nearly sorted: the lanes of a warp form a few runs of equal keys.  The first lane of a run reserves
  // the whole run with one atomic; runs are found with one shuffle and two ballots (a key that appears in two
  // separate runs simply reserves twice).
#pragma unroll
  for (int j = 0; j < SCATTER_ITEMS; ++j) {
    const bool live = k[j] >= 0;
    const int32_t prev = __shfl_up_sync(0xffffffffu, k[j], 1);
    const bool head = live && (lane == 0 || prev != k[j]);
    const unsigned heads = __ballot_sync(0xffffffffu, head), lives = __ballot_sync(0xffffffffu, live);
    const unsigned upto = 0xffffffffu >> (31 - lane);  // bits 0 .. lane
    head_lane[j] = 31 - __clz(heads & upto);
    const unsigned stops = (heads | ~lives) & ~upto;   // the next run start or dead lane above this lane
    const int end = stops ? __ffs(stops) - 1 : 32;
    base[j] = 0;
    if (head) base[j] = atomicAdd(&cursor[k[j]], end - lane);
  }
  // the particles themselves are fetched while the reservations are in flight
  double r[SCATTER_ITEMS][6];
  uint64_t pid[SCATTER_ITEMS];
#pragma unroll
  for (int j = 0; j < SCATTER_ITEMS; ++j) {
    const int64_t i = first + (int64_t)j * SCATTER_THREADS;
    if (k[j] >= 0) {
      r[j][0] = s0[i];
      r[j][1] = s1[i];
      r[j][2] = s2[i];
      r[j][3] = s3[i];
      r[j][4] = s4[i];
      r[j][5] = s5[i];
      pid[j] = sid ? sid[i] : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < SCATTER_ITEMS; ++j) {
    const int32_t b = __shfl_sync(0xffffffffu, base[j], head_lane[j] & 31);
    if (k[j] < 0) continue;
    const int32_t pos = b + (lane - head_lane[j]);
    d0[pos] = moved_coord(r[j][0], r[j][3], dtm, g.Lx);  // the same bits the key pass binned (dtm = 0: plain copy + wrap)
    d1[pos] = moved_coord(r[j][1], r[j][4], dtm, g.Ly);
    d2[pos] = moved_z(g, r[j][2], r[j][5], dtm);
    d3[pos] = r[j][3];
    d4[pos] = r[j][4];
    d5[pos] = r[j][5];
    if (sid) did[pos] = pid[j];
  }
}

// canonical order inside a bin (makes every later sum reproducible run to run although the scatter
// reserves slots with atomics): ascending particle id when ids are tracked, else ascending x
// coordinate (ties would need bit-identical x inside one half cell).  One thread per bin, insertion
// sort (bins hold ~ppc/8 particles).
template <bool BY_ID>
__global__ void k_order_bins(int64_t nbins, const int32_t* __restrict__ bin_start, double* __restrict__ p0, double* __restrict__ p1,
                             double* __restrict__ p2, double* __restrict__ p3, double* __restrict__ p4, double* __restrict__ p5,
                             uint64_t* __restrict__ id)
{
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int32_t lo = bin_start[b], hi = bin_start[b + 1];
  for (int32_t i = lo + 1; i < hi; ++i) {
    const uint64_t kid = BY_ID ? id[i] : 0;
    const double a0 = p0[i];
    int32_t j = i - 1;
    if (BY_ID ? (id[j] <= kid) : (p0[j] <= a0)) continue;
    const double a1 = p1[i], a2 = p2[i], a3 = p3[i], a4 = p4[i], a5 = p5[i];
    while (j >= lo && (BY_ID ? (id[j] > kid) : (p0[j] > a0))) {
      if (BY_ID) id[j + 1] = id[j];
      p0[j + 1] = p0[j];
      p1[j + 1] = p1[j];
      p2[j + 1] = p2[j];
      p3[j + 1] = p3[j];
      p4[j + 1] = p4[j];
      p5[j + 1] = p5[j];
      --j;
    }
    if (BY_ID) id[j + 1] = kid;
    p0[j + 1] = a0;
    p1[j + 1] = a1;
    p2[j + 1] = a2;
    p3[j + 1] = a3;
    p4[j + 1] = a4;
    p5[j + 1] = a5;
  }
}

// scan the histogram, scatter the local particles (and, in multi-rank runs, the arrivals) into the
// other SoA buffer, canonicalise the order inside every bin when ids are tracked
int sort_scan_and_scatter(xb_ctx* c, Species& s, int64_t nlocal, const MigrateBuffers* arr, int64_t n_from_down, int64_t n_from_up, double dt_move)
{
  const Grid& g = c->g;
  const int ntiles = (int)((c->nbins + SCAN_TILE - 1) / SCAN_TILE);
  XB_LAUNCH(c, k_scan_tiles, ntiles, SCAN_THREADS, 0, c->hist, s.bin_start, c->nbins, c->scan_tmp);
  XB_LAUNCH(c, k_scan_sums, 1, 1024, 0, c->scan_tmp, ntiles, s.bin_start + c->nbins);
  XB_LAUNCH(c, k_scan_add, grid_for(c->nbins), 256, 0, s.bin_start, c->nbins, c->scan_tmp, c->cursor);
  double** p = s.p[s.cur];
  double** d = s.p[1 - s.cur];
  uint64_t* did = s.id[1 - s.cur];
  if (nlocal > 0)
    XB_LAUNCH(c, k_scatter, grid_for(nlocal, SCATTER_THREADS * SCATTER_ITEMS), SCATTER_THREADS, 0, nlocal, s.key, c->cursor, p[0], p[1], p[2], p[3], p[4], p[5], s.id[s.cur], d[0], d[1], d[2], d[3],
              d[4], d[5], did, dt_move, g);
  if (arr) {
    const int64_t na[2] = {n_from_down, n_from_up};
    for (int k = 0; k < 2; ++k)
      if (na[k] > 0)
        XB_LAUNCH(c, k_scatter, grid_for(na[k], SCATTER_THREADS * SCATTER_ITEMS), SCATTER_THREADS, 0, na[k], arr->recv_key[k], c->cursor, arr->recv[k][0], arr->recv[k][1], arr->recv[k][2],
                  arr->recv[k][3], arr->recv[k][4], arr->recv[k][5], c->track_ids ? reinterpret_cast<const uint64_t*>(arr->recv[k][6]) : nullptr, d[0],
                  d[1], d[2], d[3], d[4], d[5], did, 0.0, g);
  }
  s.cur = 1 - s.cur;
  {
    double** q = s.p[s.cur];
    if (c->track_ids)
      XB_LAUNCH(c, k_order_bins<true>, grid_for(c->nbins, 128), 128, 0, c->nbins, s.bin_start, q[0], q[1], q[2], q[3], q[4], q[5], s.id[s.cur]);
    else if (c->deterministic)
      XB_LAUNCH(c, k_order_bins<false>, grid_for(c->nbins, 128), 128, 0, c->nbins, s.bin_start, q[0], q[1], q[2], q[3], q[4], q[5], nullptr);
  }
  return 0;
}

// open z boundary: particles may have left the box; the scan's total is the number that stayed
int count_after_open_sort(xb_ctx* c, Species& s)
{
  int32_t total = 0;
  XB_CUDA(cudaMemcpyAsync(&total, s.bin_start + c->nbins, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  s.count = total;
  return 0;
}

int particles_sort(xb_ctx* c, Species& s, double dt_move)
{
  const Grid& g = c->g;
  const int64_t n = s.count;
  if (g.nranks > 1) XB_FAIL("particles_sort: multi-rank runs sort through migrate_and_sort");
  XB_CUDA(cudaMemsetAsync(c->hist, 0, sizeof(int32_t) * c->nbins, c->stream));
  double** p = s.p[s.cur];
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT_KEYS));
  if (n > 0)
    XB_LAUNCH(c, k_move_key, grid_for(n), 256, 0, g, n, p[0], p[1], p[2], p[3], p[4], p[5], dt_move, s.key, c->hist, c->remove, s.m * (s.n / (double)s.Np),
              c->removed_dev);
  XB_CHECK(prof_end(c, XB_FAMILY_SORT_KEYS));
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT_SCATTER));
  XB_CHECK(sort_scan_and_scatter(c, s, n, nullptr, 0, 0, dt_move));
  XB_CHECK(prof_end(c, XB_FAMILY_SORT_SCATTER));
  if (g.open_z || c->remove.kind >= 0) XB_CHECK(count_after_open_sort(c, s));
  s.sorted = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// synthetic initial condition for the large configurations: CoordinateInBox + MaxwellianMomentum
// (src/utils/particles_load.cpp:11-18,52-76) driven by a counter-based generator instead of the
// serial mt19937 stream.  Every rank walks the same global stream and keeps the
// particles of its slab, so the initial state does not depend on the decomposition
// (src/interfaces/particles.cpp:47-57 semantics).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ double u01(uint64_t seed, uint64_t i, int k)
{
  const uint64_t h = splitmix64(seed ^ splitmix64(i * 16 + (uint64_t)k));
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0, 1)
}

__global__ void k_generate_maxwellian(Grid g, uint64_t total, uint64_t seed, double sx, double sy, double sz, double m, int tov,
                                      double* __restrict__ x, double* __restrict__ y, double* __restrict__ z, double* __restrict__ vx,
                                      double* __restrict__ vy, double* __restrict__ vz, uint64_t* __restrict__ id, uint64_t id0,
                                      unsigned long long* __restrict__ cursor, int64_t base, int64_t capacity)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const double pz = u01(seed, i, 2) * g.Lz;
    const int cz = (int)floor(pz / g.dz) - g.z0;
    if (cz < 0 || cz >= g.nzl) continue;
    const double px = u01(seed, i, 0) * g.Lx, py = u01(seed, i, 1) * g.Ly;
    double p[3];
    const double sg[3] = {sx, sy, sz};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      p[c] = sinpi(2.0 * u01(seed, i, 3 + 2 * c)) * sqrt(-2.0 * sg[c] * log(u01(seed, i, 4 + 2 * c)));
    if (tov) {
      const double den = sqrt(m * m + (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
      p[0] /= den;
      p[1] /= den;
      p[2] /= den;
    }
    const int64_t pos = base + (int64_t)atomicAdd(cursor, 1ull);
    if (pos >= capacity) continue;  // counted, reported as an error by the host
    x[pos] = px;
    y[pos] = py;
    z[pos] = pz;
    vx[pos] = p[0];
    vy[pos] = p[1];
    vz[pos] = p[2];
    if (id) id[pos] = id0 + i;
  }
}

int particles_generate(xb_ctx* c, Species& s, int64_t total, const double* T, uint64_t seed, int tov, int64_t* added)
{
  unsigned long long* cur = reinterpret_cast<unsigned long long*>(c->red_out);
  XB_CUDA(cudaMemsetAsync(cur, 0, sizeof(unsigned long long), c->stream));
  double** p = s.p[s.cur];
  const double f = s.m / 511.0;  // temperature_momentum: sqrt(-2 (T m / mec2) ln u)
  XB_LAUNCH(c, k_generate_maxwellian, 148 * 32, 256, 0, c->g, (uint64_t)total, seed, T[0] * f, T[1] * f, T[2] * f, s.m, tov, p[0], p[1], p[2], p[3],
            p[4], p[5], s.id[s.cur], s.next_id, cur, s.count, s.capacity);
  unsigned long long n = 0;
  XB_CUDA(cudaMemcpyAsync(&n, cur, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  if (s.count + (int64_t)n > s.capacity) XB_FAIL("particles_generate: species capacity exceeded");
  s.count += (int64_t)n;
  s.next_id += (uint64_t)total;
  s.sorted = false;
  if (added) *added = (int64_t)n;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// second push (ecsim): gather E^{n+1/2}, B^n at the particle, Boris update of v
// ---------------------------------------------------------------------------------------------
// One CTA per group of 64 x-consecutive cells: the 66 x 3 x 3 nodes x 3 components of E and of B
// its particles can touch are staged in shared memory once, so the 48 gathers per particle are
// shared-memory reads and HBM sees only the particle stream (72 B / particle).
constexpr int PUSH_THREADS = 256;
constexpr int PUSH_CELLS = 64;

// WORK: also accumulate the predicted field work  q n/Np * (v_old + v_new)/2 . E_p  of ecsimcorr
// (src/impls/ecsimcorr/particles.cpp:77-78), one partial per CTA (summed by a fixed tree afterwards)
template <bool WORK>
__global__ void __launch_bounds__(PUSH_THREADS) k_push_second(Grid g, const int32_t* __restrict__ bin_start, const double* __restrict__ x,
                                                              const double* __restrict__ y, const double* __restrict__ z, double* __restrict__ vx,
                                                              double* __restrict__ vy, double* __restrict__ vz, const double* __restrict__ E,
                                                              const double* __restrict__ B, double qm, int groups_x, double qn_Np,
                                                              double* __restrict__ partial)
{
  __shared__ double Et[FieldTile<PUSH_CELLS>::SIZE], Bt[FieldTile<PUSH_CELLS>::SIZE];
  const int gx = blockIdx.x % groups_x, row = blockIdx.x / groups_x;  // row = zl * ny + cy
  const int cy = row % g.ny, zl = row / g.ny;
  const int cx0 = gx * PUSH_CELLS, ncell = min(PUSH_CELLS, g.nx - cx0);
  load_field_tile<PUSH_CELLS>(g, E, cx0, cy, zl, Et, threadIdx.x, PUSH_THREADS);
  load_field_tile<PUSH_CELLS>(g, B, cx0, cy, zl, Bt, threadIdx.x, PUSH_THREADS);
  const int64_t cell0 = ((int64_t)(zl + 1) * g.ny + cy) * g.nx + cx0;  // bin plane = zl + 1
  const int32_t p0 = bin_start[cell0 << 3], p1 = bin_start[(cell0 + ncell) << 3];
  __syncthreads();
  double work = 0.0;
  // two particles per thread and round: their twelve loads are in flight together (the kernel is bound by the latency
  // of the particle stream, not by arithmetic)
  for (int32_t i0 = p0 + threadIdx.x; i0 < p1; i0 += 2 * PUSH_THREADS) {
    const int32_t idx[2] = {i0, i0 + PUSH_THREADS};
    double r[2][6];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (idx[u] < p1) {
        r[u][0] = x[idx[u]];
        r[u][1] = y[idx[u]];
        r[u][2] = z[idx[u]];
        r[u][3] = vx[idx[u]];
        r[u][4] = vy[idx[u]];
        r[u][5] = vz[idx[u]];
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (idx[u] >= p1) continue;
      const int32_t i = idx[u];
      Weights w;
      make_weights(g, r[u][0], r[u][1], r[u][2], 0, w);
      const TileIndex t = tile_index<PUSH_CELLS>(w, cx0, cy, zl);
      double Ep[3], Bp[3];
      gather_EB_tile_nested<PUSH_CELLS>(Et, Bt, w, t, Ep, Bp);
      const double vo[3] = {r[u][3], r[u][4], r[u][5]};
      double v[3] = {vo[0], vo[1], vo[2]};
      boris_update_vEB(g.dt, qm, Ep, Bp, v);
      vx[i] = v[0];
      vy[i] = v[1];
      vz[i] = v[2];
      if (WORK) {
        const double vs[3] = {vo[0] + v[0], vo[1] + v[1], vo[2] + v[2]};
        work += qn_Np * 0.5 * dot3(vs, Ep);
      }
    }
  }
  if (WORK) {
    __shared__ double sh[PUSH_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    work = warp_sum(work);
    if (lane == 0) sh[wid] = work;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < PUSH_THREADS / 32; ++q) t += sh[q];
      partial[blockIdx.x] = t;
    }
  }
}

__global__ void __launch_bounds__(RED_THREADS) k_sum_array(const double* __restrict__ v, int64_t n, double* __restrict__ partial)
{
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += v[i];
  __shared__ double sh[RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  acc = warp_sum(acc);
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q];
    partial[(int64_t)blockIdx.x * RED_MAXV] = t;
  }
}

int reduce_finish(xb_ctx* c, int nv, double* host_out);  // fields.cu

// second push of ecsimcorr without the half move: Boris update + predicted work (all ranks' sum)
int push_second_work(xb_ctx* c, Species& s, const double* Eh, const double* B, double* pred_w)
{
  const Grid& g = c->g;
  if (!s.sorted) XB_FAIL("push_second: particles are not sorted");
  double** p = s.p[s.cur];
  const int groups_x = (g.nx + PUSH_CELLS - 1) / PUSH_CELLS;
  const int64_t blocks = (int64_t)groups_x * g.ny * g.nzl;
  if (blocks > g.ntot) XB_FAIL("push_second: partial buffer too small");
  XB_LAUNCH(c, k_push_second<true>, (int)blocks, PUSH_THREADS, 0, g, s.bin_start, p[0], p[1], p[2], p[3], p[4], p[5], Eh, B, s.q / s.m, groups_x,
            s.q * s.n / s.Np, c->tmp);
  XB_LAUNCH(c, k_sum_array, RED_BLOCKS, RED_THREADS, 0, c->tmp, blocks, c->red_partial);
  return reduce_finish(c, 1, pred_w);
}

int push_second(xb_ctx* c, Species& s, const double* Eh, const double* B)
{
  if (s.count == 0) return 0;
  if (!s.sorted) XB_FAIL("push_second: particles are not sorted");
  const Grid& g = c->g;
  double** p = s.p[s.cur];
  const int groups_x = (g.nx + PUSH_CELLS - 1) / PUSH_CELLS;
  const int64_t blocks = (int64_t)groups_x * g.ny * g.nzl;
  XB_LAUNCH(c, k_push_second<false>, (int)blocks, PUSH_THREADS, 0, g, s.bin_start, p[0], p[1], p[2], p[3], p[4], p[5], Eh, B, s.q / s.m, groups_x, 0.0,
            nullptr);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// sum v^2 (fixed grid + fixed tree: reproducible for a given particle order)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) k_sum_v2(int64_t n, const double* __restrict__ vx, const double* __restrict__ vy,
                                                       const double* __restrict__ vz, double* __restrict__ partial)
{
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (vx[i] * vx[i] + vy[i] * vy[i]) + vz[i] * vz[i];
  __shared__ double sh[RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  acc = warp_sum(acc);
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q];
    partial[(int64_t)blockIdx.x * RED_MAXV] = t;
  }
}

int reduce_finish(xb_ctx* c, int nv, double* host_out);  // fields.cu

int kinetic_energy(xb_ctx* c, Species& s, double* sum_v2, double* K)
{
  double** p = s.p[s.cur];
  XB_LAUNCH(c, k_sum_v2, RED_BLOCKS, RED_THREADS, 0, s.count, p[3], p[4], p[5], c->red_partial);
  double w = 0.0;
  XB_CHECK(reduce_finish(c, 1, &w));
  if (sum_v2) *sum_v2 = w;
  if (K) *K = 0.5 * s.m * (s.n / (double)s.Np) * w;  // diagnostics/energy.cpp:70-88
  return 0;
}

__global__ void __launch_bounds__(RED_THREADS) k_moments(int64_t n, const double* __restrict__ vx, const double* __restrict__ vy,
                                                        const double* __restrict__ vz, double* __restrict__ partial)
{
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = vx[i], y = vy[i], z = vz[i];
    a[0] += x;
    a[1] += y;
    a[2] += z;
    a[3] += (x * x + y * y) + z * z;
  }
  __shared__ double sh[RED_THREADS / 32][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double t = warp_sum(a[k]);
    if (lane == 0) sh[wid][k] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * RED_MAXV + threadIdx.x] = t;
  }
}

int particle_moments(xb_ctx* c, Species& s, double* out5)
{
  double** p = s.p[s.cur];
  XB_LAUNCH(c, k_moments, RED_BLOCKS, RED_THREADS, 0, s.count, p[3], p[4], p[5], c->red_partial);
  XB_CHECK(reduce_finish(c, 4, out5));
  double n = (double)s.count;
  if (c->g.nranks > 1) {
    c->red_host[0] = n;
    XB_CUDA(cudaMemcpyAsync(c->red_out, c->red_host, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    XB_CHECK(comm_allreduce_sum(c, c->red_out, 1));
    XB_CUDA(cudaMemcpyAsync(c->red_host, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    n = c->red_host[0];
  }
  out5[4] = n;
  return 0;
}

__global__ void k_scale3(int64_t n, double* __restrict__ a, double* __restrict__ b, double* __restrict__ cc, double f)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[i] *= f;
  b[i] *= f;
  cc[i] *= f;
}

int scale_velocities(xb_ctx* c, Species& s, double lambda)
{
  if (s.count == 0) return 0;
  double** p = s.p[s.cur];
  XB_LAUNCH(c, k_scale3, grid_for(s.count), 256, 0, s.count, p[3], p[4], p[5], lambda);
  return 0;
}

}  // namespace xb
