// migrate.cu -- z-slab decomposition of the particles: migration across slab boundaries and the
// boundary-plane copies ("ghost particles") the moment deposition needs.
//
// Replaces interfaces::Particles::update_cells_mpi (src/interfaces/particles.cpp:118-248): of the
// reference's 26 neighbours only +-z exist in a slab layout.  The exchange is fused into the
// counting sort, so a step still makes one pass over the particles:
//   1. k_move_key_slab: move + wrap + key; a particle whose cell left the slab is packed straight
//      into the send buffer of its direction (and dropped from the histogram); the particles that
//      stay are NOT rewritten -- the scatter recomputes the moved position, as on one GPU;
//   2. one all-gather tells every rank every rank's counts and limits (the reference sends the
//      counts first too, :188-197); overflow is decided from the same numbers on all ranks, so
//      every rank reports it together instead of some of them hanging in the payload exchange;
//   3. payloads travel to the z neighbours over NCCL (NVLink); arrivals are keyed into the same
//      histogram; scan; locals and arrivals are scattered into the sorted SoA buffer.
// One host synchronisation per sort (the counts size the payload messages).
//
// The mass matrices of rows on a slab's boundary planes receive contributions from the
// neighbour's boundary cells (inside PETSc: the off-rank COO entries of MatSetValuesCOO,
// src/impls/ecsim/simulation.cpp:366).  Instead of exchanging 10 KB cell blocks, each rank
// receives a copy of the neighbour's boundary-plane particles (48 B each, already sorted, with
// their bin table) and computes those cells' blocks itself, bit-identically to the owner.  That
// exchange runs on a second stream underneath the moment kernel of the owned planes
// (ghost_exchange_begin / deposit_ghost_cells).
#include <algorithm>

#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"
#include "stencil.cuh"

namespace xb {

int sort_scan_and_scatter(xb_ctx* c, Species& s, int64_t nlocal, const MigrateBuffers* arrivals, int64_t n_from_down, int64_t n_from_up,
                          double dt_move);  // particles.cu
int count_after_open_sort(xb_ctx* c, Species& s);
int deposit_cells(xb_ctx* c, Species& s, const double* const* p, const int32_t* bin_start, int64_t bin_cell0, int64_t ncells, int64_t stage_cell0,
                  int zshift, double** rec, int64_t rec_stride, int64_t nparticles, int zl_first);  // deposit.cu

constexpr int SLOT = 8;  // 64-bit words every rank contributes to the all-gathered table

static int ensure_buffers(xb_ctx* c, Species& s)
{
  if (s.mig) return 0;
  MigrateBuffers* m = new MigrateBuffers();
  s.mig = m;
  const Grid& g = c->g;
  // a particle crosses a slab face when |v_z| dt exceeds its distance to the face: a few percent
  // of one plane's population per step; size for 4 planes' worth
  m->cap = std::max<int64_t>(65536, 4 * (s.capacity / std::max(1, g.nzl)));
  m->ghost_cap = std::max<int64_t>(65536, 3 * (s.capacity / std::max(1, g.nzl)));
  for (int d = 0; d < 2; ++d) {
    for (int k = 0; k < 7; ++k) {
      XB_CUDA(cudaMalloc(&m->send[d][k], sizeof(double) * m->cap));
      XB_CUDA(cudaMalloc(&m->recv[d][k], sizeof(double) * m->cap));
    }
    for (int k = 0; k < 6; ++k) XB_CUDA(cudaMalloc(&m->ghost[d][k], sizeof(double) * m->ghost_cap));
    XB_CUDA(cudaMalloc(&m->ghost_bins[d], sizeof(int32_t) * (g.plane * 8 + 1)));
    XB_CUDA(cudaMalloc(&m->ghost_bins_raw[d], sizeof(int32_t) * (g.plane * 8 + 1)));
    XB_CUDA(cudaMalloc(&m->recv_key[d], sizeof(int32_t) * m->cap));
  }
  XB_CUDA(cudaMalloc(&m->counts_dev, sizeof(unsigned long long) * 4));
  XB_CUDA(cudaMemset(m->counts_dev, 0, sizeof(unsigned long long) * 4));
  XB_CUDA(cudaMalloc(&m->table_dev, sizeof(unsigned long long) * SLOT * g.nranks));
  XB_CUDA(cudaMallocHost(&m->table_host, sizeof(unsigned long long) * SLOT * g.nranks));
  XB_CUDA(cudaEventCreateWithFlags(&m->sorted, cudaEventDisableTiming));
  XB_CUDA(cudaEventCreateWithFlags(&m->ghosts_here, cudaEventDisableTiming));
  return 0;
}

void migrate_free(Species& s)
{
  if (!s.mig) return;
  for (int d = 0; d < 2; ++d) {
    for (int k = 0; k < 7; ++k) {
      cudaFree(s.mig->send[d][k]);
      cudaFree(s.mig->recv[d][k]);
    }
    for (int k = 0; k < 6; ++k) cudaFree(s.mig->ghost[d][k]);
    cudaFree(s.mig->ghost_rec[d]);
    cudaFree(s.mig->ghost_bins[d]);
    cudaFree(s.mig->ghost_bins_raw[d]);
    cudaFree(s.mig->recv_key[d]);
  }
  cudaFree(s.mig->counts_dev);
  cudaFree(s.mig->table_dev);
  cudaFreeHost(s.mig->table_host);
  if (s.mig->sorted) cudaEventDestroy(s.mig->sorted);
  if (s.mig->ghosts_here) cudaEventDestroy(s.mig->ghosts_here);
  delete s.mig;
  s.mig = nullptr;
}

struct SendPtrs {
  double* a[2][7];
};

// as k_move_key (particles.cu), plus: leavers are packed into send[dir] and get key -1
__global__ void k_move_key_slab(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                                const uint64_t* __restrict__ id, double dtm, int32_t* __restrict__ key, int32_t* __restrict__ hist, SendPtrs sp,
                                unsigned long long* __restrict__ send_count, int64_t cap, Geometry rm, double m_mpw, unsigned long long* __restrict__ tally)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = moved_coord(x[i], dtm != 0.0 ? vx[i] : 0.0, dtm, g.Lx);
  const double py = moved_coord(y[i], dtm != 0.0 ? vy[i] : 0.0, dtm, g.Ly);
  const double pz = moved_z(g, z[i], dtm != 0.0 ? vz[i] : 0.0, dtm);
  if (left_the_box(g, pz) || removed_by_command(g, rm, px, py, pz, vx[i], vy[i], vz[i], m_mpw, tally)) {
    key[i] = -1;  // gone through an open boundary or taken out by RemoveParticles: neither kept nor sent
    return;
  }
  const int pl = slab_plane(g, pz);
  if (pl >= 1 && pl <= g.nzl) {
    const int32_t k = particle_key(g, px, py, pz, pl);
    key[i] = k;
    atomicAdd(&hist[k], 1);
    return;
  }
  const int dir = pl == 0 ? 0 : 1;  // 0: to the rank below, 1: to the rank above
  const unsigned long long slot = atomicAdd(&send_count[dir], 1ull);
  key[i] = -1;
  if ((int64_t)slot >= cap) return;  // overflow is detected from the count, by every rank (migrate_and_sort)
  // static indices into the parameter struct (a run-time `dir` index would make every thread copy it to its stack)
#define XB_SEND(k) (dir ? sp.a[1][k] : sp.a[0][k])
  XB_SEND(0)[slot] = px;
  XB_SEND(1)[slot] = py;
  XB_SEND(2)[slot] = pz;
  XB_SEND(3)[slot] = vx[i];
  XB_SEND(4)[slot] = vy[i];
  XB_SEND(5)[slot] = vz[i];
  if (id) reinterpret_cast<uint64_t*>(XB_SEND(6))[slot] = id[i];
#undef XB_SEND
}

__global__ void k_key_arrivals(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                               int32_t* __restrict__ key, int32_t* __restrict__ hist, unsigned long long* __restrict__ bad)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pl = slab_plane(g, z[i]);
  if (pl < 1 || pl > g.nzl) {  // a particle that jumped over a whole slab (the reference loses it too,
    key[i] = -1;               // src/interfaces/particles.cpp:183-208); counted, reported by every rank at the next table exchange
    atomicAdd(bad, 1ull);
    return;
  }
  const int32_t k = particle_key(g, x[i], y[i], z[i], pl);
  key[i] = k;
  atomicAdd(&hist[k], 1);
}

// this rank's row of the all-gathered table
//   migration: { to_down, to_up, particles held, species capacity, buffer capacity, lost so far, -, - }
//   ghosts   : { boundary-plane populations low / high, their first particle low / high, buffer capacity, lost so far, -, - }
__global__ void k_fill_slot(unsigned long long* __restrict__ slot, const unsigned long long* __restrict__ counts, unsigned long long n,
                            unsigned long long capacity, unsigned long long cap)
{
  slot[0] = counts[0];
  slot[1] = counts[1];
  slot[2] = n;
  slot[3] = capacity;
  slot[4] = cap;
  slot[5] = counts[2];
  slot[6] = slot[7] = 0ull;
}

__global__ void k_fill_ghost_slot(unsigned long long* __restrict__ slot, const int32_t* __restrict__ bin_start, int64_t pb, int nzl,
                                  const unsigned long long* __restrict__ counts, unsigned long long ghost_cap)
{
  const int32_t lo0 = bin_start[1 * pb], lo1 = bin_start[2 * pb], hi0 = bin_start[(int64_t)nzl * pb], hi1 = bin_start[(int64_t)(nzl + 1) * pb];
  slot[0] = (unsigned long long)(lo1 - lo0);
  slot[1] = (unsigned long long)(hi1 - hi0);
  slot[2] = (unsigned long long)lo0;
  slot[3] = (unsigned long long)hi0;
  slot[4] = ghost_cap;
  slot[5] = counts[2];
  slot[6] = slot[7] = 0ull;
}

static int gather_table(xb_ctx* c, MigrateBuffers& m, cudaStream_t stream)
{
  XB_CHECK(comm_allgather(c, m.table_dev, sizeof(unsigned long long) * SLOT, stream));
  XB_CUDA(cudaMemcpyAsync(m.table_host, m.table_dev, sizeof(unsigned long long) * SLOT * c->g.nranks, cudaMemcpyDeviceToHost, stream));
  XB_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

int migrate_and_sort(xb_ctx* c, Species& s, double dt_move)
{
  const Grid& g = c->g;
  XB_CHECK(ensure_buffers(c, s));
  MigrateBuffers& m = *s.mig;
  const int64_t n = s.count;
  XB_CUDA(cudaMemsetAsync(c->hist, 0, sizeof(int32_t) * c->nbins, c->stream));
  XB_CUDA(cudaMemsetAsync(m.counts_dev, 0, sizeof(unsigned long long) * 2, c->stream));  // [2] = particles lost so far: kept
  double** p = s.p[s.cur];
  SendPtrs sp;
  for (int d = 0; d < 2; ++d)
    for (int k = 0; k < 7; ++k) sp.a[d][k] = m.send[d][k];
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT_KEYS));
  if (n > 0) {
    const int blocks = (int)((n + 255) / 256);
    XB_LAUNCH(c, k_move_key_slab, blocks, 256, 0, g, n, p[0], p[1], p[2], p[3], p[4], p[5], s.id[s.cur], dt_move, s.key, c->hist, sp, m.counts_dev,
              m.cap, c->remove, s.m * (s.n / (double)s.Np), c->removed_dev);
  }
  XB_CHECK(prof_end(c, XB_FAMILY_SORT_KEYS));
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT_MIGRATE));
  XB_LAUNCH(c, k_fill_slot, 1, 1, 0, m.table_dev + (size_t)g.rank * SLOT, m.counts_dev, (unsigned long long)n, (unsigned long long)s.capacity,
            (unsigned long long)m.cap);
  XB_CHECK(gather_table(c, m, c->stream));
  // the same table on every rank: every rank takes the same decision
  auto row = [&](int r) { return m.table_host + (size_t)((r + g.nranks) % g.nranks) * SLOT; };
  for (int r = 0; r < g.nranks; ++r) {
    const unsigned long long* t = row(r);
    if (t[5]) XB_FAIL("a particle crossed more than one slab in a single move (rank " + std::to_string(r) + ")");
    if (t[0] > t[4] || t[1] > t[4]) XB_FAIL("particle migration buffer overflow on rank " + std::to_string(r));
    const unsigned long long in = row(r - 1)[1] + row(r + 1)[0];
    if (in > 2 * t[4] || row(r - 1)[1] > t[4] || row(r + 1)[0] > t[4]) XB_FAIL("particle migration buffer overflow on rank " + std::to_string(r));
    if (t[2] - t[0] - t[1] + in > t[3]) XB_FAIL("species capacity exceeded after migration on rank " + std::to_string(r));
  }
  const int64_t to_down = (int64_t)row(g.rank)[0], to_up = (int64_t)row(g.rank)[1];
  const int64_t from_down = (int64_t)row(g.rank - 1)[1], from_up = (int64_t)row(g.rank + 1)[0];
  const int nk = c->track_ids ? 7 : 6;
  ExchangeList l;
  l.n = nk;
  for (int k = 0; k < nk; ++k) {
    l.to_down[k] = m.send[0][k];
    l.n_to_down[k] = sizeof(double) * to_down;
    l.to_up[k] = m.send[1][k];
    l.n_to_up[k] = sizeof(double) * to_up;
    l.from_up[k] = m.recv[1][k];
    l.n_from_up[k] = sizeof(double) * from_up;
    l.from_down[k] = m.recv[0][k];
    l.n_from_down[k] = sizeof(double) * from_down;
  }
  XB_CHECK(comm_exchange_list(c, l));
  unsigned long long* bad = m.counts_dev + 2;
  if (from_down > 0)
    XB_LAUNCH(c, k_key_arrivals, (int)((from_down + 255) / 256), 256, 0, g, from_down, m.recv[0][0], m.recv[0][1], m.recv[0][2], m.recv_key[0], c->hist, bad);
  if (from_up > 0)
    XB_LAUNCH(c, k_key_arrivals, (int)((from_up + 255) / 256), 256, 0, g, from_up, m.recv[1][0], m.recv[1][1], m.recv[1][2], m.recv_key[1], c->hist, bad);
  XB_CHECK(prof_end(c, XB_FAMILY_SORT_MIGRATE));
  XB_CHECK(prof_begin(c, XB_FAMILY_SORT_SCATTER));
  XB_CHECK(sort_scan_and_scatter(c, s, n, &m, from_down, from_up, dt_move));  // the scatter recomputes the moved position of the particles that stay
  XB_CHECK(prof_end(c, XB_FAMILY_SORT_SCATTER));
  s.count = n - to_down - to_up + from_down + from_up;
  if (g.open_z || c->remove.kind >= 0) XB_CHECK(count_after_open_sort(c, s));  // minus the particles that left the box / were removed
  s.sorted = true;
  return 0;
}

__global__ void k_rebase_bins(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - src[0];
}

// Marks the point of the main stream at which the sort is complete (and the previous step's ghost copies are
// consumed): the exchange below starts there, the moment kernel of the owned planes is launched after it.
int ghost_exchange_mark(xb_ctx* c, Species& s)
{
  XB_CHECK(ensure_buffers(c, s));
  XB_CUDA(cudaEventRecord(s.mig->sorted, c->stream));
  return 0;
}

// Starts the exchange of the boundary-plane particles on the copy stream: the caller launches the moment
// kernel of the owned planes on the main stream right before, so the messages (and the host synchronisation
// that sizes them) are hidden underneath it.  Collective: every rank calls it for every sort.
int ghost_exchange_begin(xb_ctx* c, Species& s)
{
  const Grid& g = c->g;
  XB_CHECK(ensure_buffers(c, s));
  MigrateBuffers& m = *s.mig;
  cudaStream_t cs = c->copy_stream;
  const int64_t pb = g.plane * 8;  // bins per plane
  XB_CUDA(cudaStreamWaitEvent(cs, m.sorted, 0));  // recorded by ghost_exchange_mark before the owned planes' kernel
  // my boundary planes: bin plane 1 (first owned) goes down, bin plane nzl (last owned) goes up
  k_fill_ghost_slot<<<1, 1, 0, cs>>>(m.table_dev + (size_t)g.rank * SLOT, s.bin_start, pb, g.nzl, m.counts_dev, (unsigned long long)m.ghost_cap);
  c->launches++;
  XB_CUDA(cudaGetLastError());
  XB_CHECK(gather_table(c, m, cs));
  auto row = [&](int r) { return m.table_host + (size_t)((r + g.nranks) % g.nranks) * SLOT; };
  for (int r = 0; r < g.nranks; ++r) {
    const unsigned long long* t = row(r);
    if (t[5]) XB_FAIL("a particle crossed more than one slab in a single move (rank " + std::to_string(r) + ")");
    const bool has_down = !(g.open_z && r == 0), has_up = !(g.open_z && r == g.nranks - 1);
    if ((has_down && row(r - 1)[1] > t[4]) || (has_up && row(r + 1)[0] > t[4])) XB_FAIL("ghost particle buffer overflow on rank " + std::to_string(r));
  }
  const unsigned long long* me = row(g.rank);
  // open z: no neighbour across the box ends, their ghost cell planes stay empty (the staging planes keep their zeros)
  const bool down = !(g.open_z && g.rank == 0), up = !(g.open_z && g.rank == g.nranks - 1);
  const int64_t nlo = down ? (int64_t)me[0] : 0, nhi = up ? (int64_t)me[1] : 0, lo0 = (int64_t)me[2], hi0 = (int64_t)me[3];
  // from below arrives its top plane (my low ghost plane), from above its bottom plane (my high ghost plane)
  m.nghost[0] = down ? (int64_t)row(g.rank - 1)[1] : -1;
  m.nghost[1] = up ? (int64_t)row(g.rank + 1)[0] : -1;
  double** p = s.p[s.cur];
  ExchangeList l;
  l.n = 7;
  for (int k = 0; k < 6; ++k) {
    l.to_down[k] = p[k] + lo0;
    l.n_to_down[k] = sizeof(double) * nlo;
    l.to_up[k] = p[k] + hi0;
    l.n_to_up[k] = sizeof(double) * nhi;
    l.from_up[k] = m.ghost[1][k];
    l.n_from_up[k] = up ? sizeof(double) * m.nghost[1] : 0;
    l.from_down[k] = m.ghost[0][k];
    l.n_from_down[k] = down ? sizeof(double) * m.nghost[0] : 0;
  }
  // bin tables of the two planes (pb + 1 entries each, absolute offsets; rebased after receipt)
  l.to_down[6] = s.bin_start + 1 * pb;
  l.to_up[6] = s.bin_start + (int64_t)g.nzl * pb;
  l.from_up[6] = m.ghost_bins_raw[1];
  l.from_down[6] = m.ghost_bins_raw[0];
  l.n_to_down[6] = l.n_from_down[6] = down ? sizeof(int32_t) * (pb + 1) : 0;
  l.n_to_up[6] = l.n_from_up[6] = up ? sizeof(int32_t) * (pb + 1) : 0;
  XB_CHECK(comm_exchange_list(c, l, cs));
  const int blocks = (int)((pb + 1 + 255) / 256);
  for (int d = 0; d < 2; ++d) {
    if (m.nghost[d] < 0) continue;
    k_rebase_bins<<<blocks, 256, 0, cs>>>(m.ghost_bins_raw[d], m.ghost_bins[d], pb + 1);
    c->launches++;
  }
  XB_CUDA(cudaGetLastError());
  XB_CUDA(cudaEventRecord(m.ghosts_here, cs));
  return 0;
}

// Moments of the ghost cell planes below / above the slab, from the copies ghost_exchange_begin fetched, into the
// staging cells stage_lo / stage_hi on.
int deposit_ghost_cells(xb_ctx* c, Species& s, int64_t stage_lo, int64_t stage_hi, bool do_lo, bool do_hi)
{
  const Grid& g = c->g;
  MigrateBuffers& m = *s.mig;
  XB_CUDA(cudaStreamWaitEvent(c->stream, m.ghosts_here, 0));
  // low ghost plane: the neighbour below; across the periodic boundary its z is nz planes above mine
  const int zs_lo = g.rank == 0 ? -g.nz : 0;
  const int zs_hi = g.rank == g.nranks - 1 ? +g.nz : 0;
  if (do_lo && m.nghost[0] >= 0)
    XB_CHECK(deposit_cells(c, s, m.ghost[0], m.ghost_bins[0], 0, g.plane, stage_lo, zs_lo, &m.ghost_rec[0], m.ghost_cap, m.nghost[0], -1));
  if (do_hi && m.nghost[1] >= 0)
    XB_CHECK(deposit_cells(c, s, m.ghost[1], m.ghost_bins[1], 0, g.plane, stage_hi, zs_hi, &m.ghost_rec[1], m.ghost_cap, m.nghost[1], g.nzl));
  return 0;
}

}  // namespace xb
