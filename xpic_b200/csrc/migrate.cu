// migrate.cu -- z-slab decomposition of the particles: migration across slab boundaries and the
// boundary-plane copies ("ghost particles") the moment deposition needs.
//
// Replaces interfaces::Particles::update_cells_mpi (src/interfaces/particles.cpp:118-248): of the
// reference's 26 neighbours only +-z exist in a slab layout.  The exchange is fused into the
// counting sort, so a step still makes one pass over the particles:
//   1. k_move_key_slab: move + wrap + key; a particle whose cell left the slab is packed straight
//      into the send buffer of its direction (and dropped from the histogram);
//   2. counts, then payloads, travel to the z neighbours over NCCL (NVLink);
//   3. arrivals are keyed into the same histogram; scan; locals and arrivals are scattered into
//      the sorted SoA buffer.
// The mass matrices of rows on a slab's boundary planes receive contributions from the
// neighbour's boundary cells (inside PETSc: the off-rank COO entries of MatSetValuesCOO,
// src/impls/ecsim/simulation.cpp:366).  Instead of exchanging 10 KB cell blocks, each rank
// receives a copy of the neighbour's boundary-plane particles (48 B each, already sorted, with
// their bin table) and computes those cells' blocks itself, bit-identically to the owner.
#include <algorithm>

#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"
#include "stencil.cuh"

namespace xb {

int sort_scan_and_scatter(xb_ctx* c, Species& s, int64_t nlocal, const MigrateBuffers* arrivals, int64_t n_from_down, int64_t n_from_up,
                          double dt_move);  // particles.cu
int deposit_cells(xb_ctx* c, Species& s, const double* const* p, const int32_t* bin_start, int64_t bin_cell0, int64_t ncells, int64_t stage_cell0,
                  int zshift, double** rec, int64_t rec_stride, int64_t nparticles);  // deposit.cu

static int ensure_buffers(xb_ctx* c, Species& s)
{
  if (s.mig) return 0;
  MigrateBuffers* m = new MigrateBuffers();
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
    XB_CUDA(cudaMalloc(&m->recv_key[d], sizeof(int32_t) * m->cap));
  }
  XB_CUDA(cudaMalloc(&m->counts_dev, sizeof(unsigned long long) * 4));
  XB_CUDA(cudaMallocHost(&m->counts_host, sizeof(unsigned long long) * 4));
  s.mig = m;
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
    cudaFree(s.mig->recv_key[d]);
  }
  cudaFree(s.mig->counts_dev);
  cudaFreeHost(s.mig->counts_host);
  delete s.mig;
  s.mig = nullptr;
}

struct SendPtrs {
  double* a[2][7];
};

// as k_move_key (particles.cu), plus: leavers are packed into send[dir] and get key -1
__global__ void k_move_key_slab(Grid g, int64_t n, double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                                const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                                const uint64_t* __restrict__ id, double dtm, int32_t* __restrict__ key, int32_t* __restrict__ hist, SendPtrs sp,
                                unsigned long long* __restrict__ send_count, int64_t cap)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = moved_coord(x[i], dtm != 0.0 ? vx[i] : 0.0, dtm, g.Lx);
  const double py = moved_coord(y[i], dtm != 0.0 ? vy[i] : 0.0, dtm, g.Ly);
  const double pz = moved_coord(z[i], dtm != 0.0 ? vz[i] : 0.0, dtm, g.Lz);
  x[i] = px;
  y[i] = py;
  z[i] = pz;
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
  if ((int64_t)slot >= cap) return;  // overflow is detected on the host from the count
  sp.a[dir][0][slot] = px;
  sp.a[dir][1][slot] = py;
  sp.a[dir][2][slot] = pz;
  sp.a[dir][3][slot] = vx[i];
  sp.a[dir][4][slot] = vy[i];
  sp.a[dir][5][slot] = vz[i];
  if (id) reinterpret_cast<uint64_t*>(sp.a[dir][6])[slot] = id[i];
}

__global__ void k_key_arrivals(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                               int32_t* __restrict__ key, int32_t* __restrict__ hist, int* __restrict__ bad)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pl = slab_plane(g, z[i]);
  if (pl < 1 || pl > g.nzl) {  // a particle that jumped over a whole slab (the reference loses it too,
    key[i] = -1;               // src/interfaces/particles.cpp:183-208); flagged as an error here
    atomicAdd(bad, 1);
    return;
  }
  const int32_t k = particle_key(g, x[i], y[i], z[i], pl);
  key[i] = k;
  atomicAdd(&hist[k], 1);
}

int migrate_and_sort(xb_ctx* c, Species& s, double dt_move)
{
  const Grid& g = c->g;
  XB_CHECK(ensure_buffers(c, s));
  MigrateBuffers& m = *s.mig;
  const int64_t n = s.count;
  XB_CUDA(cudaMemsetAsync(c->hist, 0, sizeof(int32_t) * c->nbins, c->stream));
  XB_CUDA(cudaMemsetAsync(m.counts_dev, 0, sizeof(unsigned long long) * 4, c->stream));
  double** p = s.p[s.cur];
  SendPtrs sp;
  for (int d = 0; d < 2; ++d)
    for (int k = 0; k < 7; ++k) sp.a[d][k] = m.send[d][k];
  if (n > 0) {
    const int blocks = (int)((n + 255) / 256);
    XB_LAUNCH(c, k_move_key_slab, blocks, 256, 0, g, n, p[0], p[1], p[2], p[3], p[4], p[5], s.id[s.cur], dt_move, s.key, c->hist, sp, m.counts_dev,
              m.cap);
  }
  // counts: mine to the host, the neighbours' to me
  XB_CHECK(comm_exchange(c, m.counts_dev + 0, 8, m.counts_dev + 1, 8, m.counts_dev + 3, 8, m.counts_dev + 2, 8));
  XB_CUDA(cudaMemcpyAsync(m.counts_host, m.counts_dev, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t to_down = (int64_t)m.counts_host[0], to_up = (int64_t)m.counts_host[1];
  const int64_t from_down = (int64_t)m.counts_host[2], from_up = (int64_t)m.counts_host[3];
  if (to_down > m.cap || to_up > m.cap || from_down > m.cap || from_up > m.cap) XB_FAIL("particle migration buffer overflow");
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
  if (n - to_down - to_up + from_down + from_up > s.capacity) XB_FAIL("species capacity exceeded after migration");
  int* bad = reinterpret_cast<int*>(m.counts_dev);  // reuse slot 0 (counts are on the host already)
  XB_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), c->stream));
  if (from_down > 0)
    XB_LAUNCH(c, k_key_arrivals, (int)((from_down + 255) / 256), 256, 0, g, from_down, m.recv[0][0], m.recv[0][1], m.recv[0][2], m.recv_key[0], c->hist, bad);
  if (from_up > 0)
    XB_LAUNCH(c, k_key_arrivals, (int)((from_up + 255) / 256), 256, 0, g, from_up, m.recv[1][0], m.recv[1][1], m.recv[1][2], m.recv_key[1], c->hist, bad);
  XB_CHECK(sort_scan_and_scatter(c, s, n, &m, from_down, from_up, 0.0));  // k_move_key_slab already stored the moved positions
  s.count = n - to_down - to_up + from_down + from_up;
  int nbad = 0;
  XB_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  if (nbad) XB_FAIL("a particle crossed more than one slab in a single move");
  s.sorted = true;
  return 0;
}

__global__ void k_rebase_bins(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - src[0];
}

// Moments of the two ghost cell planes (stage planes 0 and nzl + 1) from copies of the
// neighbours' boundary-plane particles.
int deposit_ghost_cells(xb_ctx* c, Species& s, double* stage)
{
  (void)stage;
  const Grid& g = c->g;
  XB_CHECK(ensure_buffers(c, s));
  MigrateBuffers& m = *s.mig;
  const int64_t pb = g.plane * 8;  // bins per plane
  // my boundary planes: bin plane 1 (first owned) goes down, bin plane nzl (last owned) goes up
  int32_t h[4];
  XB_CUDA(cudaMemcpyAsync(&h[0], s.bin_start + 1 * pb, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaMemcpyAsync(&h[1], s.bin_start + 2 * pb, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaMemcpyAsync(&h[2], s.bin_start + (int64_t)g.nzl * pb, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaMemcpyAsync(&h[3], s.bin_start + (int64_t)(g.nzl + 1) * pb, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t lo0 = h[0], nlo = h[1] - h[0], hi0 = h[2], nhi = h[3] - h[2];
  m.counts_host[0] = (unsigned long long)nlo;
  m.counts_host[1] = (unsigned long long)nhi;
  XB_CUDA(cudaMemcpyAsync(m.counts_dev, m.counts_host, sizeof(unsigned long long) * 2, cudaMemcpyHostToDevice, c->stream));
  XB_CHECK(comm_exchange(c, m.counts_dev + 0, 8, m.counts_dev + 1, 8, m.counts_dev + 3, 8, m.counts_dev + 2, 8));
  XB_CUDA(cudaMemcpyAsync(m.counts_host + 2, m.counts_dev + 2, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  // counts_host[2] = from down (its top plane -> my low ghost), [3] = from up (its bottom plane -> my high ghost)
  const int64_t glo = (int64_t)m.counts_host[2], ghi = (int64_t)m.counts_host[3];
  if (glo > m.ghost_cap || ghi > m.ghost_cap) XB_FAIL("ghost particle buffer overflow");
  double** p = s.p[s.cur];
  ExchangeList l;
  l.n = 7;
  for (int k = 0; k < 6; ++k) {
    l.to_down[k] = p[k] + lo0;
    l.n_to_down[k] = sizeof(double) * nlo;
    l.to_up[k] = p[k] + hi0;
    l.n_to_up[k] = sizeof(double) * nhi;
    l.from_up[k] = m.ghost[1][k];
    l.n_from_up[k] = sizeof(double) * ghi;
    l.from_down[k] = m.ghost[0][k];
    l.n_from_down[k] = sizeof(double) * glo;
  }
  // bin tables of the two planes (pb + 1 entries each, absolute offsets; rebased after receipt)
  l.to_down[6] = s.bin_start + 1 * pb;
  l.to_up[6] = s.bin_start + (int64_t)g.nzl * pb;
  l.from_up[6] = c->cursor;            // scratch: nbins >= 2 (pb + 1) always holds (nzl + 2 >= 5 planes)
  l.from_down[6] = c->cursor + pb + 1;
  l.n_to_down[6] = l.n_to_up[6] = l.n_from_up[6] = l.n_from_down[6] = sizeof(int32_t) * (pb + 1);
  XB_CHECK(comm_exchange_list(c, l));
  const int blocks = (int)((pb + 1 + 255) / 256);
  XB_LAUNCH(c, k_rebase_bins, blocks, 256, 0, c->cursor, m.ghost_bins[1], pb + 1);
  XB_LAUNCH(c, k_rebase_bins, blocks, 256, 0, c->cursor + pb + 1, m.ghost_bins[0], pb + 1);
  // low ghost plane: the neighbour below; across the periodic boundary its z is nz planes above mine
  const int zs_lo = g.rank == 0 ? -g.nz : 0;
  const int zs_hi = g.rank == g.nranks - 1 ? +g.nz : 0;
  XB_CHECK(deposit_cells(c, s, m.ghost[0], m.ghost_bins[0], 0, g.plane, 0, zs_lo, &m.ghost_rec[0], m.ghost_cap, glo));
  XB_CHECK(deposit_cells(c, s, m.ghost[1], m.ghost_bins[1], 0, g.plane, (int64_t)(g.nzl + 1) * g.plane, zs_hi, &m.ghost_rec[1], m.ghost_cap, ghi));
  return 0;
}

}  // namespace xb
