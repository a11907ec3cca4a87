// esirkepov_mma.cu -- ECSIMCorr half-moves with an atomic-free, order-deterministic Esirkepov deposit.
//
// Replaces the same reference code as esirkepov.cu (ecsimcorr::Particles::first_push / second_push,
// src/impls/ecsimcorr/particles.cpp:27-91; EsirkepovDecomposition::process,
// src/algorithms/esirkepov_decomposition.cpp:20-103) with the structure of the moment deposition:
//
//  pass 1  k_esirkepov_cells: one CTA per 4 x-consecutive cells, three warps per cell (one per
//          current component).  With the 2nd-order spline the current of one particle is
//              Jx(i,j,k) = Px(i) * [ S'y(j) Az(k) + Sy(j) Bz(k) ],   A = 2 S' + S,  B = 2 S + S'
//          (Px = running sum of -q dx (S'x - Sx), esirkepov_decomposition.cpp:57-71; cyclic for y, z),
//          i.e. a sum over particles of rank-1 products on the cell's 6 x 6 x 6 node window
//          [c - 2, c + 3]^3 (the spline support of the old and the new position; nodes outside the
//          reference's own window get an exact zero).  Four particles go through one
//          mma.sync.m8n8k4.f64: rows = the prefix axis (6 of 8 used), columns = the fast transverse
//          axis (6 of 8), one DMMA per value of the slow transverse axis: 18 DMMAs per 4 particles.
//          The kernel also does the half move r += v dt/2 (the second push's Boris update and
//          predicted work run before it in k_push_second<true>, particles.cu).
//          Finished 648-entry cell blocks leave in coalesced runs to the staging area.
//  pass 2  k_esirkepov_gather: one thread per (node, component) sums, in a fixed order, the 216 cell
//          blocks whose window contains the node, into the ghosted current (+=).
#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"
#include "stencil.cuh"

namespace xb {

int reduce_finish(xb_ctx* c, int nv, double* host_out);  // fields.cu

constexpr int EW = 6;                 // window nodes per axis: c - 2 .. c + 3
constexpr int EBLOCK = 3 * EW * EW * EW;  // 648 entries per cell: [comp][k][j][i]
constexpr int ECHUNK = 32;            // particles per round
// per-axis record: S'[6], S[6], A[6], B[6], P[6] (+1: odd stride)
constexpr int EAX = 31;
constexpr int EREC = 3 * EAX;         // 93 doubles per particle
constexpr int ECELLS = CELL_GROUP;    // 4 cells per CTA
constexpr int EWARPS = 3 * ECELLS;
constexpr int ESMEM_PER_CELL = ECHUNK * EREC;
static_assert(ECHUNK * EREC >= EBLOCK, "the cell block aliases the record buffer");

__device__ __forceinline__ double spline2e(double s)
{
  s = fabs(s);
  if (s <= 0.5) return (0.75 - s * s);
  if (0.5 < s && s < 1.5) return 0.5 * (1.5 - s) * (1.5 - s);
  return 0.0;
}

__device__ __forceinline__ void dmma_e(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cell_barrier(int slot)
{
  switch (slot) {
    case 0: asm volatile("bar.sync 1, 96;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 96;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 96;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 96;" ::: "memory"); break;
  }
}

struct EsirkepovArgs {
  double* p[6];
  const int32_t* bin_start;
  double q_alpha;  // q n / Np / (6 dt), src/impls/ecsimcorr/particles.cpp:130
  double* stage;
  int* flag;       // set when a particle moves a full cell or more in one half move
};

// axis data of one particle: fills r[0..30) = S'[6], S[6], A[6], B[6], P[6]
__device__ __forceinline__ void axis_record(double po, double pn, int c0, double q_axis, double* __restrict__ r)
{
  double run = 0.0;
#pragma unroll
  for (int i = 0; i < EW; ++i) {
    const double gx = (double)(c0 - 2 + i);
    const double So = spline2e(po - gx), Sn = spline2e(pn - gx);
    r[0 + i] = Sn;
    r[6 + i] = So;
    r[12 + i] = 2.0 * Sn + So;
    r[18 + i] = 2.0 * So + Sn;
    run += -q_axis * (Sn - So);
    r[24 + i] = run;
  }
}

// One warp = one current component (AXIS) of one cell.
//   Jx: rows x (Px), tiles over y (S'y, Sy), columns z (Az, Bz)
//   Jy: rows y (Py), tiles over x (S'x, Sx), columns z (Az, Bz)
//   Jz: rows z (Pz), tiles over x (Ax, Bx),  columns y (S'y, Sy)
template <int AXIS>
__device__ __forceinline__ void cell_axis(const Grid& g, const EsirkepovArgs& a, double* __restrict__ rec, int slot, int lane, int c_axis, int32_t p0,
                                          int32_t p1, double (&acc)[EW][2])
{
  constexpr int SLOW = AXIS == 0 ? 1 : 0, FAST = AXIS == 2 ? 1 : 2;
  constexpr int SLOW_OFF = AXIS == 2 ? 12 : 0;  // (A, B) of x for Jz, (S', S) otherwise
  constexpr int FAST_OFF = AXIS == 2 ? 0 : 12;  // (S', S) of y for Jz, (A, B) of z otherwise
  const double d_axis = AXIS == 0 ? g.dx : (AXIS == 1 ? g.dy : g.dz);
  const double inv_axis = AXIS == 0 ? g.inv_dx : (AXIS == 1 ? g.inv_dy : g.inv_dz);
  const int exact = g.exact_inv & (1 << AXIS);
  const int gq = lane >> 2, q = lane & 3;
  const int gc = gq < EW ? gq : 0;
  const double h = 0.5 * g.dt;
  double* __restrict__ pos = a.p[AXIS];
  const double* __restrict__ vel = a.p[3 + AXIS];
  for (int32_t base = p0; base < p1; base += ECHUNK) {
    const int n = min(ECHUNK, p1 - base);
    const int32_t i = base + lane;
    double rn = 0.0;
    if (lane < n) {
      const double ro = pos[i];
      rn = ro + vel[i] * h;
      const double po = to_cells(ro, d_axis, inv_axis, exact), pn = to_cells(rn, d_axis, inv_axis, exact);
      if (fabs(pn - po) >= 1.0) atomicOr(a.flag, 1);
      axis_record(po, pn, c_axis, a.q_alpha * d_axis, rec + lane * EREC + AXIS * EAX);
      pos[i] = rn;  // only this warp reads or writes this axis' coordinate
    }
    cell_barrier(slot);  // all three axis records of the chunk are complete
    for (int gs = 0; gs < n; gs += 4) {
      const bool valid = gs + q < n;
      const double* r = rec + min(gs + q, n - 1) * EREC;
      // A operand: prefix sums along the own axis at node gq (rows 6, 7 are padding)
      const double av = (valid && gq < EW) ? r[AXIS * EAX + 24 + gq] : 0.0;
      double f0 = r[FAST * EAX + FAST_OFF + gc], f1 = r[FAST * EAX + FAST_OFF + 6 + gc];  // column factors at node gq
      if (gq >= EW) f0 = f1 = 0.0;
#pragma unroll
      for (int t = 0; t < EW; ++t) {
        const double s0 = r[SLOW * EAX + SLOW_OFF + t], s1 = r[SLOW * EAX + SLOW_OFF + 6 + t];  // same for all lanes of a particle
        dmma_e(acc[t][0], acc[t][1], av, s0 * f0 + s1 * f1);
      }
    }
    cell_barrier(slot);  // the records may be overwritten by the next chunk
  }
}

__global__ void __launch_bounds__(EWARPS * 32, 2) k_esirkepov_cells(Grid g, EsirkepovArgs a, int groups_x)
{
  extern __shared__ double smem[];
  double* cells = smem;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = wid / 3, axis = wid % 3;
  double* rec = cells + (size_t)slot * ESMEM_PER_CELL;

  const int gx = blockIdx.x % groups_x, row = blockIdx.x / groups_x;
  const int cy = row % g.ny, zl = row / g.ny;
  const int cx0 = gx * TILE_CELLS, ncell = min(TILE_CELLS, g.nx - cx0);
  const bool live = slot < ncell;
  const int cx = cx0 + slot;
  const int gq = lane >> 2, q = lane & 3;
  double acc[EW][2];
#pragma unroll
  for (int t = 0; t < EW; ++t) acc[t][0] = acc[t][1] = 0.0;

  if (live) {
    const int64_t cell0 = ((int64_t)(zl + 1) * g.ny + cy) * g.nx + cx;
    const int32_t p0 = a.bin_start[cell0 << 3], p1 = a.bin_start[(cell0 + 1) << 3];
    // global cell indices: positions are global coordinates
    if (axis == 0)
      cell_axis<0>(g, a, rec, slot, lane, cx, p0, p1, acc);
    else if (axis == 1)
      cell_axis<1>(g, a, rec, slot, lane, cy, p0, p1, acc);
    else
      cell_axis<2>(g, a, rec, slot, lane, zl + g.z0, p0, p1, acc);
  }
  // accumulators -> the cell's [comp][k][j][i] block (aliases the record buffer)
  cell_barrier(slot);
  if (gq < EW) {
#pragma unroll
    for (int t = 0; t < EW; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 2 * q + e;
        if (col >= EW) continue;
        int ii, jj, kk;
        if (axis == 0) { ii = gq; jj = t; kk = col; }        // rows x, tiles y, columns z
        else if (axis == 1) { jj = gq; ii = t; kk = col; }   // rows y, tiles x, columns z
        else { kk = gq; ii = t; jj = col; }                  // rows z, tiles x, columns y
        rec[axis * (EW * EW * EW) + (kk * EW + jj) * EW + ii] = live ? acc[t][e] : 0.0;
      }
  }
  __syncthreads();
  const int64_t group = ((int64_t)zl * g.ny + cy) * groups_x + gx;
  double* out = a.stage + group * (int64_t)(EBLOCK * ECELLS);
  for (int idx = threadIdx.x; idx < EBLOCK * ECELLS; idx += EWARPS * 32) {
    const int e = idx / ECELLS, w = idx % ECELLS;
    out[idx] = cells[(size_t)w * ESMEM_PER_CELL + e];
  }
}

// J(node, comp) += sum over the 216 cells whose window [c - 2, c + 3]^3 contains the node
__global__ void __launch_bounds__(128) k_esirkepov_gather(Grid g, const double* __restrict__ stage, double* __restrict__ J, int groups_x)
{
  const int64_t total = (int64_t)g.plane * (g.nzl + 2 * GZ) * 3;
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= total) return;
  // x fastest across threads (coalesced staging reads), component slowest
  const int x = (int)(id % g.nx), y = (int)((id / g.nx) % g.ny);
  const int zl = (int)((id / g.plane) % (g.nzl + 2 * GZ)) - GZ, comp = (int)(id / (g.plane * (g.nzl + 2 * GZ)));
  double s = 0.0;
  for (int K = 0; K < EW; ++K) {
    const int cz = zl + 2 - K;
    if (cz < 0 || cz >= g.nzl) continue;  // ghost planes only receive from owned cells
    for (int Jn = 0; Jn < EW; ++Jn) {
      const int cy = wrap1(y + 2 - Jn, g.ny);
#pragma unroll
      for (int I = 0; I < EW; ++I) {
        const int cx = wrap1(x + 2 - I, g.nx);
        const int64_t group = ((int64_t)cz * g.ny + cy) * groups_x + cx / TILE_CELLS;
        const int e = comp * (EW * EW * EW) + (K * EW + Jn) * EW + I;
        s += __ldg(stage + group * (int64_t)(EBLOCK * ECELLS) + (int64_t)e * ECELLS + (cx % TILE_CELLS));
      }
    }
  }
  J[g.vidx(x, y, zl, comp)] += s;
}

__global__ void __launch_bounds__(RED_THREADS) k_sum_partials(const double* __restrict__ v, int64_t n, double* __restrict__ partial)
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

// half move r += v dt/2 of every particle + Esirkepov current of that move into s.currJe (ghosted, +=)
static int run_cells(xb_ctx* c, Species& s)
{
  const Grid& g = c->g;
  if (!s.sorted) XB_FAIL("esirkepov: particles are not sorted");
  if (g.nx < EW || g.ny < EW) XB_FAIL("esirkepov (tensor-core form): the box must be at least 6 cells wide in x and y");
  const size_t smem = sizeof(double) * ((size_t)ECELLS * ESMEM_PER_CELL);
  if (!c->esirkepov_attr_set) {
    XB_CUDA(cudaFuncSetAttribute(k_esirkepov_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->esirkepov_attr_set = true;
  }
  const int groups_x = (g.nx + TILE_CELLS - 1) / TILE_CELLS;
  const int64_t groups = (int64_t)groups_x * g.ny * g.nzl;
  if (groups * EBLOCK * ECELLS > ((c->stage_cells + CELL_GROUP - 1) / CELL_GROUP) * (int64_t)CELL_GROUP * STAGE_CELL)
    XB_FAIL("esirkepov: staging area too small");
  EsirkepovArgs a;
  double** p = s.p[s.cur];
  for (int k = 0; k < 6; ++k) a.p[k] = p[k];
  a.bin_start = s.bin_start;
  a.q_alpha = s.q * s.n / s.Np / (6.0 * g.dt);
  a.stage = c->stage;
  a.flag = reinterpret_cast<int*>(c->red_out + RED_MAXV - 1);
  XB_CUDA(cudaMemsetAsync(a.flag, 0, sizeof(int), c->stream));
  XB_LAUNCH(c, k_esirkepov_cells, (int)groups, EWARPS * 32, smem, g, a, groups_x);
  const int64_t total = (int64_t)g.plane * (g.nzl + 2 * GZ) * 3;
  XB_LAUNCH(c, k_esirkepov_gather, (int)((total + 127) / 128), 128, 0, g, c->stage, s.currJe, groups_x);
  int flag = 0;
  XB_CUDA(cudaMemcpyAsync(&flag, a.flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  if (flag) XB_FAIL("esirkepov: a particle moved a full cell or more in one half move (|v| dt / 2 >= d)");
  s.sorted = false;
  return 0;
}

int push_first_corr_mma(xb_ctx* c, Species& s) { return run_cells(c, s); }
// second push = Boris update + predicted work at full occupancy (particles.cu), then the same half
// move + deposit kernel as the first push (gather + Boris inside the tensor-core kernel, three
// redundant updates at 24 warps/SM, measured 19 ms slower at 128^3 x 64)
int push_second_corr_mma(xb_ctx* c, Species& s, const double* Eh, const double* B)
{
  XB_CHECK(push_second_work(c, s, Eh, B, &s.pred_w));
  return run_cells(c, s);
}

}  // namespace xb
