// deposit.cu -- ECSIM moments: implicit current I and the 3x3-block particle mass matrices L.
//
// Replaces ecsim::Particles::fill_ecsim_current / decompose_ecsim_current
// (src/impls/ecsim/particles.cpp:33-173), Simulation::fill_matrix_indices (simulation.cpp:370-469)
// and PETSc's MatSetPreallocationCOO / MatSetValuesCOO duplicate summation (:359,366).
//
// Two atomic-free, order-deterministic passes:
//  1. k_cell_blocks: one warp per cell.  Particles are sorted by (cell, half-cell octant), so all
//     particles of a bin share one 24-point footprint; every lane owns 18 fixed entries of the
//     octant's 24 x 24 (component, point) outer-product block in registers and loops over the bin's
//     particles (per-particle weights / alpha / I_p are computed once by one lane each and
//     broadcast through shared memory).  After each octant the lane adds its registers into the
//     cell's private 9 x 12 x 12 block (the reference's coo_v layout, particles.cpp:145-163) kept
//     in shared memory; finished blocks are written to a staging area in coalesced 64-byte runs
//     (stage[group of 8 cells][entry][cell % 8]).
//  2. k_gather_rows / k_gather_current: one thread per (node, component pair); sums, in a fixed
//     order, the entries of the <= 12 neighbouring cell blocks that land on each of the node's
//     fixed-offset coefficient slots and writes the coefficient planes coalesced.
//
// Roofline: pass 1 is FP64-FMA bound (576 FMA per particle), pass 2 HBM bound
// (1332 * 8 B read + 369 * 8 B written per cell).
#include "comm.cuh"
#include "common.cuh"
#include "deposit.cuh"
#include "gather.cuh"
#include "stencil.cuh"

namespace xb {

constexpr int DEP_WARPS = CELL_GROUP;  // one warp per cell
constexpr int DEP_CHUNK = 16;          // particles staged per round
constexpr int REC = 36;                // s[3][8], A*alpha[9], I_p[3]
constexpr int DEP_SMEM_PER_WARP = BLOCK_ALL + DEP_CHUNK * REC;

__global__ void __launch_bounds__(DEP_WARPS * 32) k_cell_blocks(Grid g, DepositArgs a, const double* __restrict__ B, double* __restrict__ stage)
{
  extern __shared__ double smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* block = smem + (size_t)wid * DEP_SMEM_PER_WARP;
  double* rec = block + BLOCK_ALL;
  const int64_t cell_local = (int64_t)blockIdx.x * CELL_GROUP + wid;

  for (int e = lane; e < BLOCK_ALL; e += 32) block[e] = 0.0;
  __syncwarp();

  if (cell_local < a.ncells) {
    const int64_t bin0 = (a.bin_cell0 + cell_local) << 3;
    const int g1 = lane >> 2;      // row point t1
    const int h2 = (lane & 3) * 2;  // column points t2 = h2, h2 + 1
    const int cI = lane >> 3, tI = lane & 7;  // current entry owned by lanes < 24
    for (int oct = 0; oct < 8; ++oct) {
      const int32_t b0 = a.bin_start[bin0 + oct], b1 = a.bin_start[bin0 + oct + 1];
      if (b0 == b1) continue;
      double acc[3][3][2];
#pragma unroll
      for (int c1 = 0; c1 < 3; ++c1)
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) acc[c1][c2][0] = acc[c1][c2][1] = 0.0;
      double accI = 0.0;
      for (int32_t base = b0; base < b1; base += DEP_CHUNK) {
        const int cnt = min(DEP_CHUNK, b1 - base);
        if (lane < cnt) {
          const int32_t i = base + lane;
          const double px = a.p[0][i], py = a.p[1][i], pz = a.p[2][i];
          const double v[3] = {a.p[3][i], a.p[4][i], a.p[5][i]};
          Weights w;
          make_weights(g, px, py, pz, a.zshift, w);
          NodeOffsets off;
          make_offsets(g, w, off);
          double Bp[3], b[3];
          gather_B(g, B, w, off, Bp);
          const double f = (0.5 * g.dt) * a.q / a.m;
#pragma unroll
          for (int c = 0; c < 3; ++c) b[c] = Bp[c] * f;
          double vxb[3];
          cross3(v, b, vxb);
          const double vb = dot3(v, b), b2 = dot3(b, b);
          double* r = rec + lane * REC;
          const double ci = a.q * a.mpw / (1. + b2);
#pragma unroll
          for (int c = 0; c < 3; ++c) r[33 + c] = ci * (v[c] + vxb[c] + vb * b[c]);
          const double Ap = 0.5 * g.dt * g.dt * a.mpw * a.q * a.q / a.m / (1 + b2);
          r[24 + 0] = Ap * (1.0 + b[0] * b[0]);
          r[24 + 1] = Ap * (+b[2] + b[0] * b[1]);
          r[24 + 2] = Ap * (-b[1] + b[0] * b[2]);
          r[24 + 3] = Ap * (-b[2] + b[1] * b[0]);
          r[24 + 4] = Ap * (1.0 + b[1] * b[1]);
          r[24 + 5] = Ap * (+b[0] + b[1] * b[2]);
          r[24 + 6] = Ap * (+b[1] + b[2] * b[0]);
          r[24 + 7] = Ap * (-b[0] + b[2] * b[1]);
          r[24 + 8] = Ap * (1.0 + b[2] * b[2]);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int i1 = t & 1, j1 = (t >> 1) & 1, k1 = t >> 2;
            r[0 + t] = w.wn[2][k1] * w.wn[1][j1] * w.ws[0][i1];
            r[8 + t] = w.wn[2][k1] * w.ws[1][j1] * w.wn[0][i1];
            r[16 + t] = w.ws[2][k1] * w.wn[1][j1] * w.wn[0][i1];
          }
        }
        __syncwarp();
        for (int p = 0; p < cnt; ++p) {
          const double* r = rec + p * REC;
          double s1[3], s2[3][2];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            s1[c] = r[c * 8 + g1];
            const double2 t2 = *reinterpret_cast<const double2*>(r + c * 8 + h2);
            s2[c][0] = t2.x;
            s2[c][1] = t2.y;
          }
#pragma unroll
          for (int c1 = 0; c1 < 3; ++c1)
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
              const double t = s1[c1] * r[24 + c1 * 3 + c2];
              acc[c1][c2][0] += t * s2[c2][0];
              acc[c1][c2][1] += t * s2[c2][1];
            }
          if (lane < 24) accI += r[cI * 8 + tI] * r[33 + cI];
        }
        __syncwarp();
      }
      // fold the octant's registers into the cell block
      const int ox = oct & 1, oy = (oct >> 1) & 1, oz = oct >> 2;
#pragma unroll
      for (int c1 = 0; c1 < 3; ++c1) {
        const int row = block_pos(c1, g1, ox, oy, oz);
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) {
          const int e = (c1 * 3 + c2) * 144 + row * 12;
          block[e + block_pos(c2, h2, ox, oy, oz)] += acc[c1][c2][0];
          block[e + block_pos(c2, h2 + 1, ox, oy, oz)] += acc[c1][c2][1];
        }
      }
      if (lane < 24) block[BLOCK_MAT + cI * 12 + block_pos(cI, tI, ox, oy, oz)] += accI;
      __syncwarp();
    }
  }
  __syncthreads();
  // coalesced write-out: stage[group][entry][cell % 8]
  double* out = stage + ((a.stage_cell0 / CELL_GROUP) + blockIdx.x) * (int64_t)(BLOCK_ALL * CELL_GROUP);
  for (int idx = threadIdx.x; idx < BLOCK_ALL * CELL_GROUP; idx += DEP_WARPS * 32) {
    const int e = idx / CELL_GROUP, w = idx % CELL_GROUP;
    out[idx] = smem[(size_t)w * DEP_SMEM_PER_WARP + e];
  }
}

// ---------------------------------------------------------------------------------------------
// pass 1, tensor-core form.  For one octant the 24 x 24 block is a sum of rank-1 updates
//   D[(c1,t1)][(c2,t2)] += s_c1(t1) * (A_p alpha_c1c2) * s_c2(t2)       over the bin's particles,
// i.e. nine 8 x k x 8 products with k = particles.  mma.sync.m8n8k4.f64 (DMMA, the only fp64
// tensor path on sm_100a) takes four particles per instruction: lane (g = lane / 4, q = lane % 4)
// supplies A[t1 = g][k = q] = s_c1(g) of particle q and B[k = q][t2 = g] = (A alpha)_c1c2 * s_c2(g)
// of the same particle, and receives D[g][2q], D[g][2q + 1] -- exactly the 18 entries per lane of
// the scalar kernel above, so the fold into the cell block is shared.  The current I uses three
// more DMMAs with B = I_p in column 0.  Per four particles a lane issues 12 shared loads,
// 9 multiplies, 9 DMMAs and 3 FMAs (the current) instead of ~180 scalar instructions.
// ---------------------------------------------------------------------------------------------
constexpr int MMA_CHUNK = 32;
constexpr int SREC = 25;  // shape record: 24 weights, odd stride (conflict-free stores and loads)
constexpr int FREC = 13;  // field record: 9 A_p alpha + 3 I_p, odd stride
// per cell: the 12 x 12 x 9 block + currents, 32 shape records, and two buffers of asynchronously
// staged inputs of the next chunk (x, y, z and the 12 field values of each particle)
constexpr int MMA_SMEM_PER_CELL = BLOCK_ALL + MMA_CHUNK * SREC + 2 * MMA_CHUNK * FREC + 2 * 3 * MMA_CHUNK;
constexpr int MMA_WARPS = 2 * CELL_GROUP;  // two warps per cell
static_assert(CELL_GROUP == 4, "pair_barrier enumerates four cell slots");

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void pair_barrier(int slot)
{
  // immediate barrier ids: the compiler then reserves 1 + CELL_GROUP barriers, not all 16
  switch (slot) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// Per-particle record, prepared once per particle and read by the MMA lanes:
//   r[0..23]  = s_c(t): E-like CIC weights of the 8 corners, 3 components   (written by warp half 0)
//   r[24..32] = A_p * alpha[c1][c2],  r[33..35] = I_p                      (written by warp half 1)
// Shape part of the record: E-like CIC weights of the 8 corners (src/impls/ecsim/particles.cpp:76-105,
// 129-131).  Warp half 0 writes components X and Y, half 1 component Z.
template <int HALF>
__device__ __forceinline__ void prepare_shapes(const Grid& g, int zshift, double px, double py, double pz, double* __restrict__ r)
{
  Weights w;
  make_weights(g, px, py, pz, zshift, w);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int i1 = t & 1, j1 = (t >> 1) & 1, k1 = t >> 2;
    if (HALF == 0) {
      r[0 + t] = w.wn[2][k1] * w.wn[1][j1] * w.ws[0][i1];
      r[8 + t] = w.wn[2][k1] * w.ws[1][j1] * w.wn[0][i1];
    }
    else {
      r[16 + t] = w.ws[2][k1] * w.wn[1][j1] * w.wn[0][i1];
    }
  }
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Field part of the record in its own kernel at full occupancy (inlined in the tensor-core kernel it
// needs ~100 registers and two dependent global loads, which starved that kernel); one CTA per 4 cells,
// B staged in a shared-memory tile:
// gather B^n, beta, A_p alpha[3][3] and I_p (src/impls/ecsim/particles.cpp:107-115).
constexpr int PF_THREADS = 256;
constexpr int PF_CELLS = 16;  // x-consecutive cells per CTA (one B tile of 18 x 3 x 3 nodes)

__global__ void __launch_bounds__(PF_THREADS) k_particle_fields(Grid g, DepositArgs a, const double* __restrict__ B, int groups_x, int zl_off)
{
  __shared__ double Bt[FieldTile<PF_CELLS>::SIZE];
  const int gx = blockIdx.x % groups_x, row = blockIdx.x / groups_x;  // row = plane index * ny + cy within this launch
  const int cy = row % g.ny, pl = row / g.ny;
  const int zl = pl + zl_off;
  const int cx0 = gx * PF_CELLS, ncell = min(PF_CELLS, g.nx - cx0);
  load_field_tile<PF_CELLS>(g, B, cx0, cy, zl, Bt, threadIdx.x, PF_THREADS);
  const int64_t cell0 = a.bin_cell0 + ((int64_t)pl * g.ny + cy) * g.nx + cx0;
  const int32_t p0 = a.bin_start[cell0 << 3], p1 = a.bin_start[(cell0 + ncell) << 3];
  __syncthreads();
  const double f = (0.5 * g.dt) * a.q / a.m;
  const int64_t st = a.rec_stride;
  for (int32_t i = p0 + threadIdx.x; i < p1; i += PF_THREADS) {
    const double v[3] = {a.p[3][i], a.p[4][i], a.p[5][i]};
    Weights w;
    make_weights(g, a.p[0][i], a.p[1][i], a.p[2][i], a.zshift, w);
    const TileIndex t = tile_index<PF_CELLS>(w, cx0, cy, zl);
    double Bp[3], b[3];
    gather_B_tile<PF_CELLS>(Bt, w, t, Bp);
#pragma unroll
    for (int c = 0; c < 3; ++c) b[c] = Bp[c] * f;
    double vxb[3];
    cross3(v, b, vxb);
    const double vb = dot3(v, b), b2 = dot3(b, b);
    double* r = a.rec + i;
    const double ci = a.q * a.mpw / (1. + b2);
#pragma unroll
    for (int c = 0; c < 3; ++c) r[(9 + c) * st] = ci * (v[c] + vxb[c] + vb * b[c]);
    const double Ap = 0.5 * g.dt * g.dt * a.mpw * a.q * a.q / a.m / (1 + b2);
    r[0 * st] = Ap * (1.0 + b[0] * b[0]);
    r[1 * st] = Ap * (+b[2] + b[0] * b[1]);
    r[2 * st] = Ap * (-b[1] + b[0] * b[2]);
    r[3 * st] = Ap * (-b[2] + b[1] * b[0]);
    r[4 * st] = Ap * (1.0 + b[1] * b[1]);
    r[5 * st] = Ap * (+b[0] + b[1] * b[2]);
    r[6 * st] = Ap * (+b[1] + b[2] * b[0]);
    r[7 * st] = Ap * (-b[0] + b[2] * b[1]);
    r[8 * st] = Ap * (1.0 + b[2] * b[2]);
  }
}

// The 9 DMMAs + 3 current FMAs of a particle group are split between the two warps of a cell so that their
// accumulators (and therefore their folds into the cell block) are disjoint:
//   half 0: (0,0) (0,1) (0,2) I_0 (1,0) (1,1)      half 1: (1,2) I_1 (2,0) (2,1) (2,2) I_2
// slot j of a half: row component op_row, column component op_col (3 = current).
//
// Where a slot's 8 x 8 tile lands inside the cell's 12 x 12 block depends only on the stagger bit
// of its row component and of its column component (src/impls/ecsim/particles.cpp:145-147), so a
// slot has 2 (c1 == c2, current) or 4 (c1 != c2) distinct placements ("variants") over the eight
// octants.  Each lane keeps one accumulator pair per (slot, variant) -- 18 pairs per half -- and
// the octants of a cell accumulate straight into them: the block in shared memory is touched once
// per cell instead of once per octant.
__host__ __device__ constexpr int op_row(int half, int j) { return half == 0 ? (j < 4 ? 0 : 1) : (j < 2 ? 1 : 2); }
__host__ __device__ constexpr int op_col(int half, int j)
{
  return half == 0 ? (j < 3 ? j : (j == 3 ? 3 : j - 4)) : (j == 0 ? 2 : (j == 1 ? 3 : (j < 5 ? j - 2 : 3)));
}
__host__ __device__ constexpr int op_nvar(int half, int j) { return (op_col(half, j) == 3 || op_col(half, j) == op_row(half, j)) ? 2 : 4; }
__host__ __device__ constexpr int op_base(int half, int j)
{
  int b = 0;
  for (int i = 0; i < j; ++i) b += op_nvar(half, i);
  return b;
}
__host__ __device__ constexpr int op_variant(int half, int j, int oct)
{
  const int c1 = op_row(half, j), c2 = op_col(half, j);
  const int o1 = (oct >> c1) & 1;
  return (c2 == 3 || c2 == c1) ? o1 : o1 * 2 + ((oct >> c2) & 1);
}
constexpr int NACC = 18;
static_assert(op_base(0, 5) + op_nvar(0, 5) == NACC && op_base(1, 5) + op_nvar(1, 5) == NACC, "18 accumulator pairs per half");

// all groups of one octant segment: cnt particles whose records start at rs0 (shapes) / rf0 (fields)
template <int HALF, int OCT>
__device__ __forceinline__ void octant_segment(const double* __restrict__ rs0, const double* __restrict__ rf0, int cnt, int gq, int q,
                                               double (&acc)[NACC][2])
{
  for (int gs = 0; gs < cnt; gs += 4) {
    const bool valid = gs + q < cnt;
    const int pi = min(gs + q, cnt - 1);  // clamp: operands of padded lanes stay finite
    const double* rs = rs0 + pi * SREC;
    const double* rf = rf0 + pi * FREC;
    double s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] = rs[c * 8 + gq];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      constexpr int dummy = 0;
      (void)dummy;
      const int c1 = op_row(HALF, j), c2 = op_col(HALF, j);
      const double av = valid ? s[c1] : 0.0;
      const int v = op_base(HALF, j) + op_variant(HALF, j, OCT);
      if (c2 < 3)
        dmma(acc[v][0], acc[v][1], av, rf[c1 * 3 + c2] * s[c2]);
      else
        // the current: a DMMA would be 1/8 filled (one column); each lane keeps s_c1(g) I_c1 of its own
        // particle instead, the four particles of a row are summed once per cell (cell_half)
        acc[v][0] += av * rf[9 + c1];
    }
  }
}

template <int HALF>
__device__ __forceinline__ void cell_half(const Grid& g, const DepositArgs& a, double* __restrict__ block, double* __restrict__ cellmem, int slot,
                                          int lane, int64_t bin0)
{
  double* srec = cellmem;                              // [32][SREC]
  double* frec = srec + MMA_CHUNK * SREC;              // [2][32][FREC]
  double* xyz = frec + 2 * MMA_CHUNK * FREC;           // [2][3][32]
  const int gq = lane >> 2, q = lane & 3;
  const int32_t bs = lane < 9 ? a.bin_start[bin0 + lane] : 0;  // bin boundaries of the 8 octants
  const int32_t p0 = __shfl_sync(0xffffffffu, bs, 0), p1 = __shfl_sync(0xffffffffu, bs, 8);
  int oct = 0;
  int32_t oend = __shfl_sync(0xffffffffu, bs, 1);
  double acc[NACC][2];
#pragma unroll
  for (int v = 0; v < NACC; ++v) acc[v][0] = acc[v][1] = 0.0;

  // asynchronous staging of a chunk's inputs: half 0 fetches x, y, z, half 1 the 12 field values
  auto stage_chunk = [&](int32_t base, int buf) {
    const int n = min(MMA_CHUNK, p1 - base);
    if (lane < n) {
      const int32_t i = base + lane;
      if (HALF == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async8(xyz + (buf * 3 + c) * MMA_CHUNK + lane, a.p[c] + i);
      }
      else {
        double* dst = frec + (buf * MMA_CHUNK + lane) * FREC;
#pragma unroll
        for (int j = 0; j < 12; ++j) cp_async8(dst + j, a.rec + i + j * a.rec_stride);
      }
    }
  };
  if (p0 < p1) stage_chunk(p0, 0);
  int buf = 0;
  for (int32_t base = p0; base < p1; base += MMA_CHUNK, buf ^= 1) {
    const int n = min(MMA_CHUNK, p1 - base);
    cp_async_wait_all();
    pair_barrier(slot);  // both halves' copies of this chunk have landed; srec is free again
    if (lane < n) {
      const double* xs = xyz + buf * 3 * MMA_CHUNK;
      prepare_shapes<HALF>(g, a.zshift, xs[lane], xs[MMA_CHUNK + lane], xs[2 * MMA_CHUNK + lane], srec + lane * SREC);
    }
    if (base + MMA_CHUNK < p1) stage_chunk(base + MMA_CHUNK, buf ^ 1);  // overlaps with the MMA phase below
    pair_barrier(slot);  // shape records of both halves are complete
    int32_t pos = base;
    const int32_t cend = base + n;
    while (pos < cend) {
      while (oend <= pos) {  // the octant's particles are exhausted (warp-uniform)
        ++oct;
        oend = __shfl_sync(0xffffffffu, bs, oct + 1);
      }
      const int32_t seg_end = min(oend, cend);
      const double* rs0 = srec + (pos - base) * SREC;
      const double* rf0 = frec + (buf * MMA_CHUNK + (pos - base)) * FREC;
      const int cnt = seg_end - pos;
      switch (oct) {
        case 0: octant_segment<HALF, 0>(rs0, rf0, cnt, gq, q, acc); break;
        case 1: octant_segment<HALF, 1>(rs0, rf0, cnt, gq, q, acc); break;
        case 2: octant_segment<HALF, 2>(rs0, rf0, cnt, gq, q, acc); break;
        case 3: octant_segment<HALF, 3>(rs0, rf0, cnt, gq, q, acc); break;
        case 4: octant_segment<HALF, 4>(rs0, rf0, cnt, gq, q, acc); break;
        case 5: octant_segment<HALF, 5>(rs0, rf0, cnt, gq, q, acc); break;
        case 6: octant_segment<HALF, 6>(rs0, rf0, cnt, gq, q, acc); break;
        default: octant_segment<HALF, 7>(rs0, rf0, cnt, gq, q, acc); break;
      }
      pos = seg_end;
    }
  }

  // the current partials of the four lanes of a row (one per particle of a group) -> lane q == 0
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (op_col(HALF, j) != 3) continue;
#pragma unroll
    for (int v = 0; v < op_nvar(HALF, j); ++v) {
      double t = acc[op_base(HALF, j) + v][0];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      acc[op_base(HALF, j) + v][0] = t;
    }
  }
  // one fold per cell: variants of a slot may land on the same entry, so they go in separate passes
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    constexpr int dummy = 0;
    (void)dummy;
    const int c1 = op_row(HALF, j), c2 = op_col(HALF, j);
#pragma unroll
    for (int v = 0; v < op_nvar(HALF, j); ++v) {
      const int o1 = op_nvar(HALF, j) == 2 ? v : (v >> 1), o2 = op_nvar(HALF, j) == 2 ? v : (v & 1);
      // block_pos only looks at the stagger bit of its own component
      const int row = block_pos(c1, gq, o1, o1, o1);
      const int k = op_base(HALF, j) + v;
      if (c2 < 3) {
        const int e = (c1 * 3 + c2) * 144 + row * 12;
        block[e + block_pos(c2, 2 * q, o2, o2, o2)] += acc[k][0];
        block[e + block_pos(c2, 2 * q + 1, o2, o2, o2)] += acc[k][1];
      }
      else if (q == 0) {
        block[BLOCK_MAT + c1 * 12 + row] += acc[k][0];
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(MMA_WARPS * 32, 2) k_cell_blocks_mma(Grid g, DepositArgs a, const double* __restrict__ B, double* __restrict__ stage)
{
  extern __shared__ double smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = wid >> 1, half = wid & 1;
  double* block = smem + (size_t)slot * MMA_SMEM_PER_CELL;
  double* cellmem = block + BLOCK_ALL;
  const int64_t cell_local = (int64_t)blockIdx.x * CELL_GROUP + slot;

  for (int e = half * 32 + lane; e < BLOCK_ALL; e += 64) block[e] = 0.0;
  pair_barrier(slot);

  if (cell_local < a.ncells) {
    const int64_t bin0 = (a.bin_cell0 + cell_local) << 3;
    if (half == 0)
      cell_half<0>(g, a, block, cellmem, slot, lane, bin0);
    else
      cell_half<1>(g, a, block, cellmem, slot, lane, bin0);
  }
  __syncthreads();
  // coalesced write-out of the CTA's four blocks: stage[group][entry][cell % 4] (per-pair strided
  // writes were measured 13 % slower: partial-sector stores)
  double* out = stage + ((a.stage_cell0 / CELL_GROUP) + blockIdx.x) * (int64_t)(BLOCK_ALL * CELL_GROUP);
  for (int idx = threadIdx.x; idx < BLOCK_ALL * CELL_GROUP; idx += MMA_WARPS * 32) {
    const int e = idx / CELL_GROUP, w = idx % CELL_GROUP;
    out[idx] = smem[(size_t)w * MMA_SMEM_PER_CELL + e];
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: gather cell blocks into the fixed-offset rows
// ---------------------------------------------------------------------------------------------
struct GatherArgs {
  const double* stage;
  int wrap;     // 1: the staging area holds exactly the owned planes of a single slab, z wraps
  int base;     // else: staging plane of cell plane zl = zl - base (whole slab + 2 ghost planes: -1; a batch: first plane - 1)
  int zl0, nplanes;  // planes whose rows this launch gathers
};

// staging id of cell (x, y, zl), zl in [-1, nzl]; -1: the cell lies outside an open z boundary
__device__ __forceinline__ int64_t stage_cell(const Grid& g, const GatherArgs& a, int x, int y, int zl)
{
  if (g.open_z && (g.z0 + zl < 0 || g.z0 + zl >= g.nz)) return -1;
  const int pz = a.wrap ? wrapi(zl, g.nzl) : zl - a.base;
  return ((int64_t)pz * g.ny + y) * g.nx + x;
}

__device__ __forceinline__ double stage_read(const double* __restrict__ stage, int64_t cell, int e)
{
  if (cell < 0) return 0.0;
  return __ldg(stage + (cell / CELL_GROUP) * (int64_t)(BLOCK_ALL * CELL_GROUP) + (int64_t)e * CELL_GROUP + (cell % CELL_GROUP));
}

template <int C1, int C2>
__global__ void __launch_bounds__(128) k_gather_rows(Grid g, GatherArgs a, double* __restrict__ coef, int accumulate)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.plane * a.nplanes) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = a.zl0 + (int)(node / g.plane);
  constexpr int NS = pair_size(C1, C2);
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = 0.0;
#pragma unroll
  for (int k1 = 0; k1 < win_n(C1, 2); ++k1)
#pragma unroll
    for (int j1 = 0; j1 < win_n(C1, 1); ++j1)
#pragma unroll
      for (int i1 = 0; i1 < win_n(C1, 0); ++i1) {
        const int o1x = win_lo(C1, 0) + i1, o1y = win_lo(C1, 1) + j1, o1z = win_lo(C1, 2) + k1;
        // the cell whose window point (i1, j1, k1) is this node
        const int64_t cell = stage_cell(g, a, wrapi(x - o1x, g.nx), wrapi(y - o1y, g.ny), zl - o1z);
        const int ebase = (C1 * 3 + C2) * 144 + win_index(C1, i1, j1, k1) * 12;
#pragma unroll
        for (int k2 = 0; k2 < win_n(C2, 2); ++k2)
#pragma unroll
          for (int j2 = 0; j2 < win_n(C2, 1); ++j2)
#pragma unroll
            for (int i2 = 0; i2 < win_n(C2, 0); ++i2) {
              const int dx = win_lo(C2, 0) + i2 - o1x, dy = win_lo(C2, 1) + j2 - o1y, dz = win_lo(C2, 2) + k2 - o1z;
              if (in_range(C1, C2, dx, dy, dz))
                acc[coef_slot(C1, C2, dx, dy, dz) - pair_base(C1, C2)] += stage_read(a.stage, cell, ebase + win_index(C2, i2, j2, k2));
            }
      }
  const TileMap tm = make_tilemap(g.nx, g.ny, g.nzl);
  double* out = coef + tm.node_offset(x, y, zl) + (int64_t)pair_base(C1, C2) * TILE_NODES;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (accumulate)
      out[s * TILE_NODES] += acc[s];
    else
      out[s * TILE_NODES] = acc[s];
  }
}

// ---- the same from variant tiles (deposit.cuh): the production path ----------------------------------------------
// row gq of a tile: 64 contiguous bytes, two 256-bit loads (whole sectors although the threads of a warp read different cells)
// The cells around a node in the staging area stage[plane][tile][cell of the plane][64]: offset d - 1 in {-1, 0, 1} per axis
struct Around {
  int xs[3], ys[3], ps[3];
  bool zok[3];  // false: the plane lies outside an open z boundary and contributes nothing
};

template <bool OPENZ>
__device__ __forceinline__ Around cells_around(const Grid& g, const GatherArgs& a, int x, int y, int zl)
{
  Around r;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int o = d - 1, zz = zl - o;
    r.xs[d] = wrapi(x - o, g.nx);
    r.ys[d] = wrapi(y - o, g.ny);
    r.ps[d] = a.wrap ? wrapi(zz, g.nzl) : zz - a.base;
    r.zok[d] = !OPENZ || !(g.z0 + zz < 0 || g.z0 + zz >= g.nz);
  }
  return r;
}

// row gq of a tile: 64 contiguous bytes, two 256-bit loads (whole sectors although the threads of a warp read different cells)
__device__ __forceinline__ void tile_row(const double* __restrict__ p, double (&r)[8])
{
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r[4]), "=d"(r[5]), "=d"(r[6]), "=d"(r[7]) : "l"(p + 4));
}

// One thread per node and component pair, as k_gather_rows; the contributions of a cell arrive as the 8 x 8 tiles of
// the pair's (ox, oy, oz) variants: tile row gq sits at window position corner_pos(C1, gq, bit of C1), column t2 at
// corner_pos(C2, t2, bit of C2).  Every index below is a compile-time constant; the sum runs in a fixed
// order (variant, tile row, column), so the result does not depend on the decomposition.
template <int C1, int C2, bool OPENZ>
__global__ void __launch_bounds__(128, 3) k_gather_tiles(Grid g, GatherArgs a, double* __restrict__ coef, int accumulate)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.plane * a.nplanes) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = a.zl0 + (int)(node / g.plane);
  constexpr int SL = C1 * 3 + C2, NS = pair_size(C1, C2), NZ = dep(SL, 2) ? 2 : 1;
  const Around ar = cells_around<OPENZ>(g, a, x, y, zl);
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = 0.0;
#pragma unroll
  for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
    for (int v = 0; v < nvar(SL); ++v) {
      const int tile = tile_id(SL, v, oz);
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        // the cell whose window point (tile row gq of this variant) is this node
        const TileTarget row = tile_target(C1, C2, v, oz, gq, 0);
        const double* p = a.stage + (((int64_t)ar.ps[row.oz + 1] * NTILE + tile) * g.plane + (ar.ys[row.oy + 1] * g.nx + ar.xs[row.ox + 1])) * 64 + gq * 8;
        double r[8];
        tile_row(p, r);
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) {
          const double val = (OPENZ && !ar.zok[row.oz + 1]) ? 0.0 : r[t2];
          acc[tile_target(C1, C2, v, oz, gq, t2).slot - pair_base(C1, C2)] += val;
        }
      }
    }
  const TileMap tm = make_tilemap(g.nx, g.ny, g.nzl);
  double* out = coef + tm.node_offset(x, y, zl) + (int64_t)pair_base(C1, C2) * TILE_NODES;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (accumulate)
      out[s * TILE_NODES] += acc[s];
    else
      out[s * TILE_NODES] = acc[s];
  }
}

__global__ void k_gather_current_tiles(Grid g, GatherArgs a, double* __restrict__ sort_currI, double* __restrict__ sim_currI)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.plane * a.nplanes) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = a.zl0 + (int)(node / g.plane);
  const Around ar = cells_around<true>(g, a, x, y, zl);
  double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        const int p = corner_pos(c, gq, o);
        const int ox = win_lo(c, 0) + pos_i(c, p), oy = win_lo(c, 1) + pos_j(c, p), oz = win_lo(c, 2) + pos_k(c, p);
        const double* t = a.stage + (((int64_t)ar.ps[oz + 1] * NTILE + TILE_CUR) * g.plane + (ar.ys[oy + 1] * g.nx + ar.xs[ox + 1])) * 64;
        const double val = __ldg(t + (gq * 4 + c) * 2 + o);
        r[c] += (g.open_z && !ar.zok[oz + 1]) ? 0.0 : val;
      }
  const int64_t o = g.vidx(x, y, zl, 0);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    sort_currI[o + c] = r[c];
    sim_currI[o + c] += r[c];
  }
}

// currI of one sort at the owned nodes; also accumulated into the simulation's total current
__global__ void k_gather_current(Grid g, GatherArgs a, double* __restrict__ sort_currI, double* __restrict__ sim_currI)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.plane * a.nplanes) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = a.zl0 + (int)(node / g.plane);
  double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < win_n(c, 2); ++k)
#pragma unroll
      for (int j = 0; j < win_n(c, 1); ++j)
#pragma unroll
        for (int i = 0; i < win_n(c, 0); ++i) {
          const int64_t cell = stage_cell(g, a, wrapi(x - (win_lo(c, 0) + i), g.nx), wrapi(y - (win_lo(c, 1) + j), g.ny), zl - (win_lo(c, 2) + k));
          r[c] += stage_read(a.stage, cell, BLOCK_MAT + c * 12 + win_index(c, i, j, k));
        }
  const int64_t o = g.vidx(x, y, zl, 0);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    sort_currI[o + c] = r[c];
    sim_currI[o + c] += r[c];
  }
}

// plain [k][node] (the C ABI's layout) <-> blocked [tile][k][t]
__global__ void k_coef_convert(Grid g, double* __restrict__ blocked, double* __restrict__ plain, int to_blocked)
{
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= g.ncl) return;
  const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = (int)(node / g.plane);
  const TileMap tm = make_tilemap(g.nx, g.ny, g.nzl);
  double* b = blocked + tm.node_offset(x, y, zl);
  for (int k = 0; k < NCOEF; ++k) {
    if (to_blocked)
      b[k * TILE_NODES] = plain[(int64_t)k * g.ncl + node];
    else
      plain[(int64_t)k * g.ncl + node] = b[k * TILE_NODES];
  }
}

int coef_convert(xb_ctx* c, double* plain_dev, bool to_blocked)
{
  const int blocks = (int)((c->g.ncl + 127) / 128);
  XB_LAUNCH(c, k_coef_convert, blocks, 128, 0, c->g, c->coef, plain_dev, to_blocked ? 1 : 0);
  return 0;
}

template <int C1, int C2>
static int launch_gather(xb_ctx* c, const GatherArgs& ga, int accumulate)
{
  const int blocks = (int)((c->g.plane * ga.nplanes + 127) / 128);
  if (c->stage_tiles && c->g.open_z)
    XB_LAUNCH(c, (k_gather_tiles<C1, C2, true>), blocks, 128, 0, c->g, ga, c->coef, accumulate);
  else if (c->stage_tiles)
    XB_LAUNCH(c, (k_gather_tiles<C1, C2, false>), blocks, 128, 0, c->g, ga, c->coef, accumulate);
  else
    XB_LAUNCH(c, (k_gather_rows<C1, C2>), blocks, 128, 0, c->g, ga, c->coef, accumulate);
  return 0;
}

// migrate.cu (multi-rank only)
int ghost_exchange_mark(xb_ctx* c, Species& s);
int ghost_exchange_begin(xb_ctx* c, Species& s);
int deposit_ghost_cells(xb_ctx* c, Species& s, int64_t stage_lo, int64_t stage_hi, bool do_lo, bool do_hi);

// cell blocks of `ncells` consecutive cells (bin space) into the staging area
// zl_first: local plane index of the first cell (the cells of one launch are whole planes: owned planes 0 .. nzl - 1,
// or one ghost plane, -1 / nzl)
int deposit_cells(xb_ctx* c, Species& s, const double* const* p, const int32_t* bin_start, int64_t bin_cell0, int64_t ncells, int64_t stage_cell0,
                  int zshift, double** rec, int64_t rec_stride, int64_t nparticles, int zl_first)
{
  // variants (xb_set_option(ctx, 0, v)): 0 fused warp-specialised DMMA kernel, variant tiles in the staging area (the
  // production path); cross-checks, all with folded cell blocks in the staging area: 4 the same kernel with the fold in
  // shared memory, 3 fused kernel without role split, 2 round-1 pipeline (field records in HBM + two-warp DMMA kernel),
  // 1 scalar-FMA cell blocks
  const int variant = c->deposit_variant;
  c->stage_tiles = variant == 0;
  DepositArgs a;
  for (int k = 0; k < 6; ++k) a.p[k] = p[k];
  a.bin_start = bin_start;
  a.bin_cell0 = bin_cell0;
  a.ncells = ncells;
  a.stage_cell0 = stage_cell0;
  a.zshift = zshift;
  a.q = s.q;
  a.m = s.m;
  a.mpw = s.n / (double)s.Np;
  a.f_beta = (0.5 * c->g.dt) * a.q / a.m;
  a.num_A = 0.5 * c->g.dt * c->g.dt * a.mpw * a.q * a.q / a.m;
  a.num_I = a.q * a.mpw;
  a.rec = nullptr;
  a.rec_stride = rec_stride;
  if ((a.stage_cell0 % CELL_GROUP) != 0) XB_FAIL("deposit: plane size must be a multiple of the staging group in multi-rank runs");
  const Grid& g = c->g;
  const int zl_off = zl_first;
  if (variant == 0 || variant == 3 || variant == 4) return launch_cell_moments(c, a, zl_off, variant);

  const bool use_mma = variant == 2;
  const size_t smem = sizeof(double) * (use_mma ? MMA_SMEM_PER_CELL : DEP_SMEM_PER_WARP) * CELL_GROUP;
  if (!c->deposit_attr_set) {
    XB_CUDA(cudaFuncSetAttribute(k_cell_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * DEP_SMEM_PER_WARP * CELL_GROUP)));
    XB_CUDA(cudaFuncSetAttribute(k_cell_blocks_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * MMA_SMEM_PER_CELL * CELL_GROUP)));
    c->deposit_attr_set = true;
  }
  const int groups = (int)((a.ncells + CELL_GROUP - 1) / CELL_GROUP);
  if (use_mma) {
    if (!*rec) XB_CUDA(cudaMalloc(rec, sizeof(double) * 12 * rec_stride));  // only the cross-check pipeline stores field records
    a.rec = *rec;
    if (nparticles > 0) {
      const int groups_x = (g.nx + PF_CELLS - 1) / PF_CELLS;
      const int64_t planes = ncells / g.plane;
      XB_LAUNCH(c, k_particle_fields, (int)(groups_x * g.ny * planes), PF_THREADS, 0, g, a, c->B, groups_x, zl_off);
    }
    XB_LAUNCH(c, k_cell_blocks_mma, groups, MMA_WARPS * 32, smem, c->g, a, c->B, c->stage);
  }
  else
    XB_LAUNCH(c, k_cell_blocks, groups, DEP_WARPS * 32, smem, c->g, a, c->B, c->stage);
  return 0;
}

static int gather_rows_of(xb_ctx* c, Species& s, const GatherArgs& ga, int acc)
{
  XB_CHECK((launch_gather<0, 0>(c, ga, acc)));
  XB_CHECK((launch_gather<0, 1>(c, ga, acc)));
  XB_CHECK((launch_gather<0, 2>(c, ga, acc)));
  XB_CHECK((launch_gather<1, 0>(c, ga, acc)));
  XB_CHECK((launch_gather<1, 1>(c, ga, acc)));
  XB_CHECK((launch_gather<1, 2>(c, ga, acc)));
  XB_CHECK((launch_gather<2, 0>(c, ga, acc)));
  XB_CHECK((launch_gather<2, 1>(c, ga, acc)));
  XB_CHECK((launch_gather<2, 2>(c, ga, acc)));
  const int blocks = (int)((c->g.plane * ga.nplanes + 127) / 128);
  if (c->stage_tiles)
    XB_LAUNCH(c, k_gather_current_tiles, blocks, 128, 0, c->g, ga, s.currI, c->currI);
  else
    XB_LAUNCH(c, k_gather_current, blocks, 128, 0, c->g, ga, s.currI, c->currI);
  return 0;
}

// My boundary planes' cell blocks -> the ghost planes of the z neighbours' staging areas (staging plane = zl + 1), on the
// exchange stream; no neighbour across an open box end (those staging planes are never read: stage_cell).
static int boundary_blocks_begin(xb_ctx* c)
{
  const Grid& g = c->g;
  if (!c->blocks_ready) {
    XB_CUDA(cudaEventCreateWithFlags(&c->blocks_ready, cudaEventDisableTiming));
    XB_CUDA(cudaEventCreateWithFlags(&c->blocks_here, cudaEventDisableTiming));
  }
  XB_CUDA(cudaEventRecord(c->blocks_ready, c->stream));
  XB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->blocks_ready, 0));
  const size_t per_plane = (size_t)g.plane * (c->stage_tiles ? STAGE_CELL : BLOCK_ALL);  // doubles; g.plane is a multiple of CELL_GROUP (deposit_cells checks)
  auto plane = [&](int sp) { return c->stage + (size_t)sp * per_plane; };
  const bool down = !(g.open_z && g.rank == 0), up = !(g.open_z && g.rank == g.nranks - 1);
  const size_t bytes = sizeof(double) * per_plane;
  XB_CHECK(comm_exchange(c, plane(1), down ? bytes : 0, plane(g.nzl), up ? bytes : 0, plane(g.nzl + 1), up ? bytes : 0, plane(0), down ? bytes : 0,
                         c->copy_stream));
  XB_CUDA(cudaEventRecord(c->blocks_here, c->copy_stream));
  return 0;
}

int deposit_moments(xb_ctx* c)
{
  const Grid& g = c->g;
  XB_CHECK(halo_fill(c, c->B, GZ));
  const bool single = g.nranks == 1;
  bool first = true;
  for (auto& s : c->sorts) {
    if (!s.sorted) XB_FAIL("deposit: particles are not sorted");
    const int acc = first ? 0 : 1;
    if (c->batch_planes == 0 && single) {
      // ---- one GPU, the staging area holds the cell blocks of the whole box: one launch, the gather wraps in z ------
      XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_CELLS));
      XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane, g.ncl, 0, 0, &s.rec, s.capacity, s.count, 0));
      XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_CELLS));
      XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_ROWS));
      XB_CHECK(gather_rows_of(c, s, GatherArgs{c->stage, 1, -1, 0, g.nzl}, acc));
      XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_ROWS));
    }
    else if (c->batch_planes == 0) {
      // ---- several slabs, the staging area holds the slab's planes plus one ghost plane on either side.  Every cell
      // block is computed once, by its owner: the two boundary planes go first and their blocks (10.6 KB per cell, 0.7 GB
      // per plane of 256 x 256 cells) travel to the z neighbours on the exchange stream while the other planes are
      // computed.  (Round 1 and the batched mode below fetch the neighbours' boundary PARTICLES instead and compute the
      // ghost planes again: 2 / nzl more cell blocks and last-bit differences across the periodic boundary.)
      // The kernel is persistent (one wave of CTAs that own their SMs until the launch ends): the inner planes go in
      // four launches so that NCCL's kernels, waiting on the high-priority stream, get their SMs at the latest where the
      // first of them ends; the dynamic work distribution of the kernel absorbs the SMs NCCL holds meanwhile.
      auto planes = [&](int p0, int np) -> int {
        return deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane * (1 + p0), g.plane * np, g.plane * (1 + p0), 0, &s.rec, s.capacity, s.count, p0);
      };
      XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_CELLS));
      XB_CHECK(planes(0, 1));
      if (g.nzl > 1) XB_CHECK(planes(g.nzl - 1, 1));
      XB_CHECK(boundary_blocks_begin(c));
      const int inner = g.nzl - 2, chunk = (inner + 3) / 4;
      for (int p0 = 1; p0 < g.nzl - 1; p0 += chunk) XB_CHECK(planes(p0, chunk < g.nzl - 1 - p0 ? chunk : g.nzl - 1 - p0));
      XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_CELLS));
      XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_GHOST));  // what is left of the exchange after the inner planes
      XB_CUDA(cudaStreamWaitEvent(c->stream, c->blocks_here, 0));
      XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_GHOST));
      XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_ROWS));
      XB_CHECK(gather_rows_of(c, s, GatherArgs{c->stage, 0, -1, 0, g.nzl}, acc));
      XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_ROWS));
    }
    else {
      // ---- large slabs: batches of P planes through a staging area of P + 2 planes (xb_create chose P so that it
      // fits the budget): staging plane 0 = the plane below the batch, 1 .. np = the batch, np + 1 = the plane above.
      // The two neighbour planes of a batch are computed again with the next / previous batch (2 / P more cell blocks).
      if (!single) {
        XB_CHECK(ghost_exchange_mark(c, s));
        XB_CHECK(ghost_exchange_begin(c, s));
      }
      for (int p0 = 0; p0 < g.nzl; p0 += c->batch_planes) {
        const int np = c->batch_planes < g.nzl - p0 ? c->batch_planes : g.nzl - p0;
        XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_CELLS));
        XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane * (1 + p0), g.plane * np, g.plane, 0, &s.rec, s.capacity, s.count, p0));
        const int64_t hi_stage = (int64_t)(np + 1) * g.plane;
        const bool open_lo = g.open_z && g.z0 + p0 - 1 < 0, open_hi = g.open_z && g.z0 + p0 + np >= g.nz;
        if (p0 > 0)  // the plane below is a plane of this slab
          XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane * p0, g.plane, 0, 0, &s.rec, s.capacity, s.count, p0 - 1));
        else if (single && !open_lo)  // ... or the top plane, through the periodic boundary
          XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane * g.nzl, g.plane, 0, -g.nz, &s.rec, s.capacity, s.count, -1));
        if (p0 + np < g.nzl)
          XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane * (1 + p0 + np), g.plane, hi_stage, 0, &s.rec, s.capacity, s.count, p0 + np));
        else if (single && !open_hi)
          XB_CHECK(deposit_cells(c, s, s.p[s.cur], s.bin_start, g.plane, g.plane, hi_stage, +g.nz, &s.rec, s.capacity, s.count, g.nzl));
        XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_CELLS));
        if (!single && (p0 == 0 || p0 + np >= g.nzl)) {  // ... or a plane of the neighbour slab
          XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_GHOST));
          XB_CHECK(deposit_ghost_cells(c, s, 0, hi_stage, p0 == 0, p0 + np >= g.nzl));
          XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_GHOST));
        }
        XB_CHECK(prof_begin(c, XB_FAMILY_MOMENTS_ROWS));
        XB_CHECK(gather_rows_of(c, s, GatherArgs{c->stage, 0, p0 - 1, p0, np}, acc));
        XB_CHECK(prof_end(c, XB_FAMILY_MOMENTS_ROWS));
      }
    }
    first = false;
  }
  if (c->sorts.empty()) XB_CUDA(cudaMemsetAsync(c->coef, 0, sizeof(double) * c->coef_elems, c->stream));
  c->coef_valid = true;
  return 0;
}

}  // namespace xb
