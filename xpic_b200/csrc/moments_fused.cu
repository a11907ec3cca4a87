// moments_fused.cu -- pass 1 of the ECSIM moment deposition in ONE kernel: per-particle field
// records (B gather, beta, A_p alpha, I_p), the rank-1 updates on the fp64 tensor path and the
// coalesced write of the finished cell blocks.
//
// Replaces ecsim::Particles::decompose_ecsim_current (src/impls/ecsim/particles.cpp:62-173); the
// arithmetic per particle is the reference's, operation by operation (weights :76-105, beta / I_p /
// alpha :107-115, s1 * s2 * A_p alpha into the cell's 9 x 12 x 12 block :145-165).
//
// Differences to the round-1 pipeline (k_particle_fields + k_cell_blocks_mma, kept in deposit.cu as a
// cross-check): no per-particle record round trip through HBM (96 B written + 120 B read per particle),
// one warp per cell instead of two (every operand is loaded once per group of four particles and feeds
// all nine DMMAs), records of 26 instead of 63 doubles (the 24 corner weights are rebuilt from the 12 axis
// weights by six multiplies per group), 128-bit shared loads, weights from the known cell of the bin
// instead of floor(), and half the shared-memory folds: accumulators are kept per (slot, ox, oy) variant
// in registers (21 pairs); the oz bit is handled by folding the z-dependent slots once when the octant
// walk passes from oz = 0 to oz = 1 and once at the end of the cell.
//
// Roofline: fp64 tensor path (mma.sync.m8n8k4.f64, nine per four particles) and the shared-memory
// pipe; HBM sees 48 B per particle in and 10.6 KB per cell out.
#include "common.cuh"
#include "deposit.cuh"
#include "gather.cuh"
#include "stencil.cuh"

namespace xb {

namespace {

constexpr int FM_CELLS = CELL_GROUP;       // cells (= warps) per CTA, one staging group
constexpr int FM_THREADS = 32 * FM_CELLS;
constexpr int FM_CHUNK = 32;               // particles per record round: one per lane
constexpr int FM_REC = 26;                 // doubles per record = 13 chunks of 16 bytes (odd: conflict-free 128-bit stores)
constexpr int FM_BLOCK = 1344;             // BLOCK_ALL = 1332 padded to 21 x 32 double2
constexpr int FM_TILE = 94;                // 3 x 3 x 3 nodes x 3 components of B (81), padded
// per cell: block, 32 records + one all-zero record (the operand of padded lanes), B tile: 2296 doubles = 18 368 B
constexpr int FM_CELL = FM_BLOCK + (FM_CHUNK + 1) * FM_REC + FM_TILE;
static_assert(FM_BLOCK >= BLOCK_ALL && FM_BLOCK % 64 == 0, "block padding");
static_assert(FM_CELL % 16 == 8 && FM_CELL % 2 == 0, "write-out reads the four blocks of a CTA without bank conflicts");
static_assert(FM_REC % 2 == 0 && (FM_REC / 2) % 2 == 1, "records are an odd number of 16-byte chunks");

// record layout in 16-byte chunks:
//   0..2  (wn, ws)[axis][lower]     3  (a00, a01)
//   4..6  (wn, ws)[axis][upper]     7  (a02, a10)
//   8 (a11, a12)   9 (a20, a21)   10 (a22, I0)   11 (I1, I2)   12 pad
// a[c1][c2] = A_p alpha[c1][c2].  The lower / upper chunks of an axis are 64 bytes apart, so the eight
// chunks a warp touches in one weight load (4 particles x 2) fall into eight different 16-byte bank groups.

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct Lane {
  int gq, q;
  int wofs[3];     // offset (doubles) of this lane's (wn, ws) chunk of axis a inside a record
  int rowpos[3];   // block_pos(c, gq) with octant bits 0
  int colpos[3];   // block_pos(c, 2 q) with octant bits 0 (2 q + 1 is the next position: i is the fastest index)
};

__device__ __forceinline__ const double2& ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// Row order of the 12 x 12 slots inside the shared-memory cell block.  A quarter warp of a 128-bit fold holds the two
// rows p, p + 1 (corner bit i) of four column chunks: with the rows in natural order they are 96 bytes apart and hit
// eight banks twice.  The order p -> 2 p (p < 6), 2 (p - 6) + 1 puts
// p and p + 1 two rows = 192 bytes apart (p = 5, 6 never form a pair: block_pos), i.e. on disjoint halves of the banks.
// The write-out walks the block in this physical order and un-permutes the staging address.
__device__ __forceinline__ constexpr int smem_row(int p) { return p < 6 ? 2 * p : 2 * (p - 6) + 1; }
__device__ __forceinline__ constexpr int slot_row(int r) { return (r & 1) ? 6 + (r >> 1) : (r >> 1); }
// staging entry of the element at physical index f of a cell block
__device__ __forceinline__ int staged_entry(int f)
{
  if (f >= BLOCK_MAT) return f;  // the 36 current partials are not permuted
  const int sl = f / 144, rem = f - sl * 144, r = rem / 12, p2 = rem - r * 12;
  return sl * 144 + slot_row(r) * 12 + p2;
}

// all groups of four particles of one octant segment: cnt particles whose records start at r0
template <int OXY>
__device__ __forceinline__ void octant_segment(const double* __restrict__ r0, const double* __restrict__ zero_rec, int cnt, const Lane& L,
                                               double (&acc)[NMAT][2], double (&cur)[NCUR])
{
  for (int gs = 0; gs < cnt; gs += 4) {
    const double* r = gs + L.q < cnt ? r0 + (gs + L.q) * FM_REC : zero_rec;  // padded lanes multiply zeros
    const double2 wx = ld2(r + L.wofs[0]), wy = ld2(r + L.wofs[1]), wz = ld2(r + L.wofs[2]);
    // E-like CIC weights of corner gq (src/impls/ecsim/particles.cpp:129-131), z * y * x as the reference multiplies
    double s[3];
    s[0] = (wz.x * wy.x) * wx.y;
    s[1] = (wz.x * wy.y) * wx.x;
    s[2] = (wz.y * wy.x) * wx.x;
    const double2 f0 = ld2(r + 6), f1 = ld2(r + 14), f2 = ld2(r + 16), f3 = ld2(r + 18), f4 = ld2(r + 20), f5 = ld2(r + 22);
    const double al[9] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y, f4.x};
    const double ip[3] = {f4.y, f5.x, f5.y};
    const double (&a)[3] = s;
#pragma unroll
    for (int c1 = 0; c1 < 3; ++c1) {
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2) {
        constexpr int dummy = 0;
        (void)dummy;
        const int v = vidx(c1 * 3 + c2, OXY);
        dmma(acc[v][0], acc[v][1], a[c1], al[c1 * 3 + c2] * s[c2]);
      }
      // the current is a matrix-vector product (a DMMA would be 1/8 filled): every lane keeps s_c1(gq) I_c1 of
      // its own particle, the four particles of a row are summed once per fold
      cur[cidx(c1, OXY)] += a[c1] * ip[c1];
    }
  }
}

// ---- the same for the variant-tile consumer, software-pipelined across groups AND octant segments: the raw operands of
// the next group of the round (the next four records, whichever octant they belong to) are requested as soon as the
// products of this group are formed, and arrive underneath its nine DMMAs.
struct Raw {
  double2 w[3], f[6];
};
__device__ __forceinline__ void load_raw(Raw& r, const double* __restrict__ rec, const Lane& L)
{
#pragma unroll
  for (int a = 0; a < 3; ++a) r.w[a] = ld2(rec + L.wofs[a]);
  r.f[0] = ld2(rec + 6);
#pragma unroll
  for (int k = 1; k < 6; ++k) r.f[k] = ld2(rec + 12 + 2 * k);
}
// live: this lane's particle belongs to the group (a padded lane holds some other record of the round: its weights are
// replaced by zeros, so it adds exact zeros as the all-zero record of octant_segment does)
template <int OXY>
__device__ __forceinline__ void group_update(Raw& raw, bool live, const double* __restrict__ next_rec, const Lane& L, double (&acc)[NMAT][2],
                                             double (&cur)[NCUR])
{
  double s[3];
  s[0] = (raw.w[2].x * raw.w[1].x) * raw.w[0].y;
  s[1] = (raw.w[2].x * raw.w[1].y) * raw.w[0].x;
  s[2] = (raw.w[2].y * raw.w[1].x) * raw.w[0].x;
  if (!live) s[0] = s[1] = s[2] = 0.0;
  const double al[9] = {raw.f[0].x, raw.f[0].y, raw.f[1].x, raw.f[1].y, raw.f[2].x, raw.f[2].y, raw.f[3].x, raw.f[3].y, raw.f[4].x};
  const double ip[3] = {raw.f[4].y, raw.f[5].x, raw.f[5].y};
  double p[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = al[k] * s[k % 3];
  // the current is a matrix-vector product (see octant_segment); done first: the raw record is dead from here on
#pragma unroll
  for (int c1 = 0; c1 < 3; ++c1) cur[cidx(c1, OXY)] += s[c1] * ip[c1];
  load_raw(raw, next_rec, L);
#pragma unroll
  for (int c1 = 0; c1 < 3; ++c1) {
#pragma unroll
    for (int c2 = 0; c2 < 3; ++c2) {
      const int v = vidx(c1 * 3 + c2, OXY);
      dmma(acc[v][0], acc[v][1], s[c1], p[c1 * 3 + c2]);
    }
  }
}

// fold register variants into the cell block.  ZDEP: the z-dependent slots (row or column component Z) and
// the Z current, at octant bit OZ, and clear them; else the five z-independent slots and the X, Y currents.
template <bool ZDEP, int OZ>
__device__ __forceinline__ void fold(double* __restrict__ block, const Lane& L, double (&acc)[NMAT][2], double (&cur)[NCUR])
{
  constexpr int st[3] = {1, 2, 4};  // block_pos moves by this much per octant bit of the component's staggered axis
  // variants of one slot may land on the same entry: one pass per variant index, all slots inside a pass
#pragma unroll
  for (int v = 0; v < 4; ++v) {
#pragma unroll
    for (int sl = 0; sl < 9; ++sl) {
      constexpr int dummy = 0;
      (void)dummy;
      if (dep(sl, 2) != ZDEP || v >= nvar(sl)) continue;
      const int c1 = sl / 3, c2 = sl % 3;
      const int o1 = vbit(sl, v, c1, OZ), o2 = vbit(sl, v, c2, OZ);
      double* e = block + sl * 144 + smem_row(L.rowpos[c1] + o1 * st[c1]) * 12 + L.colpos[c2] + o2 * st[c2];
      const int k = vbase(sl) + v;
      if (c2 == 0) {  // columns 2 q, 2 q + 1 are adjacent but not 16-byte aligned when ox = 1
        e[0] += acc[k][0];
        e[1] += acc[k][1];
      }
      else {
        double2 t = *reinterpret_cast<double2*>(e);
        t.x += acc[k][0];
        t.y += acc[k][1];
        *reinterpret_cast<double2*>(e) = t;
      }
      if (ZDEP) acc[k][0] = acc[k][1] = 0.0;
    }
    __syncwarp();
  }
#pragma unroll
  for (int v = 0; v < 2; ++v) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if ((c == 2) != ZDEP || (c == 2 && v > 0)) continue;
      const int k = cbase(c) + v;
      double t = cur[k];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      const int o = c == 2 ? OZ : v;
      if (L.q == 0) block[BLOCK_MAT + c * 12 + L.rowpos[c] + o * st[c]] += t;
      if (ZDEP) cur[k] = 0.0;
    }
    __syncwarp();
  }
}

// Persistent CTAs: one warp per cell, four x-consecutive cells (one staging group) per round, rounds strided
// over the grid.  What a cell needs from HBM before its first instruction (bin table, B tile, first 32
// particles) is requested while the previous round is folded and written out.
template <int MINB>
__global__ void __launch_bounds__(FM_THREADS, MINB) k_cell_moments(Grid g, DepositArgs a, const double* __restrict__ B, double* __restrict__ stage, int zl_off,
                                                                   int groups)
{
  extern __shared__ __align__(16) double smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* block = smem + (size_t)wid * FM_CELL;
  double* recs = block + FM_BLOCK;
  double* zero_rec = recs + FM_CHUNK * FM_REC;
  double* Bt = zero_rec + FM_REC;

  {
    double2* b2 = reinterpret_cast<double2*>(block);
#pragma unroll
    for (int k = 0; k < FM_BLOCK / 64; ++k) b2[k * 32 + lane] = make_double2(0.0, 0.0);
    if (lane < FM_REC) zero_rec[lane] = 0.0;
  }
  Lane L;
  L.gq = lane >> 2;
  L.q = lane & 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    L.wofs[c] = 2 * (c + 4 * ((L.gq >> c) & 1));
    L.rowpos[c] = block_pos(c, L.gq, 0, 0, 0);
    L.colpos[c] = block_pos(c, 2 * L.q, 0, 0, 0);
  }
  // the three B-tile elements this lane fetches per cell: e = lane + 32 j -> (x, y, z, c) of the 3 x 3 x 3 x 3 tile
  int te[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int e = lane + 32 * j;
    te[j] = e < 81 ? ((e % 3) | (((e / 3) % 3) << 2) | (((e / 9) % 3) << 4) | ((e / 27) << 6)) : -1;
  }
  const double f = a.f_beta;

  // cell of this warp in round grp (the cells of one launch are whole planes: owned planes or one ghost plane)
  int cx = 0, cy = 0, zl = 0;
  int32_t bs = 0;
  double tl[3] = {0.0, 0.0, 0.0};
  double pin[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  auto locate = [&](int64_t cell_local) {
    const int pl = (int)(cell_local / g.plane), rem = (int)(cell_local % g.plane);
    cy = rem / g.nx;
    cx = rem % g.nx;
    zl = pl + zl_off;
  };
  auto request_cell = [&](int64_t cell_local) {  // bin boundaries of the 8 octants and the B tile
    bs = lane < 9 ? __ldg(a.bin_start + ((a.bin_cell0 + cell_local) << 3) + lane) : 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (te[j] < 0) continue;
      const int x = wrap1(cx - 1 + (te[j] & 3), g.nx), y = wrap1(cy - 1 + ((te[j] >> 2) & 3), g.ny), z = zl - 1 + ((te[j] >> 4) & 3);
      tl[j] = __ldg(&B[g.vidx(x, y, z, te[j] >> 6)]);
    }
  };
  auto request_particles = [&](int32_t base, int32_t p1) {
    const int32_t i = base + lane;
    if (i < p1) {
#pragma unroll
      for (int k = 0; k < 6; ++k) pin[k] = __ldg(a.p[k] + i);
    }
  };

  int64_t cell_local = (int64_t)blockIdx.x * FM_CELLS + wid;
  if (blockIdx.x < groups && cell_local < a.ncells) {
    locate(cell_local);
    request_cell(cell_local);
    request_particles(__shfl_sync(0xffffffffu, bs, 0), __shfl_sync(0xffffffffu, bs, 8));
  }
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    cell_local = (int64_t)grp * FM_CELLS + wid;
    if (cell_local < a.ncells) {
      const int32_t p0 = __shfl_sync(0xffffffffu, bs, 0), p1 = __shfl_sync(0xffffffffu, bs, 8), b4 = __shfl_sync(0xffffffffu, bs, 4);
      const int32_t bsc = bs;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (te[j] >= 0) Bt[lane + 32 * j] = tl[j];
      // the lower node of the cell as the reference's floor() gives it for every particle binned here
      const double cd[3] = {(double)cx, (double)cy, (double)(zl + g.z0 - a.zshift)};
      const int ci[3] = {cx, cy, zl};

      double acc[NMAT][2], cur[NCUR];
#pragma unroll
      for (int v = 0; v < NMAT; ++v) acc[v][0] = acc[v][1] = 0.0;
#pragma unroll
      for (int v = 0; v < NCUR; ++v) cur[v] = 0.0;
      __syncwarp();  // the B tile is complete

      int oct = 0;
      int32_t oend = __shfl_sync(0xffffffffu, bsc, 1);
      bool zdone = false;
      for (int32_t base = p0; base < p1;) {
        // a round ends at the oz = 0 / oz = 1 boundary when that keeps it within 32 particles: no octant is split
        const int n = (base < b4 && b4 - base <= FM_CHUNK) ? b4 - base : min(FM_CHUNK, p1 - base);
        // ---- records of this round: lane = particle ---------------------------------------------
        if (lane < n) {
          Weights w;
          const double xn[3] = {to_cells(pin[0], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(pin[1], g.dy, g.inv_dy, g.exact_inv & 2),
                                to_cells(pin[2], g.dz, g.inv_dz, g.exact_inv & 4)};
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            // src/impls/ecsim/particles.cpp:76-105 with floor(xn) = cell and floor(xn - 0.5) = cell - 1 + octant bit;
            // xn - cell and xn - 0.5 are exact, so the bit is (xn - cell >= 0.5), the one the key pass binned by
            const double fx = xn[ax] - cd[ax];
            const int o = fx >= 0.5 ? 1 : 0;
            w.in[ax] = ci[ax];
            w.is[ax] = ci[ax] - 1 + o;
            w.wn[ax][1] = fx;
            w.wn[ax][0] = 1 - w.wn[ax][1];
            w.ws[ax][1] = (xn[ax] - 0.5) - (cd[ax] - 1.0 + (double)o);
            w.ws[ax][0] = 1 - w.ws[ax][1];
          }
          const TileIndex t = tile_index<1>(w, cx, cy, zl);
          double Bp[3], b[3];
          gather_B_tile<1>(Bt, w, t, Bp);
          const double v[3] = {pin[3], pin[4], pin[5]};
#pragma unroll
          for (int c = 0; c < 3; ++c) b[c] = Bp[c] * f;
          double vxb[3];
          cross3(v, b, vxb);
          const double vb = dot3(v, b), b2 = dot3(b, b);
          const double cI = a.num_I / (1. + b2);  // q mpw / (1 + b^2)
          double ip[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) ip[c] = cI * (v[c] + vxb[c] + vb * b[c]);
          const double Ap = a.num_A / (1 + b2);   // dt^2/2 mpw q^2 / m / (1 + b^2)
          double al[9];
          al[0] = Ap * (1.0 + b[0] * b[0]);
          al[1] = Ap * (+b[2] + b[0] * b[1]);
          al[2] = Ap * (-b[1] + b[0] * b[2]);
          al[3] = Ap * (-b[2] + b[1] * b[0]);
          al[4] = Ap * (1.0 + b[1] * b[1]);
          al[5] = Ap * (+b[0] + b[1] * b[2]);
          al[6] = Ap * (+b[1] + b[2] * b[0]);
          al[7] = Ap * (-b[0] + b[2] * b[1]);
          al[8] = Ap * (1.0 + b[2] * b[2]);
          double2* r = reinterpret_cast<double2*>(recs + lane * FM_REC);
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            r[ax] = make_double2(w.wn[ax][0], w.ws[ax][0]);
            r[4 + ax] = make_double2(w.wn[ax][1], w.ws[ax][1]);
          }
          r[3] = make_double2(al[0], al[1]);
          r[7] = make_double2(al[2], al[3]);
          r[8] = make_double2(al[4], al[5]);
          r[9] = make_double2(al[6], al[7]);
          r[10] = make_double2(al[8], ip[0]);
          r[11] = make_double2(ip[1], ip[2]);
        }
        __syncwarp();
        if (base + n < p1) request_particles(base + n, p1);  // the next round's particles travel during the MMA phase
        // ---- rank-1 updates, octant segment by octant segment -----------------------------------
        int32_t pos = base;
        const int32_t cend = base + n;
        while (pos < cend) {
          while (oend <= pos) {  // the octant's particles are exhausted (warp-uniform)
            ++oct;
            oend = __shfl_sync(0xffffffffu, bsc, oct + 1);
          }
          if (oct >= 4 && !zdone) {  // first particle with oz = 1: the z-dependent slots change their place
            fold<true, 0>(block, L, acc, cur);
            zdone = true;
          }
          const int32_t seg_end = min(oend, cend);
          const double* r0 = recs + (pos - base) * FM_REC;
          const int cnt = seg_end - pos;
          switch (oct & 3) {
            case 0: octant_segment<0>(r0, zero_rec, cnt, L, acc, cur); break;
            case 1: octant_segment<1>(r0, zero_rec, cnt, L, acc, cur); break;
            case 2: octant_segment<2>(r0, zero_rec, cnt, L, acc, cur); break;
            default: octant_segment<3>(r0, zero_rec, cnt, L, acc, cur); break;
          }
          pos = seg_end;
        }
        __syncwarp();  // the records may be overwritten by the next round
        base += n;
      }
      if (p0 < p1) {
        if (zdone)
          fold<true, 1>(block, L, acc, cur);
        else
          fold<true, 0>(block, L, acc, cur);
        fold<false, 0>(block, L, acc, cur);
      }
    }
    // what the next round's cell needs first is requested before this round's blocks leave
    const int next = grp + gridDim.x;
    const int64_t next_cell = (int64_t)next * FM_CELLS + wid;
    const bool have_next = next < groups && next_cell < a.ncells;
    if (have_next) {
      locate(next_cell);
      request_cell(next_cell);
    }
    __syncthreads();
    if (have_next) request_particles(__shfl_sync(0xffffffffu, bs, 0), __shfl_sync(0xffffffffu, bs, 8));
    // coalesced write-out of the CTA's four blocks, stage[group][entry][cell % 4]: one thread per entry reads it
    // from the four blocks, stores 32 contiguous bytes and leaves zeros behind for the next round
    double* out = stage + ((a.stage_cell0 / CELL_GROUP) + grp) * (int64_t)(BLOCK_ALL * CELL_GROUP);
    for (int e = threadIdx.x; e < BLOCK_ALL; e += FM_THREADS) {
      double v[FM_CELLS];
#pragma unroll
      for (int w = 0; w < FM_CELLS; ++w) {
        v[w] = smem[(size_t)w * FM_CELL + e];
        smem[(size_t)w * FM_CELL + e] = 0.0;
      }
      double2* o2 = reinterpret_cast<double2*>(out + (size_t)staged_entry(e) * CELL_GROUP);
      o2[0] = make_double2(v[0], v[1]);
      o2[1] = make_double2(v[2], v[3]);
    }
    __syncthreads();
  }
}

// =====================================================================================================
// Warp-specialised form (the default).  In the kernel above a warp that carries 94 accumulator registers
// spends three quarters of its time on work that needs none of them (particle loads, B gather, two fp64
// divisions per particle, record stores: all latency-bound), and those registers cap the SM at 12 warps.
// Here a CTA is two warpgroups: warps 0-3 are CONSUMERS (rank-1 updates, folds, write-out; 152 registers
// after setmaxnreg.inc), warps 4-7 are PRODUCERS (bin table, B tile, particle loads, records; 104 registers
// after setmaxnreg.dec).  Producer 4 + i feeds consumer i through a two-stage ring of 32-particle record
// buffers in shared memory, one mbarrier pair (full / empty) per stage; it runs up to two rounds ahead,
// also across cells, so the consumer never sees HBM latency.  Two CTAs per SM: 8 + 8 warps.
// =====================================================================================================
constexpr int WS_THREADS = 256;
constexpr int WS_STAGES = 2;   // ring depth with the cell block in shared memory
constexpr int WT_STAGES = 2;   // ... of the variant-tile form
constexpr int WT_PBUF = 2 * 6 * FM_CHUNK;  // two rounds of particles (SoA) requested ahead with cp.async (and a second B tile)
constexpr int WS_META = 16;  // ints per stage: [0..8] bin boundaries of the cell (first round only), [9] base, [10] n, [11] 1 = first | 2 = last | 4 = stop, [12] group
// per cell: block, WS_STAGES x 32 records, zero record, B tile, meta, 2 x WS_STAGES mbarriers (+ pad to 16 k + 8)
constexpr int WS_CELL = FM_BLOCK + WS_STAGES * FM_CHUNK * FM_REC + FM_REC + FM_TILE + WS_STAGES * WS_META / 2 + 2 * WS_STAGES + 12;
static_assert(WS_CELL % 16 == 8, "write-out reads the four blocks of a CTA without bank conflicts");
// TILES form: no cell block in shared memory (the accumulators go to the staging area as variant tiles, deposit.cuh)
constexpr int WT_CELL = WT_STAGES * FM_CHUNK * FM_REC + FM_REC + 2 * FM_TILE + WT_PBUF + WT_STAGES * WS_META / 2 + 2 * WT_STAGES + 12;
static_assert(WT_CELL % 2 == 0, "16-byte alignment of the records");
template <bool TILES>
__host__ __device__ constexpr int ws_cell() { return TILES ? WT_CELL : WS_CELL; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, int parity)
{
  unsigned ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// a warp that finds its stage not ready backs off before it polls again: the polls of the waiting role share the
// issue slots and the shared-memory pipe with the working role
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity, unsigned backoff_ns)
{
  const unsigned addr = smem_u32(bar);
  while (!mbar_try(addr, parity))
    if (backoff_ns) __nanosleep(backoff_ns);
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// PROF: per-role phase clocks (clock64 deltas summed over the warps of a role) into prof[16]; XPIC_WS_PROF=1 selects it.
//   consumer: 0 wait for a stage, 1 rank-1 updates, 2 folds, 3 barriers, 4 write-out, 5 total
//   producer: 8 hand-out + barrier, 9 cell header (bins, B tile), 10 particle loads -> stage free, 11 records, 12 total
template <bool PROF, bool TILES>
__global__ void __launch_bounds__(WS_THREADS, 2) k_cell_moments_ws(Grid g, DepositArgs a, const double* __restrict__ B, double* __restrict__ stage, int zl_off,
                                                                   int groups, unsigned backoff_ns, int* __restrict__ next_group, unsigned long long* __restrict__ prof)
{
  long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = 0, pt0 = 0;
  auto tick = [&](int k) {
    if (PROF) {
      const long long t = clock64();
      pc[k] += t - pt;
      pt = t;
    }
  };
  if (PROF) pt = pt0 = clock64();
  extern __shared__ __align__(16) double smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = wid & 3;
  const bool producer = wid >= 4;
  constexpr int NST = TILES ? WT_STAGES : WS_STAGES;
  double* block = smem + (size_t)slot * ws_cell<TILES>();
  double* recs = block + (TILES ? 0 : FM_BLOCK);
  double* zero_rec = recs + NST * FM_CHUNK * FM_REC;
  double* Bt = zero_rec + FM_REC;
  double* pbuf = Bt + 2 * FM_TILE;  // TILES only
  int* meta = reinterpret_cast<int*>(TILES ? pbuf + WT_PBUF : Bt + FM_TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(meta + NST * WS_META);
  uint64_t* empty = full + NST;
  // Work is handed out dynamically: producer warp 4 takes the next group of four cells from a global counter and
  // tells the other producers through these two slots (one named barrier per round); the consumers read the group
  // from the first message of the round.  A CTA that starts late -- another kernel held its SM -- simply takes less.
  __shared__ int sched[2];

  if (!producer) {
    if (!TILES) {
      double2* b2 = reinterpret_cast<double2*>(block);
#pragma unroll
      for (int k = 0; k < FM_BLOCK / 64; ++k) b2[k * 32 + lane] = make_double2(0.0, 0.0);
    }
    if (lane < FM_REC) zero_rec[lane] = 0.0;
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < NST; ++st) {
        mbar_init(full + st, 1);
        mbar_init(empty + st, 1);
      }
    }
  }
  __syncthreads();

  if (producer) {
    // ================================ producer: records =============================================
    if (TILES)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    else
      asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    int te[3];  // the three B-tile elements this lane fetches per cell: e = lane + 32 j -> (x, y, z, c)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int e = lane + 32 * j;
      te[j] = e < 81 ? ((e % 3) | (((e / 3) % 3) << 2) | (((e / 9) % 3) << 4) | ((e / 27) << 6)) : -1;
    }
    const double f = a.f_beta;
    int seq = 0;
    if constexpr (TILES) {
      // ---- producer of the variant-tile form: everything it needs from HBM is requested one step ahead with cp.async
      // (no registers are tied up while the data travels).  The bin table and the B tile of the NEXT cell travel while the
      // records of this cell are formed, the particles of the NEXT round (of this cell, or the first of the next cell) while
      // the records of this round are formed.  One commit group per round; wait_group 1 = everything but the newest.
      auto cp8 = [&](double* dst, const double* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      };
      auto round_size = [&](int32_t base, int32_t b4, int32_t p1) {
        // a round ends at the oz = 0 / oz = 1 boundary when that keeps it within 32 particles: no octant is split
        return (base < b4 && b4 - base <= FM_CHUNK) ? b4 - base : min(FM_CHUNK, p1 - base);
      };
      const int ncells = (int)a.ncells, plane = (int)g.plane;  // launch_cell_moments checks the range
      // bin boundaries of the 8 octants (register) and the B tile (into tile buffer tb)
      auto request_header = [&](int cell_local, int tb) {
        const int pl = cell_local / plane, rem = cell_local - pl * plane;
        const int cy = rem / g.nx, cx = rem - cy * g.nx, zl = pl + zl_off;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (te[j] < 0) continue;
          const int x = wrap1(cx - 1 + (te[j] & 3), g.nx), y = wrap1(cy - 1 + ((te[j] >> 2) & 3), g.ny), z = zl - 1 + ((te[j] >> 4) & 3);
          cp8(Bt + tb * FM_TILE + lane + 32 * j, &B[g.vidx(x, y, z, te[j] >> 6)]);
        }
        return lane < 9 ? __ldg(a.bin_start + ((a.bin_cell0 + (int64_t)cell_local) << 3) + lane) : 0;
      };
      // Work is handed out cell by cell from a global counter, to every producer warp on its own (the variant tiles of a
      // cell are stored by its consumer alone: nothing ties the four cells of a CTA together any more).  The counter is
      // read two cells ahead, so that its latency is never waited for.
      auto take = [&]() { return lane == 0 ? atomicAdd(next_group, 1) : 0; };
      auto request_particles = [&](int32_t base, int n, int pb) {
        if (lane < n) {
#pragma unroll
          for (int k = 0; k < 6; ++k) cp8(pbuf + (pb * 6 + k) * FM_CHUNK + lane, a.p[k] + base + lane);
        }
      };
      // first round of the cell with bin table bs: its size; its particles are requested into particle buffer pb
      auto request_first_round = [&](int32_t bs, int pb) {
        const int32_t p0 = __shfl_sync(0xffffffffu, bs, 0), p1 = __shfl_sync(0xffffffffu, bs, 8), b4 = __shfl_sync(0xffffffffu, bs, 4);
        const int n = round_size(p0, b4, p1);
        request_particles(p0, n, pb);
        return n;
      };
      auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };
      int cell = __shfl_sync(0xffffffffu, take(), 0), cell_n = __shfl_sync(0xffffffffu, take(), 0), taken = take();
      int32_t bs = 0, bs_n = 0;
      int n = 0, pb = 0, tb = 0;  // size of the round whose particles are in buffer pb; tile buffer of the current cell
      if (cell < ncells) {
        bs = request_header(cell, tb);
        n = request_first_round(bs, pb);
      }
      commit();
      for (;;) {
        tick(0);
        const bool valid_n = cell_n < ncells;
        if (valid_n) bs_n = request_header(cell_n, tb ^ 1);  // joins the commit group of the next request of particles
        if (cell >= ncells) {
          // no cell left: the stop message
          const int st = seq & (NST - 1);
          mbar_wait(empty + st, ((seq / NST) & 1) ^ 1, backoff_ns);
          int* mt = meta + st * WS_META;
          if (lane == 0) {
            mt[9] = 0;
            mt[10] = 0;
            mt[11] = 1 | 2 | 4;
            mt[12] = cell;
          }
          if (lane < 9) mt[lane] = 0;
          __syncwarp();
          if (lane == 0) mbar_arrive(full + st);
          break;
        }
        {
          const int pl = cell / plane, rem = cell - pl * plane;
          const int cy = rem / g.nx, cx = rem - cy * g.nx, zl = pl + zl_off;
          const int32_t p0 = __shfl_sync(0xffffffffu, bs, 0), p1 = __shfl_sync(0xffffffffu, bs, 8), b4 = __shfl_sync(0xffffffffu, bs, 4);
          const double* Btc = Bt + tb * FM_TILE;
          tick(1);
          // the lower node of the cell as the reference's floor() gives it for every particle binned here
          const double cd[3] = {(double)cx, (double)cy, (double)(zl + g.z0 - a.zshift)};
          const int ci[3] = {cx, cy, zl};
          int32_t base = p0;
          bool first = true;
          do {
            // the particles of the next round: of this cell, or the first round of the next cell
            const int32_t nb = base + n;
            int n_next = 0;
            if (nb < p1) {
              n_next = round_size(nb, b4, p1);
              request_particles(nb, n_next, pb ^ 1);
            }
            else if (valid_n)
              n_next = request_first_round(bs_n, pb ^ 1);
            commit();
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // this round's particles (and this cell's B tile) have arrived
            __syncwarp();                                         // ... those of the other lanes as well (the tile)
            const int st = seq & (NST - 1);
            mbar_wait(empty + st, ((seq / NST) & 1) ^ 1, backoff_ns);  // the consumer has released this stage
            tick(2);
            if (lane < n) {
              double pin[6];
#pragma unroll
              for (int k = 0; k < 6; ++k) pin[k] = pbuf[(pb * 6 + k) * FM_CHUNK + lane];
              Weights w;
              const double xn[3] = {to_cells(pin[0], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(pin[1], g.dy, g.inv_dy, g.exact_inv & 2),
                                    to_cells(pin[2], g.dz, g.inv_dz, g.exact_inv & 4)};
#pragma unroll
              for (int ax = 0; ax < 3; ++ax) {
                // src/impls/ecsim/particles.cpp:76-105 with floor(xn) = cell and floor(xn - 0.5) = cell - 1 + octant bit;
                // xn - cell and xn - 0.5 are exact, so the bit is (xn - cell >= 0.5), the one the key pass binned by
                const double fx = xn[ax] - cd[ax];
                const int o = fx >= 0.5 ? 1 : 0;
                w.in[ax] = ci[ax];
                w.is[ax] = ci[ax] - 1 + o;
                w.wn[ax][1] = fx;
                w.wn[ax][0] = 1 - w.wn[ax][1];
                w.ws[ax][1] = (xn[ax] - 0.5) - (cd[ax] - 1.0 + (double)o);
                w.ws[ax][0] = 1 - w.ws[ax][1];
              }
              const TileIndex t = tile_index<1>(w, cx, cy, zl);
              double Bp[3], b[3];
              gather_B_tile<1>(Btc, w, t, Bp);
              const double v[3] = {pin[3], pin[4], pin[5]};
#pragma unroll
              for (int c = 0; c < 3; ++c) b[c] = Bp[c] * f;
              double vxb[3];
              cross3(v, b, vxb);
              const double vb = dot3(v, b), b2 = dot3(b, b);
              const double cI = a.num_I / (1. + b2);  // q mpw / (1 + b^2)
              double ip[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) ip[c] = cI * (v[c] + vxb[c] + vb * b[c]);
              const double Ap = a.num_A / (1 + b2);   // dt^2/2 mpw q^2 / m / (1 + b^2)
              double2* r = reinterpret_cast<double2*>(recs + (st * FM_CHUNK + lane) * FM_REC);
#pragma unroll
              for (int ax = 0; ax < 3; ++ax) {
                r[ax] = make_double2(w.wn[ax][0], w.ws[ax][0]);
                r[4 + ax] = make_double2(w.wn[ax][1], w.ws[ax][1]);
              }
              r[3] = make_double2(Ap * (1.0 + b[0] * b[0]), Ap * (+b[2] + b[0] * b[1]));
              r[7] = make_double2(Ap * (-b[1] + b[0] * b[2]), Ap * (-b[2] + b[1] * b[0]));
              r[8] = make_double2(Ap * (1.0 + b[1] * b[1]), Ap * (+b[0] + b[1] * b[2]));
              r[9] = make_double2(Ap * (+b[1] + b[2] * b[0]), Ap * (-b[0] + b[2] * b[1]));
              r[10] = make_double2(Ap * (1.0 + b[2] * b[2]), ip[0]);
              r[11] = make_double2(ip[1], ip[2]);
            }
            int* mt = meta + st * WS_META;
            if (first && lane < 9) mt[lane] = bs;
            if (lane == 0) {
              mt[9] = base;
              mt[10] = n;
              mt[11] = (first ? 1 : 0) | (nb >= p1 ? 2 : 0);
              mt[12] = cell;
            }
            __syncwarp();  // every lane's records are written (and its tile gathers done) before lane 0 publishes the stage
            if (lane == 0) mbar_arrive(full + st);
            tick(3);
            base = nb;
            first = false;
            ++seq;
            n = n_next;
            pb ^= 1;
          } while (base < p1);
          tb ^= 1;
        }
        cell = cell_n;
        bs = bs_n;
        cell_n = __shfl_sync(0xffffffffu, taken, 0);
        taken = take();
      }
      if (PROF && lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(prof + 8 + k, (unsigned long long)pc[k]);
        atomicAdd(prof + 12, (unsigned long long)(clock64() - pt0));
      }
      return;
    }
    for (int round = 0;; ++round) {
      if (wid == 4 && lane == 0) sched[round & 1] = atomicAdd(next_group, 1);
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four producer warps
      tick(0);
      const int grp = sched[round & 1];
      const int64_t cell_local = (int64_t)grp * FM_CELLS + slot;
      if (grp >= groups || cell_local >= a.ncells) {
        // nothing to produce for this slot: an empty message keeps the consumer in step (and stops it after the last group)
        const int st = seq & (NST - 1);
        mbar_wait(empty + st, ((seq / NST) & 1) ^ 1, backoff_ns);
        int* mt = meta + st * WS_META;
        if (lane == 0) {
          mt[9] = 0;
          mt[10] = 0;
          mt[11] = 1 | 2 | 8 | (grp >= groups ? 4 : 0);  // 8: no cell behind this slot (nothing is stored)
          mt[12] = grp;
        }
        if (lane < 9) mt[lane] = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(full + st);
        ++seq;
        if (grp >= groups) break;
        continue;
      }
      // the cells of one launch are whole planes (deposit_cells): owned planes or one ghost plane
      const int pl = (int)(cell_local / g.plane), rem = (int)(cell_local % g.plane);
      const int cy = rem / g.nx, cx = rem % g.nx, zl = pl + zl_off;
      const int32_t bs = lane < 9 ? __ldg(a.bin_start + ((a.bin_cell0 + cell_local) << 3) + lane) : 0;  // bin boundaries of the 8 octants
      double tl[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (te[j] < 0) continue;
        const int x = wrap1(cx - 1 + (te[j] & 3), g.nx), y = wrap1(cy - 1 + ((te[j] >> 2) & 3), g.ny), z = zl - 1 + ((te[j] >> 4) & 3);
        tl[j] = __ldg(&B[g.vidx(x, y, z, te[j] >> 6)]);
      }
      const int32_t p0 = __shfl_sync(0xffffffffu, bs, 0), p1 = __shfl_sync(0xffffffffu, bs, 8), b4 = __shfl_sync(0xffffffffu, bs, 4);
      __syncwarp();  // the gathers of the previous cell are done with the tile
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (te[j] >= 0) Bt[lane + 32 * j] = tl[j];
      __syncwarp();
      tick(1);
      // the lower node of the cell as the reference's floor() gives it for every particle binned here
      const double cd[3] = {(double)cx, (double)cy, (double)(zl + g.z0 - a.zshift)};
      const int ci[3] = {cx, cy, zl};
      int32_t base = p0;
      bool first = true;
      do {
        // a round ends at the oz = 0 / oz = 1 boundary when that keeps it within 32 particles: no octant is split
        const int n = (base < b4 && b4 - base <= FM_CHUNK) ? b4 - base : min(FM_CHUNK, p1 - base);
        double pin[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (lane < n) {
#pragma unroll
          for (int k = 0; k < 6; ++k) pin[k] = __ldg(a.p[k] + base + lane);
        }
        const int st = seq & (NST - 1);
        mbar_wait(empty + st, ((seq / NST) & 1) ^ 1, backoff_ns);  // the consumer has released this stage
        tick(2);
        if (lane < n) {
          Weights w;
          const double xn[3] = {to_cells(pin[0], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(pin[1], g.dy, g.inv_dy, g.exact_inv & 2),
                                to_cells(pin[2], g.dz, g.inv_dz, g.exact_inv & 4)};
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            // src/impls/ecsim/particles.cpp:76-105 with floor(xn) = cell and floor(xn - 0.5) = cell - 1 + octant bit;
            // xn - cell and xn - 0.5 are exact, so the bit is (xn - cell >= 0.5), the one the key pass binned by
            const double fx = xn[ax] - cd[ax];
            const int o = fx >= 0.5 ? 1 : 0;
            w.in[ax] = ci[ax];
            w.is[ax] = ci[ax] - 1 + o;
            w.wn[ax][1] = fx;
            w.wn[ax][0] = 1 - w.wn[ax][1];
            w.ws[ax][1] = (xn[ax] - 0.5) - (cd[ax] - 1.0 + (double)o);
            w.ws[ax][0] = 1 - w.ws[ax][1];
          }
          const TileIndex t = tile_index<1>(w, cx, cy, zl);
          double Bp[3], b[3];
          gather_B_tile<1>(Bt, w, t, Bp);
          const double v[3] = {pin[3], pin[4], pin[5]};
#pragma unroll
          for (int c = 0; c < 3; ++c) b[c] = Bp[c] * f;
          double vxb[3];
          cross3(v, b, vxb);
          const double vb = dot3(v, b), b2 = dot3(b, b);
          const double cI = a.num_I / (1. + b2);  // q mpw / (1 + b^2)
          double ip[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) ip[c] = cI * (v[c] + vxb[c] + vb * b[c]);
          const double Ap = a.num_A / (1 + b2);   // dt^2/2 mpw q^2 / m / (1 + b^2)
          double2* r = reinterpret_cast<double2*>(recs + (st * FM_CHUNK + lane) * FM_REC);
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            r[ax] = make_double2(w.wn[ax][0], w.ws[ax][0]);
            r[4 + ax] = make_double2(w.wn[ax][1], w.ws[ax][1]);
          }
          r[3] = make_double2(Ap * (1.0 + b[0] * b[0]), Ap * (+b[2] + b[0] * b[1]));
          r[7] = make_double2(Ap * (-b[1] + b[0] * b[2]), Ap * (-b[2] + b[1] * b[0]));
          r[8] = make_double2(Ap * (1.0 + b[1] * b[1]), Ap * (+b[0] + b[1] * b[2]));
          r[9] = make_double2(Ap * (+b[1] + b[2] * b[0]), Ap * (-b[0] + b[2] * b[1]));
          r[10] = make_double2(Ap * (1.0 + b[2] * b[2]), ip[0]);
          r[11] = make_double2(ip[1], ip[2]);
        }
        int* mt = meta + st * WS_META;
        if (first && lane < 9) mt[lane] = bs;
        if (lane == 0) {
          mt[9] = base;
          mt[10] = n;
          mt[11] = (first ? 1 : 0) | (base + n >= p1 ? 2 : 0);
          mt[12] = grp;
        }
        __syncwarp();  // every lane's records are written before lane 0 publishes the stage
        if (lane == 0) mbar_arrive(full + st);
        tick(3);
        base += n;
        first = false;
        ++seq;
      } while (base < p1);
    }
    if (PROF && lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(prof + 8 + k, (unsigned long long)pc[k]);
      atomicAdd(prof + 12, (unsigned long long)(clock64() - pt0));
    }
    return;
  }

  // ================================== consumer: rank-1 updates, folds, write-out =====================
  if (TILES)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
  else
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
  if constexpr (TILES) {
    // ---- variant tiles: the accumulators are stored as they stand; no cell block, no fold, no barrier between consumers ----
    Lane L;
    L.gq = lane >> 2;
    L.q = lane & 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      L.wofs[c] = 2 * (c + 4 * ((L.gq >> c) & 1));
      L.rowpos[c] = L.colpos[c] = 0;
    }
    int seq = 0;
    for (;;) {
      double acc[NMAT][2], cur[NCUR], curz0 = 0.0;
#pragma unroll
      for (int v = 0; v < NMAT; ++v) acc[v][0] = acc[v][1] = 0.0;
#pragma unroll
      for (int v = 0; v < NCUR; ++v) cur[v] = 0.0;
      int grp = 0;
      int32_t bsc = 0;
      bool zdone = false;
      int flags = 0;  // of the last message: a stop / no-cell message is the first and the last of its "cell"
      // the z-dependent variants (and the Z current) of the lower half of the cell leave; the registers restart at zero
      auto tile0 = [&]() -> double2* {  // staging cell -> [plane][tile][cell of the plane][lane]
        const int sc = (int)a.stage_cell0 + grp, pz = sc / (int)g.plane;  // grp: the cell of the launch; launch_cell_moments checks the range
        return reinterpret_cast<double2*>(stage) + ((int64_t)pz * (NTILE - 1) * g.plane + sc) * 32 + lane;
      };
      auto flush_z = [&]() {
        double2* out2 = tile0();
        const int64_t tstride = g.plane * 32;  // double2 units between two tiles of a staging plane
#pragma unroll
        for (int sl = 0; sl < 9; ++sl) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (!dep(sl, 2) || v >= nvar(sl)) continue;
            const int k = vbase(sl) + v;
            out2[tile_id(sl, v, 0) * tstride] = make_double2(acc[k][0], acc[k][1]);
            acc[k][0] = acc[k][1] = 0.0;
          }
        }
        curz0 = cur[cbase(2)];
        cur[cbase(2)] = 0.0;
      };
      while (true) {
        const int st = seq & (NST - 1);
        mbar_wait(full + st, (seq / NST) & 1, backoff_ns);  // the producer has published this stage
        tick(0);
        const int* mt = meta + st * WS_META;
        const int32_t base = mt[9];
        const int n = mt[10];
        flags = mt[11];
        if (flags & 1) {
          bsc = lane < 9 ? mt[lane] : 0;  // bin boundaries of the cell's octants
          grp = mt[12];
        }
        const double* rbuf = recs + st * FM_CHUNK * FM_REC;
        if (n > 0) {
          Raw raw;
          load_raw(raw, rbuf + min(L.q, n - 1) * FM_REC, L);
          int e0 = 0;  // the groups of octant o cover the records [e0, e1) of the round
          for (int oz = 0; oz < 2; ++oz) {
            if (oz == 1) {
              if (e0 >= n) break;
              if (!zdone) {  // first particle with oz = 1
                tick(1);
                flush_z();
                zdone = true;
                tick(2);
              }
            }
            int32_t ev[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) ev[k] = __shfl_sync(0xffffffffu, bsc, oz * 4 + k + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e1 = min(max(ev[k] - base, 0), n);
              for (int pos = e0; pos < e1; pos += 4) {
                const int nxt = min(pos + 4, e1) + L.q;  // the next group: of this octant, or the first of the next non-empty one
                const double* next_rec = rbuf + min(nxt, n - 1) * FM_REC;
                const bool live = L.q < e1 - pos;
                if (k == 0) group_update<0>(raw, live, next_rec, L, acc, cur);
                if (k == 1) group_update<1>(raw, live, next_rec, L, acc, cur);
                if (k == 2) group_update<2>(raw, live, next_rec, L, acc, cur);
                if (k == 3) group_update<3>(raw, live, next_rec, L, acc, cur);
              }
              e0 = e1;
            }
          }
        }
        __syncwarp();  // every lane is done with the stage before lane 0 releases it
        if (lane == 0) mbar_arrive(empty + st);
        tick(1);
        ++seq;
        if (flags & 2) break;
      }
      if (flags & 4) break;     // stop
      if (flags & 8) continue;  // no cell behind this slot: nothing is stored
      if (!zdone) flush_z();  // no particle in the upper half (or none at all): the oz = 1 tiles below are zeros
      double2* out2 = tile0();
      const int64_t tstride = g.plane * 32;
#pragma unroll
      for (int sl = 0; sl < 9; ++sl) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (v >= nvar(sl)) continue;
          const int k = vbase(sl) + v;
          out2[tile_id(sl, v, 1) * tstride] = make_double2(acc[k][0], acc[k][1]);
        }
      }
      {
        // currents: the four particles of a row are summed; lane (gq, q = c) keeps component c at octant bits 0 and 1
        double v[6] = {cur[0], cur[1], cur[2], cur[3], curz0, cur[4]};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1);
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 2);
        }
        const double2 o = L.q == 0 ? make_double2(v[0], v[1]) : (L.q == 1 ? make_double2(v[2], v[3]) : (L.q == 2 ? make_double2(v[4], v[5]) : make_double2(0.0, 0.0)));
        out2[TILE_CUR * tstride] = o;
      }
      tick(2);
    }
    if (PROF && lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) atomicAdd(prof + k, (unsigned long long)pc[k]);
      atomicAdd(prof + 5, (unsigned long long)(clock64() - pt0));
    }
    return;
  }
  Lane L;
  L.gq = lane >> 2;
  L.q = lane & 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    L.wofs[c] = 2 * (c + 4 * ((L.gq >> c) & 1));
    L.rowpos[c] = block_pos(c, L.gq, 0, 0, 0);
    L.colpos[c] = block_pos(c, 2 * L.q, 0, 0, 0);
  }
  int seq = 0;
  for (;;) {
    int grp = 0;
    bool stop = false;
    {
      double acc[NMAT][2], cur[NCUR];
#pragma unroll
      for (int v = 0; v < NMAT; ++v) acc[v][0] = acc[v][1] = 0.0;
#pragma unroll
      for (int v = 0; v < NCUR; ++v) cur[v] = 0.0;
      int oct = 0;
      int32_t bsc = 0, oend = 0;
      bool zdone = false, any = false;
      while (true) {
        const int st = seq & (NST - 1);
        mbar_wait(full + st, (seq / NST) & 1, backoff_ns);  // the producer has published this stage
        tick(0);
        const int* mt = meta + st * WS_META;
        const int32_t base = mt[9];
        const int n = mt[10], flags = mt[11];
        if (flags & 1) {
          bsc = lane < 9 ? mt[lane] : 0;
          oend = __shfl_sync(0xffffffffu, bsc, 1);
          grp = mt[12];
          stop = (flags & 4) != 0;
        }
        const double* rbuf = recs + st * FM_CHUNK * FM_REC;
        int32_t pos = base;
        const int32_t cend = base + n;
        any = any || n > 0;
        while (pos < cend) {
          while (oend <= pos) {  // the octant's particles are exhausted (warp-uniform)
            ++oct;
            oend = __shfl_sync(0xffffffffu, bsc, oct + 1);
          }
          if (oct >= 4 && !zdone) {  // first particle with oz = 1: the z-dependent slots change their place
            tick(1);
            fold<true, 0>(block, L, acc, cur);
            zdone = true;
            tick(2);
          }
          const int32_t seg_end = min(oend, cend);
          const double* r0 = rbuf + (pos - base) * FM_REC;
          const int cnt = seg_end - pos;
          switch (oct & 3) {
            case 0: octant_segment<0>(r0, zero_rec, cnt, L, acc, cur); break;
            case 1: octant_segment<1>(r0, zero_rec, cnt, L, acc, cur); break;
            case 2: octant_segment<2>(r0, zero_rec, cnt, L, acc, cur); break;
            default: octant_segment<3>(r0, zero_rec, cnt, L, acc, cur); break;
          }
          pos = seg_end;
        }
        __syncwarp();  // every lane is done with the stage before lane 0 releases it
        if (lane == 0) mbar_arrive(empty + st);
        tick(1);
        ++seq;
        if (flags & 2) break;
      }
      if (any) {
        if (zdone)
          fold<true, 1>(block, L, acc, cur);
        else
          fold<true, 0>(block, L, acc, cur);
        fold<false, 0>(block, L, acc, cur);
      }
      tick(2);
    }
    if (stop) break;  // every slot of the CTA receives the stop message in the same round
    consumer_barrier();
    tick(3);
    // coalesced write-out of the CTA's four blocks, stage[group][entry][cell % 4]: one consumer thread per entry
    // reads it from the four blocks, stores 32 contiguous bytes and leaves zeros behind for the next round
    double* out = stage + ((a.stage_cell0 / CELL_GROUP) + grp) * (int64_t)(BLOCK_ALL * CELL_GROUP);
    for (int e = threadIdx.x; e < BLOCK_ALL; e += 128) {
      double v[FM_CELLS];
#pragma unroll
      for (int w = 0; w < FM_CELLS; ++w) {
        v[w] = smem[(size_t)w * WS_CELL + e];
        smem[(size_t)w * WS_CELL + e] = 0.0;
      }
      double2* o2 = reinterpret_cast<double2*>(out + (size_t)staged_entry(e) * CELL_GROUP);
      o2[0] = make_double2(v[0], v[1]);
      o2[1] = make_double2(v[2], v[3]);
    }
    tick(4);
    consumer_barrier();
    tick(3);
  }
  if (PROF && lane == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) atomicAdd(prof + k, (unsigned long long)pc[k]);
    atomicAdd(prof + 5, (unsigned long long)(clock64() - pt0));
  }
}

}  // namespace

int launch_cell_moments(xb_ctx* c, const DepositArgs& a, int zl_off, int form)
{
  // form 0: warp-specialised, variant tiles (the default); 4: warp-specialised, cell blocks folded in shared memory;
  // 3: every warp does everything for its cell (cell blocks)
  const size_t smem = sizeof(double) * FM_CELL * FM_CELLS, smem_ws = sizeof(double) * WS_CELL * FM_CELLS, smem_wt = sizeof(double) * WT_CELL * FM_CELLS;
  if (!c->fused_attr_set) {
    XB_CUDA(cudaFuncSetAttribute(k_cell_moments<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XB_CUDA(cudaFuncSetAttribute((k_cell_moments_ws<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
    XB_CUDA(cudaFuncSetAttribute((k_cell_moments_ws<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
    XB_CUDA(cudaFuncSetAttribute((k_cell_moments_ws<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt));
    XB_CUDA(cudaFuncSetAttribute((k_cell_moments_ws<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt));
    c->fused_attr_set = true;
  }
  const int groups = (int)((a.ncells + FM_CELLS - 1) / FM_CELLS);
  if (a.stage_cell0 + a.ncells + FM_CELLS >= (int64_t)1 << 31) XB_FAIL("deposit: more than 2^31 staging cells");
  if (c->sm_count == 0) XB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
  const int resident = c->sm_count * (form == 3 ? 3 : 2);  // persistent CTAs: one wave
  const int grid = groups < resident ? groups : resident;
  if (grid < 1) return 0;
  if (form == 3) {
    XB_LAUNCH(c, k_cell_moments<3>, grid, FM_THREADS, smem, c->g, a, c->B, c->stage, zl_off, groups);
    return 0;
  }
  // producers (records) and consumers (DMMA) in one CTA, groups handed out dynamically
  const bool tiles = form == 0;
  if (!c->work_counter) XB_CUDA(cudaMalloc(&c->work_counter, sizeof(int)));
  XB_CUDA(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));
  static const bool prof = getenv("XPIC_WS_PROF") != nullptr;  // debugging aid: phase clocks of the two roles on stderr
  unsigned long long* d = nullptr;
  if (prof) {
    XB_CUDA(cudaMalloc(&d, 16 * sizeof(unsigned long long)));
    XB_CUDA(cudaMemsetAsync(d, 0, 16 * sizeof(unsigned long long), c->stream));
  }
  const unsigned backoff = (unsigned)c->ws_backoff_ns;
  if (tiles && prof)
    XB_LAUNCH(c, (k_cell_moments_ws<true, true>), grid, WS_THREADS, smem_wt, c->g, a, c->B, c->stage, zl_off, groups, backoff, c->work_counter, d);
  else if (tiles)
    XB_LAUNCH(c, (k_cell_moments_ws<false, true>), grid, WS_THREADS, smem_wt, c->g, a, c->B, c->stage, zl_off, groups, backoff, c->work_counter, d);
  else if (prof)
    XB_LAUNCH(c, (k_cell_moments_ws<true, false>), grid, WS_THREADS, smem_ws, c->g, a, c->B, c->stage, zl_off, groups, backoff, c->work_counter, d);
  else
    XB_LAUNCH(c, (k_cell_moments_ws<false, false>), grid, WS_THREADS, smem_ws, c->g, a, c->B, c->stage, zl_off, groups, backoff, c->work_counter, d);
  if (prof) {
    unsigned long long h[16];
    XB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    XB_CUDA(cudaStreamSynchronize(c->stream));
    XB_CUDA(cudaFree(d));
    const double nc = 4.0 * grid, cells = (double)a.ncells;
    fprintf(stderr, "k_cell_moments_ws<tiles = %d> phase clocks, cycles per consumer warp (%d CTAs, %.0f cells, %.1f cells per warp):\n", (int)tiles, grid, cells, cells / nc);
    fprintf(stderr, "  consumer: wait %.0f  mma %.0f  fold / store %.0f  barrier %.0f  write-out %.0f  total %.0f\n", h[0] / nc, h[1] / nc, h[2] / nc, h[3] / nc, h[4] / nc, h[5] / nc);
    fprintf(stderr, "  producer: hand-out %.0f  header %.0f  loads+wait %.0f  records %.0f  total %.0f\n", h[8] / nc, h[9] / nc, h[10] / nc, h[11] / nc, h[12] / nc);
  }
  return 0;
}

}  // namespace xb
