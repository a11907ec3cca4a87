// deposit.cuh -- declarations shared by the moment-deposition kernels (deposit.cu, moments_fused.cu).
#pragma once
#include "common.cuh"
#include "stencil.cuh"

namespace xb {

// window position of point t = (k, j, i) of component c for a particle of octant (ox, oy, oz)
// src/impls/ecsim/particles.cpp:145-147
__device__ __forceinline__ int block_pos(int c, int t, int ox, int oy, int oz)
{
  const int i = t & 1, j = (t >> 1) & 1, k = t >> 2;
  if (c == 0) return (k * 2 + j) * 3 + (ox + i);
  if (c == 1) return (k * 3 + (oy + j)) * 2 + i;
  return ((oz + k) * 2 + j) * 2 + i;
}

// ---- accumulator bookkeeping (compile time) -----------------------------------------------------
// slot sl = c1 * 3 + c2.  A slot's 8 x 8 tile moves inside the cell's 12 x 12 block with the octant bit of
// its row component (along that component's staggered axis) and of its column component.  Variants over
// (ox, oy) live in registers; oz is sequenced (see the header comment).
__host__ __device__ constexpr bool dep(int sl, int a) { return sl / 3 == a || sl % 3 == a; }
__host__ __device__ constexpr int nvar(int sl) { return (dep(sl, 0) ? 2 : 1) * (dep(sl, 1) ? 2 : 1); }
__host__ __device__ constexpr int vbase(int sl)
{
  int b = 0;
  for (int i = 0; i < sl; ++i) b += nvar(i);
  return b;
}
// register variant of slot sl for a particle of (ox, oy) = (oxy & 1, oxy >> 1)
__host__ __device__ constexpr int vidx(int sl, int oxy)
{
  return vbase(sl) + (dep(sl, 0) ? (oxy & 1) : 0) + (dep(sl, 1) ? (oxy >> 1) * (dep(sl, 0) ? 2 : 1) : 0);
}
// octant bit of axis a that register variant v of slot sl stands for (oz: the sequenced bit)
__host__ __device__ constexpr int vbit(int sl, int v, int a, int oz)
{
  if (a == 2) return oz;
  if (!dep(sl, a)) return 0;
  if (a == 0) return v & 1;
  return dep(sl, 0) ? (v >> 1) : (v & 1);
}
constexpr int NMAT = vbase(8) + nvar(8);  // 21 accumulator pairs
static_assert(NMAT == 21, "21 (slot, ox, oy) variants");
// currents: component c moves with its own octant bit only: X 2 variants, Y 2, Z 1 (sequenced)
__host__ __device__ constexpr int cbase(int c) { return c == 0 ? 0 : (c == 1 ? 2 : 4); }
__host__ __device__ constexpr int cidx(int c, int oxy) { return cbase(c) + (c == 0 ? (oxy & 1) : (c == 1 ? (oxy >> 1) : 0)); }
constexpr int NCUR = 5;


// ---- variant tiles: the staging layout of the production kernel (moments_fused.cu, k_cell_moments_ws<.., true>) ----
// The accumulator pairs leave the registers as they are: one 8 x 8 tile per (slot, ox, oy, oz) variant, element
// (row corner gq, column corners 2 q, 2 q + 1) in lane 4 gq + q as a double2.  Tiles 0..20: the 21 register variants at
// oz = 0 (for a slot without a Z component: its only tile); 21..29: the nine z-dependent variants at oz = 1; 30: the
// currents, lane (gq, q = c) = (sum at octant bit 0, at octant bit 1) of component c at corner gq.  Overlapping
// variants are summed by the row gather (k_gather_tiles), not in shared memory.
// stage[plane][tile][cell of the plane][lane] in double2 units: a consumer warp stores 512 contiguous bytes per tile (whole
// sectors; 16-byte pieces interleaved over the four cells of a group were measured at 0.2 TB/s), the gather threads of a
// warp read 64-byte rows of x-consecutive cells, 512 bytes apart.  A plane is contiguous (the exchange of boundary planes).
__host__ __device__ constexpr int zbase(int sl)  // z-dependent variants below slot sl
{
  int b = 0;
  for (int i = 0; i < sl; ++i) b += dep(i, 2) ? nvar(i) : 0;
  return b;
}
// tile of variant v of slot sl at octant bit oz
__host__ __device__ constexpr int tile_id(int sl, int v, int oz) { return (oz && dep(sl, 2)) ? NMAT + zbase(sl) + v : vbase(sl) + v; }
constexpr int TILE_CUR = NMAT + zbase(9);
constexpr int NTILE = TILE_CUR + 1;
static_assert(NTILE == 31 && NTILE * 64 == STAGE_CELL, "31 tiles of 64 doubles per cell");
// window position of corner t of component c at octant bit o of the component's staggered axis (block_pos, compile time)
__host__ __device__ constexpr int corner_pos(int c, int t, int o)
{
  const int i = t & 1, j = (t >> 1) & 1, k = t >> 2;
  if (c == 0) return (k * 2 + j) * 3 + (o + i);
  if (c == 1) return (k * 3 + (o + j)) * 2 + i;
  return ((o + k) * 2 + j) * 2 + i;
}

// window coordinates (i, j, k) of position p of component c
__host__ __device__ constexpr int pos_i(int c, int p) { return p % win_n(c, 0); }
__host__ __device__ constexpr int pos_j(int c, int p) { return (p / win_n(c, 0)) % win_n(c, 1); }
__host__ __device__ constexpr int pos_k(int c, int p) { return p / (win_n(c, 0) * win_n(c, 1)); }

// Where element (row gq, column t2) of the tile of variant v, octant bit oz, of the component pair (c1, c2) belongs: the
// row's node is the cell + (ox, oy, oz), the coefficient is stencil slot `slot` of that node (k_gather_tiles;
// tests/host/tile_map_check.cu derives the same from the particle's footprint, src/impls/ecsim/particles.cpp:119-171)
struct TileTarget {
  int ox, oy, oz, slot;
};
__host__ __device__ constexpr TileTarget tile_target(int c1, int c2, int v, int oz, int gq, int t2)
{
  const int sl = c1 * 3 + c2;
  const int p1 = corner_pos(c1, gq, vbit(sl, v, c1, oz)), p2 = corner_pos(c2, t2, vbit(sl, v, c2, oz));
  const int o1x = win_lo(c1, 0) + pos_i(c1, p1), o1y = win_lo(c1, 1) + pos_j(c1, p1), o1z = win_lo(c1, 2) + pos_k(c1, p1);
  return TileTarget{o1x, o1y, o1z,
                    coef_slot(c1, c2, win_lo(c2, 0) + pos_i(c2, p2) - o1x, win_lo(c2, 1) + pos_j(c2, p2) - o1y, win_lo(c2, 2) + pos_k(c2, p2) - o1z)};
}

struct DepositArgs {
  const double* p[6];
  const int32_t* bin_start;
  int64_t bin_cell0;    // first cell (in bin space) of this launch
  int64_t ncells;       // cells in this launch
  int64_t stage_cell0;  // staging cell id of the first cell
  int zshift;
  double q, m, mpw;
  // particle-independent factors of src/impls/ecsim/particles.cpp:107-115, rounded as the reference rounds them:
  // beta = B_p * f_beta; A_p = num_A / (1 + beta^2); I_p = num_I / (1 + beta^2) * (...)
  double f_beta, num_A, num_I;
  double* rec;         // per-particle field record, SoA [12][rec_stride]: A_p alpha (9), I_p (3)
  int64_t rec_stride;
};


// moments_fused.cu: the production pass 1 (field records, DMMA cell blocks, staging write in one kernel)
int launch_cell_moments(xb_ctx* c, const DepositArgs& a, int zl_off, int form);

}  // namespace xb
