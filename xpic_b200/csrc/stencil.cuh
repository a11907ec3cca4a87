// stencil.cuh -- the fixed-offset layout of the implicit Maxwell operator's particle part L.
//
// L[(c1, r), (c2, r + d)] exists for the offsets d a single particle can couple
// (src/impls/ecsim/particles.cpp:119-171): per axis a
//   a != c1, a != c2 : both weights nodal            d in {-1, 0, 1}
//   a == c1 == c2    : both staggered                d in {-1, 0, 1}
//   a == c1 != c2    : row staggered, column nodal   d in {-1, 0, 1, 2}
//   a == c2 != c1    : row nodal, column staggered   d in {-2, -1, 0, 1}
// which gives 27 slots for c1 == c2 and 48 otherwise: 123 per row, 369 per cell.
// (The reference's COO pattern also holds 18 explicit zeros per row, |d| = 2 along the
// staggered axis of a diagonal pair; they never receive a value and are not stored here.)
#pragma once

namespace xb {

constexpr int NCOEF = 369;

struct DRange {
  int lo, n;
};

__host__ __device__ constexpr DRange drange(int c1, int c2, int a)
{
  if (a == c1 && a == c2) return {-1, 3};
  if (a == c1) return {-1, 4};
  if (a == c2) return {-2, 4};
  return {-1, 3};
}

__host__ __device__ constexpr int pair_size(int c1, int c2)
{
  return drange(c1, c2, 0).n * drange(c1, c2, 1).n * drange(c1, c2, 2).n;
}

__host__ __device__ constexpr int pair_base(int c1, int c2)
{
  int b = 0;
  for (int p = 0; p < c1 * 3 + c2; ++p) b += pair_size(p / 3, p % 3);
  return b;
}

// slot of (c1, c2, dx, dy, dz); the caller guarantees the offset is inside the ranges
__host__ __device__ constexpr int coef_slot(int c1, int c2, int dx, int dy, int dz)
{
  const DRange rx = drange(c1, c2, 0), ry = drange(c1, c2, 1), rz = drange(c1, c2, 2);
  return pair_base(c1, c2) + ((dz - rz.lo) * ry.n + (dy - ry.lo)) * rx.n + (dx - rx.lo);
}

__host__ __device__ constexpr bool in_range(int c1, int c2, int dx, int dy, int dz)
{
  const DRange rx = drange(c1, c2, 0), ry = drange(c1, c2, 1), rz = drange(c1, c2, 2);
  return dx >= rx.lo && dx < rx.lo + rx.n && dy >= ry.lo && dy < ry.lo + ry.n && dz >= rz.lo && dz < rz.lo + rz.n;
}

static_assert(pair_base(2, 2) + pair_size(2, 2) == NCOEF, "stencil layout must have 369 slots");

// ---- cell-local window of component c (src/impls/ecsim/simulation.cpp:405-411) ---------------
// 3 wide from -1 along the staggered axis, 2 wide from 0 elsewhere; 12 positions per component.
__host__ __device__ constexpr int win_lo(int c, int a) { return a == c ? -1 : 0; }
__host__ __device__ constexpr int win_n(int c, int a) { return a == c ? 3 : 2; }
// position index inside the 12-point window, x fastest (src/impls/ecsim/simulation.cpp:448-449)
__host__ __device__ constexpr int win_index(int c, int i, int j, int k)
{
  return (k * win_n(c, 1) + j) * win_n(c, 0) + i;
}

// ---- blocked storage of the coefficients ------------------------------------------------------
// coef[tile][k][t]: a tile is TX x TY x TZ nodes (the SpMV CTA), t = (tz * TY + ty) * TX + tx.
// One tile's 369 x 256 coefficients are 756 KB contiguous: every load of a thread is one base
// register + an immediate offset, and a CTA touches one 2 MB page instead of 369.
constexpr int TX = 32, TY = 4, TZ = 2;
constexpr int TILE_NODES = TX * TY * TZ;

struct TileMap {
  int tiles_x, tiles_y, tiles_z;
  __host__ __device__ inline long long ntiles() const { return (long long)tiles_x * tiles_y * tiles_z; }
  // offset of slot 0 of node (x, y, zl); slot k lives TILE_NODES * k doubles further
  __host__ __device__ inline long long node_offset(int x, int y, int zl) const
  {
    const long long tile = ((long long)(zl / TZ) * tiles_y + (y / TY)) * tiles_x + (x / TX);
    const int t = ((zl % TZ) * TY + (y % TY)) * TX + (x % TX);
    return tile * (long long)(NCOEF * TILE_NODES) + t;
  }
};

__host__ __device__ inline TileMap make_tilemap(int nx, int ny, int nzl)
{
  return TileMap{(nx + TX - 1) / TX, (ny + TY - 1) / TY, (nzl + TZ - 1) / TZ};
}

constexpr int BLOCK_MAT = 9 * 144;  // 1296 mass-matrix entries of one cell (ecsim/simulation.cpp:488)
constexpr int BLOCK_CUR = 36;       // 3 x 12 current partials of one cell
constexpr int BLOCK_ALL = BLOCK_MAT + BLOCK_CUR;
constexpr int STAGE_CELL = 31 * 64;  // doubles of staging per cell in the variant-tile layout (deposit.cuh); the staging area is sized for it
constexpr int CELL_GROUP = 4;       // cells per CTA and per staging group: stage[group][entry][cell % 4] (32-byte runs)

}  // namespace xb
