// gather.cuh -- per-particle CIC weights on the Yee lattice, field gather and the Boris update.
// Arithmetic follows src/impls/ecsim/simulation.cpp:8-118 (weights :17-43, E :51-57, B :107-113)
// and src/algorithms/boris_push.cpp:48-57 operation by operation.
#pragma once
#include "common.cuh"

namespace xb {

struct Weights {
  int in[3], is[3];  // nodal / staggered lower index; z entries are LOCAL plane indices
  double wn[3][2], ws[3][2];
};

__device__ __forceinline__ void cell_and_octant(double xn, int& i, int& o)
{
  i = (int)floor(xn);
  const int is = (int)floor(xn - 0.5);
  o = is - i + 1;  // src/impls/ecsim/particles.cpp:92-94
}

// r / d as the reference computes it; when d is a power of two the product with 1/d is the same
// number bit for bit and avoids the fp64 division sequence
__device__ __forceinline__ double to_cells(double r, double d, double inv_d, int exact)
{
  return exact ? r * inv_d : r / d;
}

__device__ __forceinline__ void axis_weights(double xn, int& in, int& is, double* wn, double* ws)
{
  const double xs = xn - 0.5;
  in = (int)floor(xn);
  is = (int)floor(xs);
  wn[1] = xn - in;
  wn[0] = 1 - wn[1];
  ws[1] = xs - is;
  ws[0] = 1 - ws[1];
}

// zshift: planes added to the z indices (0 for owned particles; -+nz for ghost copies that came
// across the periodic boundary, so their weights stay bit-identical to the owner's)
__device__ __forceinline__ void make_weights(const Grid& g, double px, double py, double pz, int zshift, Weights& w)
{
  axis_weights(to_cells(px, g.dx, g.inv_dx, g.exact_inv & 1), w.in[0], w.is[0], w.wn[0], w.ws[0]);
  axis_weights(to_cells(py, g.dy, g.inv_dy, g.exact_inv & 2), w.in[1], w.is[1], w.wn[1], w.ws[1]);
  axis_weights(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4), w.in[2], w.is[2], w.wn[2], w.ws[2]);
  w.in[2] += zshift - g.z0;
  w.is[2] += zshift - g.z0;
}

// element offsets (component 0) of the 2 + 2 node columns a particle touches per axis:
// [0] = nodal lower/upper, [1] = staggered lower/upper; x and y wrap periodically (indices lie in
// [-1, n]), z is a local plane index into the ghosted array
struct NodeOffsets {
  int x[2][2], y[2][2], z[2][2];
};

__device__ __forceinline__ int wrap1(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

__device__ __forceinline__ void make_offsets(const Grid& g, const Weights& w, NodeOffsets& o)
{
  const int sy = 3 * g.nx, sz = 3 * (int)g.plane;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    o.x[0][i] = 3 * wrap1(w.in[0] + i, g.nx);
    o.x[1][i] = 3 * wrap1(w.is[0] + i, g.nx);
    o.y[0][i] = sy * wrap1(w.in[1] + i, g.ny);
    o.y[1][i] = sy * wrap1(w.is[1] + i, g.ny);
    o.z[0][i] = sz * (w.in[2] + i + GZ);
    o.z[1][i] = sz * (w.is[2] + i + GZ);
  }
}

__device__ __forceinline__ void gather_E(const Grid& g, const double* __restrict__ E, const Weights& w, const NodeOffsets& o, double* Ep)
{
  Ep[0] = Ep[1] = Ep[2] = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double sx = w.wn[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sy = w.wn[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sz = w.ws[2][k] * w.wn[1][j] * w.wn[0][i];
        Ep[0] += __ldg(&E[o.z[0][k] + o.y[0][j] + o.x[1][i] + 0]) * sx;
        Ep[1] += __ldg(&E[o.z[0][k] + o.y[1][j] + o.x[0][i] + 1]) * sy;
        Ep[2] += __ldg(&E[o.z[1][k] + o.y[0][j] + o.x[0][i] + 2]) * sz;
      }
}

__device__ __forceinline__ void gather_B(const Grid& g, const double* __restrict__ B, const Weights& w, const NodeOffsets& o, double* Bp)
{
  Bp[0] = Bp[1] = Bp[2] = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double sx = w.ws[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sy = w.ws[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sz = w.wn[2][k] * w.ws[1][j] * w.ws[0][i];
        Bp[0] += __ldg(&B[o.z[1][k] + o.y[1][j] + o.x[0][i] + 0]) * sx;
        Bp[1] += __ldg(&B[o.z[1][k] + o.y[0][j] + o.x[1][i] + 1]) * sy;
        Bp[2] += __ldg(&B[o.z[0][k] + o.y[1][j] + o.x[1][i] + 2]) * sz;
      }
}

// ---- gathers from a shared-memory field tile --------------------------------------------------
// A CTA that owns NC x-consecutive cells stages the nodes its particles can touch:
// x in [cx0 - 1, cx0 + NC], y in [cy - 1, cy + 1], z in [cz - 1, cz + 1], 3 components.
constexpr int TILE_CELLS = 4;  // cells per CTA of the deposit-side kernels (= staging group)

template <int NC>
struct FieldTile {
  static constexpr int NX = NC + 2;
  static constexpr int SIZE = 3 * 3 * 3 * NX;  // [c][z][y][x]
};
constexpr int FIELD_TILE = FieldTile<TILE_CELLS>::SIZE;

template <int NC>
__device__ __forceinline__ void load_field_tile(const Grid& g, const double* __restrict__ F, int cx0, int cy, int zl, double* __restrict__ T, int tid,
                                                int nthreads)
{
  constexpr int NX = FieldTile<NC>::NX;
  for (int e = tid; e < FieldTile<NC>::SIZE; e += nthreads) {
    const int x = e % NX, y = (e / NX) % 3, z = (e / (NX * 3)) % 3, c = e / (NX * 9);
    T[e] = __ldg(&F[g.vidx(wrapi(cx0 - 1 + x, g.nx), wrap1(cy - 1 + y, g.ny), zl - 1 + z, c)]);
  }
}

struct TileIndex {
  int xn, xs, yn, ys, zn, zs;  // offsets of the particle's lower nodal / staggered node inside the tile
};

template <int NC>
__device__ __forceinline__ TileIndex tile_index(const Weights& w, int cx0, int cy, int zl)
{
  constexpr int NX = FieldTile<NC>::NX;
  TileIndex t;
  t.xn = w.in[0] - (cx0 - 1);
  t.xs = w.is[0] - (cx0 - 1);
  t.yn = (w.in[1] - (cy - 1)) * NX;
  t.ys = (w.is[1] - (cy - 1)) * NX;
  t.zn = (w.in[2] - (zl - 1)) * (3 * NX);
  t.zs = (w.is[2] - (zl - 1)) * (3 * NX);
  return t;
}

template <int NC>
__device__ __forceinline__ void gather_E_tile(const double* __restrict__ T, const Weights& w, const TileIndex& t, double* Ep)
{
  Ep[0] = Ep[1] = Ep[2] = 0.0;
  constexpr int NX = FieldTile<NC>::NX, C = 9 * NX, Y = NX, Z = 3 * NX;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double sx = w.wn[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sy = w.wn[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sz = w.ws[2][k] * w.wn[1][j] * w.wn[0][i];
        Ep[0] += T[0 * C + t.zn + k * Z + t.yn + j * Y + t.xs + i] * sx;
        Ep[1] += T[1 * C + t.zn + k * Z + t.ys + j * Y + t.xn + i] * sy;
        Ep[2] += T[2 * C + t.zs + k * Z + t.yn + j * Y + t.xn + i] * sz;
      }
}

template <int NC>
__device__ __forceinline__ void gather_B_tile(const double* __restrict__ T, const Weights& w, const TileIndex& t, double* Bp)
{
  Bp[0] = Bp[1] = Bp[2] = 0.0;
  constexpr int NX = FieldTile<NC>::NX, C = 9 * NX, Y = NX, Z = 3 * NX;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double sx = w.ws[2][k] * w.ws[1][j] * w.wn[0][i];
        const double sy = w.ws[2][k] * w.wn[1][j] * w.ws[0][i];
        const double sz = w.wn[2][k] * w.ws[1][j] * w.ws[0][i];
        Bp[0] += T[0 * C + t.zs + k * Z + t.ys + j * Y + t.xn + i] * sx;
        Bp[1] += T[1 * C + t.zs + k * Z + t.yn + j * Y + t.xs + i] * sy;
        Bp[2] += T[2 * C + t.zn + k * Z + t.ys + j * Y + t.xs + i] * sz;
      }
}

// The same two gathers as nested linear interpolations (x, then y, then z): 7 multiply-adds and 7 multiplies per
// component instead of 16 multiplies and 8 multiply-adds for the eight explicit weight products -- the second push is
// bound by fp64 issue, not by HBM.  Same weights, same nodes as interpolate_E_s1 / B_s1
// (src/impls/ecsim/simulation.cpp:8-118); the sum is associated differently (agreement at round-off level).
__device__ __forceinline__ double lerp3(const double* __restrict__ T, int base, int Y, int Z, const double* wx, const double* wy, const double* wz)
{
  const double a00 = T[base] * wx[0] + T[base + 1] * wx[1];
  const double a01 = T[base + Y] * wx[0] + T[base + Y + 1] * wx[1];
  const double a10 = T[base + Z] * wx[0] + T[base + Z + 1] * wx[1];
  const double a11 = T[base + Z + Y] * wx[0] + T[base + Z + Y + 1] * wx[1];
  return (a00 * wy[0] + a01 * wy[1]) * wz[0] + (a10 * wy[0] + a11 * wy[1]) * wz[1];
}

template <int NC>
__device__ __forceinline__ void gather_EB_tile_nested(const double* __restrict__ Et, const double* __restrict__ Bt, const Weights& w, const TileIndex& t,
                                                      double* Ep, double* Bp)
{
  constexpr int NX = FieldTile<NC>::NX, C = 9 * NX, Y = NX, Z = 3 * NX;
  // E_x: staggered in x; E_y: in y; E_z: in z.  B_x: nodal in x, staggered in y, z; ...
  Ep[0] = lerp3(Et, 0 * C + t.zn + t.yn + t.xs, Y, Z, w.ws[0], w.wn[1], w.wn[2]);
  Ep[1] = lerp3(Et, 1 * C + t.zn + t.ys + t.xn, Y, Z, w.wn[0], w.ws[1], w.wn[2]);
  Ep[2] = lerp3(Et, 2 * C + t.zs + t.yn + t.xn, Y, Z, w.wn[0], w.wn[1], w.ws[2]);
  Bp[0] = lerp3(Bt, 0 * C + t.zs + t.ys + t.xn, Y, Z, w.wn[0], w.ws[1], w.ws[2]);
  Bp[1] = lerp3(Bt, 1 * C + t.zs + t.yn + t.xs, Y, Z, w.ws[0], w.wn[1], w.ws[2]);
  Bp[2] = lerp3(Bt, 2 * C + t.zn + t.ys + t.xs, Y, Z, w.ws[0], w.ws[1], w.wn[2]);
}

// ---- binning ------------------------------------------------------------------------------------
// periodic wrap of one coordinate, src/interfaces/point.cpp:18-26
__device__ __forceinline__ double wrap_coord(double s, double L)
{
  if (s < 0.0)
    s = L - (0.0 - s);
  else if (s > L)
    s = 0.0 + (s - L);
  // s == L is the same point as 0 (the reference would index cell N there and drop the particle,
  // src/interfaces/particles.cpp:101-104; a measure-zero event we fold back instead)
  if (s >= L) s = 0.0;
  if (s < 0.0) s = 0.0;
  return s;
}

// r + v * dtm followed by the periodic wrap; unfused multiply and add with explicit rounding so that the
// key pass and the scatter pass, which both evaluate it, get the same bits
__device__ __forceinline__ double moved_coord(double r, double v, double dtm, double L)
{
  return wrap_coord(dtm != 0.0 ? __dadd_rn(r, __dmul_rn(v, dtm)) : r, L);
}

// z after the move: periodic wrap, or none when the z boundary is open (Particles::correct_coordinates wraps the
// periodic axes only, src/interfaces/particles.cpp:329-338); the same bits in the key pass and in the scatter
__device__ __forceinline__ double moved_z(const Grid& g, double r, double v, double dtm)
{
  if (g.open_z) return dtm != 0.0 ? __dadd_rn(r, __dmul_rn(v, dtm)) : r;
  return moved_coord(r, v, dtm, g.Lz);
}

// a particle beyond an open z boundary is removed by the re-binning (src/interfaces/particles.cpp:100-103)
__device__ __forceinline__ bool left_the_box(const Grid& g, double pz)
{
  return g.open_z && ((int)floor(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4)) < 0 || (int)floor(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4)) >= g.nz);
}

// bin plane of a (wrapped) z: 1..nzl inside the slab, 0 / nzl + 1 for the neighbour below / above
__device__ __forceinline__ int slab_plane(const Grid& g, double pz)
{
  int iz = (int)floor(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4));
  iz = min(max(iz, 0), g.nz - 1);
  if (g.open_z && (iz < g.z0 || iz >= g.z0 + g.nzl)) return iz < g.z0 ? 0 : g.nzl + 1;  // no wrap-around neighbour
  const int rel = iz - g.z0;
  if (rel >= 0 && rel < g.nzl) return rel + 1;
  const int up = (iz - (g.z0 + g.nzl) + 2 * g.nz) % g.nz;  // planes above the slab top (periodic)
  const int dn = (g.z0 - 1 - iz + 2 * g.nz) % g.nz;        // planes below the slab bottom
  return up <= dn ? g.nzl + 1 : 0;
}

// bin = (cell << 3) | octant, cell over the nzl + 2 bin planes.  The octant bit is taken relative to the
// CLAMPED cell index, so that (cell, octant) always names floor(xn - 0.5) = cell - 1 + octant: when r / d rounds
// up to exactly n (r one ulp below L, d not a power of two) the particle sits at the top of cell n - 1, octant 1.
__device__ __forceinline__ void clamped_cell_and_octant(double xn, int n, int& i, int& o)
{
  i = min(max((int)floor(xn), 0), n - 1);
  o = min(max((int)floor(xn - 0.5) - i + 1, 0), 1);
}

// RemoveParticles::execute (src/commands/remove_particles.cpp:22-39): the particle goes when the corner of its cell lies
// outside the geometry; tallies { particles, kinetic energy 0.5 m v^2 n / Np } (Energy::get_kinetic)
__device__ __forceinline__ bool removed_by_command(const Grid& g, const Geometry& rm, double px, double py, double pz, double vx, double vy, double vz,
                                                   double m_mpw, unsigned long long* __restrict__ tally)
{
  if (rm.kind < 0) return false;
  const int ix = min(max((int)floor(to_cells(px, g.dx, g.inv_dx, g.exact_inv & 1)), 0), g.nx - 1);
  const int iy = min(max((int)floor(to_cells(py, g.dy, g.inv_dy, g.exact_inv & 2)), 0), g.ny - 1);
  const int iz = min(max((int)floor(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4)), 0), g.nz - 1);
  if (within_geometry(rm, ix * g.dx, iy * g.dy, iz * g.dz)) return false;
  atomicAdd(&tally[0], 1ull);
  atomicAdd(reinterpret_cast<double*>(&tally[1]), 0.5 * (m_mpw * ((vx * vx + vy * vy) + vz * vz)));
  return true;
}

__device__ __forceinline__ int32_t particle_key(const Grid& g, double px, double py, double pz, int pl)
{
  int ix, iy, iz, ox, oy, oz;
  clamped_cell_and_octant(to_cells(px, g.dx, g.inv_dx, g.exact_inv & 1), g.nx, ix, ox);
  clamped_cell_and_octant(to_cells(py, g.dy, g.inv_dy, g.exact_inv & 2), g.ny, iy, oy);
  clamped_cell_and_octant(to_cells(pz, g.dz, g.inv_dz, g.exact_inv & 4), g.nz, iz, oz);
  return (int32_t)(((((int64_t)pl * g.ny + iy) * g.nx + ix) << 3) | (oz << 2) | (oy << 1) | ox);
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
  o[0] = +(a[1] * b[2] - a[2] * b[1]);
  o[1] = -(a[0] * b[2] - a[2] * b[0]);
  o[2] = +(a[0] * b[1] - a[1] * b[0]);
}

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ __forceinline__ void boris_update_vEB(double dt, double qm, const double* Ep, const double* Bp, double* v)
{
  const double alpha = dt * qm;
  double a[3], b[3], w[3], bw[3], bbw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a[c] = +alpha * Ep[c];
    b[c] = -alpha * Bp[c];
    w[c] = v[c] + 0.5 * a[c];
  }
  cross3(b, w, bw);
  cross3(b, bw, bbw);
  const double den = 1.0 + 0.25 * dot3(b, b);
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] += a[c] + (bw[c] + 0.5 * bbw[c]) / den;
}

}  // namespace xb
