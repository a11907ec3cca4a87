// spmv.cu -- y = (L + M) x on the fixed-offset stencil layout.
//
// Replaces MatMult on matA = matL + matM inside KSPSolve (src/impls/ecsim/simulation.cpp:197-201,266)
// and the MatDuplicate + MatAXPY(DIFFERENT_NONZERO_PATTERN) rebuild of matA every step: L's 369
// coefficients per cell are streamed once from HBM (coalesced planes coef[k][node]), M's 13-point
// curl-curl stencil is applied matrix-free from the same shared-memory tile of x.
//
// Roofline: HBM.  Algorithmic bytes per cell = 369 * 8 + 24 (x) + 24 (y) = 3000.
#include "common.cuh"
#include "operators.cuh"
#include "stencil.cuh"

namespace xb {

constexpr int HALO = 2;
constexpr int SX = TX + 2 * HALO, SY = TY + 2 * HALO, SZ = TZ + 2 * HALO;

// sum over the fixed offsets of one component pair; every index is a compile-time constant, so the
// loads carry immediate offsets into the shared tile (xt points at this thread's own node)
template <int C1, int C2>
__device__ __forceinline__ double pair_sum(const double* __restrict__ cp, const double* __restrict__ xt)
{
  constexpr DRange rx = drange(C1, C2, 0), ry = drange(C1, C2, 1), rz = drange(C1, C2, 2);
  constexpr int SLICE = rx.n * ry.n;  // 9, 12 or 16 coefficients per z-slice
  const double* c0 = cp + pair_base(C1, C2) * TILE_NODES;
  const double* x0 = xt + ((C2 * SZ + rz.lo) * SY + ry.lo) * SX + rx.lo;
  double a0 = 0.0, a1 = 0.0;  // two chains
  // the z loop stays a real loop: one slice of loads is in flight per thread, the compiler cannot
  // hoist all 369 loads (it spills when it does) and the code stays inside the instruction cache
#pragma unroll 1
  for (int dz = 0; dz < rz.n; ++dz) {
    double cv[SLICE];
#pragma unroll
    for (int i = 0; i < SLICE; ++i) cv[i] = __ldg(c0 + i * TILE_NODES);
#pragma unroll
    for (int dy = 0; dy < ry.n; ++dy)
#pragma unroll
      for (int dx = 0; dx < rx.n; ++dx) {
        const double xv = x0[dy * SX + dx];
        if ((dx + dy) & 1)
          a1 += cv[dy * rx.n + dx] * xv;
        else
          a0 += cv[dy * rx.n + dx] * xv;
      }
    c0 += SLICE * TILE_NODES;
    x0 += SY * SX;
  }
  return a0 + a1;
}

template <int OP>
__global__ void __launch_bounds__(TX * TY * TZ, 3) k_spmv(Grid g, const double* __restrict__ coef, const double* __restrict__ x, double* __restrict__ y,
                                                      int tiles_x, int tiles_y, int bz0)
{
  __shared__ double xs[3][SZ][SY][SX];
  // the launch covers the z-tiles from bz0 on (a multi-rank SpMV sweeps the tiles that need no ghost planes first)
  const int bx = blockIdx.x % tiles_x, by = (blockIdx.x / tiles_x) % tiles_y, bz = bz0 + blockIdx.x / (tiles_x * tiles_y);
  const int64_t tile = ((int64_t)bz * tiles_y + by) * tiles_x + bx;
  const int x0 = bx * TX, y0 = by * TY, z0 = bz * TZ;

  // stage the tile + halo; global reads are contiguous in (x, c)
  constexpr int ROW = 3 * SX;
  for (int i = threadIdx.x; i < SZ * SY * ROW; i += TX * TY * TZ) {
    const int q = i % ROW, j = (i / ROW) % SY, k = i / (ROW * SY);
    const int xx = q / 3, c = q % 3;
    int zl = z0 + k - HALO;
    zl = zl > g.nzl + GZ - 1 ? g.nzl + GZ - 1 : zl;  // partial z tiles: value unused
    const int gx = wrapi(x0 + xx - HALO, g.nx), gy = wrapi(y0 + j - HALO, g.ny);
    xs[c][k][j][xx] = x[g.vidx(gx, gy, zl, c)];
  }
  __syncthreads();

  const int tx = threadIdx.x % TX, ty = (threadIdx.x / TX) % TY, tz = threadIdx.x / (TX * TY);
  const int gx = x0 + tx, gy = y0 + ty, zl = z0 + tz;
  if (gx >= g.nx || gy >= g.ny || zl >= g.nzl) return;

  double acc[3] = {0.0, 0.0, 0.0};
  if (OP & XB_OP_L) {
    const double* cp = coef + tile * (NCOEF * TILE_NODES) + threadIdx.x;
    const double* xt = &xs[0][tz + HALO][ty + HALO][tx + HALO];
    acc[0] = pair_sum<0, 0>(cp, xt) + pair_sum<0, 1>(cp, xt) + pair_sum<0, 2>(cp, xt);
    acc[1] = pair_sum<1, 0>(cp, xt) + pair_sum<1, 1>(cp, xt) + pair_sum<1, 2>(cp, xt);
    acc[2] = pair_sum<2, 0>(cp, xt) + pair_sum<2, 1>(cp, xt) + pair_sum<2, 2>(cp, xt);
  }
  if (OP & XB_OP_M) {
    const double inv_d[3] = {1.0 / g.dx, 1.0 / g.dy, 1.0 / g.dz};
    auto f = [&](int comp, int ox, int oy, int oz) { return xs[comp][tz + HALO + oz][ty + HALO + oy][tx + HALO + ox]; };
    const double h = 0.5 * g.dt * g.dt;
    const bool cut = g.open_z && g.z0 + zl == 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += 2.0 * f(c, 0, 0, 0) + h * curlcurl(c, inv_d, f, cut);
  }
  const int64_t o = g.vidx(gx, gy, zl, 0);
  y[o + 0] = acc[0];
  y[o + 1] = acc[1];
  y[o + 2] = acc[2];
}

static int launch_spmv(xb_ctx* c, int op, const double* x, double* y, int tiles_x, int tiles_y, int bz0, int ntz, cudaStream_t st)
{
  if (ntz <= 0) return 0;
  const Grid& g = c->g;
  const int grid = tiles_x * tiles_y * ntz;
  switch (op) {
    case XB_OP_L: k_spmv<XB_OP_L><<<grid, TX * TY * TZ, 0, st>>>(g, c->coef, x, y, tiles_x, tiles_y, bz0); break;
    case XB_OP_M: k_spmv<XB_OP_M><<<grid, TX * TY * TZ, 0, st>>>(g, c->coef, x, y, tiles_x, tiles_y, bz0); break;
    case XB_OP_A: k_spmv<XB_OP_A><<<grid, TX * TY * TZ, 0, st>>>(g, c->coef, x, y, tiles_x, tiles_y, bz0); break;
    default: XB_FAIL("spmv: unknown operator selector");
  }
  c->launches++;
  XB_CUDA(cudaGetLastError());
  return 0;
}

int spmv(xb_ctx* c, int op, double* x, double* y)
{
  const Grid& g = c->g;
  if ((op & XB_OP_L) && !c->coef_valid) XB_FAIL("spmv: operator L has not been deposited / uploaded");
  const int w = (op & XB_OP_L) ? 2 : 1;
  const int tiles_x = (g.nx + TX - 1) / TX, tiles_y = (g.ny + TY - 1) / TY, tiles_z = (g.nzl + TZ - 1) / TZ;
  const bool prof = (op & XB_OP_L) != 0;
  // z-tile tz covers planes [2 tz, 2 tz + 2) and reads x on [2 tz - 2, 2 tz + 4): tiles 1 .. hi need no ghost plane
  const int hi = (g.nzl - 4) / TZ;
  if (g.nranks > 1 && hi >= 1) {
    XB_CHECK(halo_begin(c, x, w));  // the ghost planes travel while the inner tiles are swept
    if (prof) XB_CHECK(prof_begin(c, XB_FAMILY_SPMV));
    XB_CHECK(launch_spmv(c, op, x, y, tiles_x, tiles_y, 1, hi, c->stream));
    // the edge tiles follow the exchange on ITS stream: they start the moment the ghost planes are there and fill the
    // tail of the inner launch instead of adding two short launches (two more tails) behind it
    XB_CHECK(launch_spmv(c, op, x, y, tiles_x, tiles_y, 0, 1, c->copy_stream));
    XB_CHECK(launch_spmv(c, op, x, y, tiles_x, tiles_y, hi + 1, tiles_z - hi - 1, c->copy_stream));
    XB_CUDA(cudaEventRecord(c->halo_done, c->copy_stream));  // now: exchange and edge tiles done
    XB_CHECK(halo_end(c));  // inside the timed interval: what is left after the inner tiles counts as SpMV time
    if (prof) XB_CHECK(prof_end(c, XB_FAMILY_SPMV));
    return 0;
  }
  XB_CHECK(halo_fill(c, x, w));
  if (prof) XB_CHECK(prof_begin(c, XB_FAMILY_SPMV));
  XB_CHECK(launch_spmv(c, op, x, y, tiles_x, tiles_y, 0, tiles_z, c->stream));
  if (prof) XB_CHECK(prof_end(c, XB_FAMILY_SPMV));
  return 0;
}

}  // namespace xb
