// spmv.cu -- y = (L + M) x on the fixed-offset stencil layout.
//
// Replaces MatMult on matA = matL + matM inside KSPSolve (src/impls/ecsim/simulation.cpp:197-201,266)
// and the MatDuplicate + MatAXPY(DIFFERENT_NONZERO_PATTERN) rebuild of matA every step: L's 369
// coefficients per cell are streamed once from HBM (coalesced planes coef[k][node]), M's 13-point
// curl-curl stencil is applied matrix-free from the same shared-memory tile of x.
//
// Roofline: HBM.  Algorithmic bytes per cell = 369 * 8 + 24 (x) + 24 (y) = 3000.
#include "common.cuh"
#include "stencil.cuh"

namespace xb {

constexpr int TX = 32, TY = 4, TZ = 2;
constexpr int HALO = 2;
constexpr int SX = TX + 2 * HALO, SY = TY + 2 * HALO, SZ = TZ + 2 * HALO;

// (curl^- curl^+ f)_c at the tile point, f(comp, ox, oy, oz) reads the shared tile.
// (CC f)_c = - d_a^- d_a^+ f_c - d_b^- d_b^+ f_c + d_a^- d_c^+ f_a + d_b^- d_c^+ f_b,  {a, b} = axes != c
template <class F>
__device__ __forceinline__ double curlcurl(int c, const double* inv_d, F&& f)
{
  double r = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (a == c) continue;
    int ea[3] = {0, 0, 0}, ec[3] = {0, 0, 0};
    ea[a] = 1;
    ec[c] = 1;
    const double lap = (f(c, ea[0], ea[1], ea[2]) - 2.0 * f(c, 0, 0, 0) + f(c, -ea[0], -ea[1], -ea[2])) * (inv_d[a] * inv_d[a]);
    const double mix = ((f(a, ec[0], ec[1], ec[2]) - f(a, 0, 0, 0)) - (f(a, ec[0] - ea[0], ec[1] - ea[1], ec[2] - ea[2]) - f(a, -ea[0], -ea[1], -ea[2]))) *
                       (inv_d[a] * inv_d[c]);
    r += mix - lap;
  }
  return r;
}

template <int OP>
__global__ void __launch_bounds__(TX * TY * TZ) k_spmv(Grid g, const double* __restrict__ coef, const double* __restrict__ x, double* __restrict__ y,
                                                      int tiles_x, int tiles_y)
{
  __shared__ double xs[3][SZ][SY][SX];
  const int bx = blockIdx.x % tiles_x, by = (blockIdx.x / tiles_x) % tiles_y, bz = blockIdx.x / (tiles_x * tiles_y);
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
  const int64_t node = ((int64_t)zl * g.ny + gy) * g.nx + gx;

  double acc[3] = {0.0, 0.0, 0.0};
  if (OP & XB_OP_L) {
    const double* cp = coef + node;
    const int64_t ncl = g.ncl;
#pragma unroll
    for (int c1 = 0; c1 < 3; ++c1)
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2) {
        constexpr int dummy = 0;
        (void)dummy;
        const DRange rx = drange(c1, c2, 0), ry = drange(c1, c2, 1), rz = drange(c1, c2, 2);
#pragma unroll
        for (int dz = 0; dz < rz.n; ++dz)
#pragma unroll
          for (int dy = 0; dy < ry.n; ++dy)
#pragma unroll
            for (int dx = 0; dx < rx.n; ++dx) {
              const int k = pair_base(c1, c2) + (dz * ry.n + dy) * rx.n + dx;
              acc[c1] += __ldg(cp + (int64_t)k * ncl) * xs[c2][tz + HALO + rz.lo + dz][ty + HALO + ry.lo + dy][tx + HALO + rx.lo + dx];
            }
      }
  }
  if (OP & XB_OP_M) {
    const double inv_d[3] = {1.0 / g.dx, 1.0 / g.dy, 1.0 / g.dz};
    auto f = [&](int comp, int ox, int oy, int oz) { return xs[comp][tz + HALO + oz][ty + HALO + oy][tx + HALO + ox]; };
    const double h = 0.5 * g.dt * g.dt;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += 2.0 * f(c, 0, 0, 0) + h * curlcurl(c, inv_d, f);
  }
  const int64_t o = g.vidx(gx, gy, zl, 0);
  y[o + 0] = acc[0];
  y[o + 1] = acc[1];
  y[o + 2] = acc[2];
}

int spmv(xb_ctx* c, int op, double* x, double* y)
{
  const Grid& g = c->g;
  if ((op & XB_OP_L) && !c->coef_valid) XB_FAIL("spmv: operator L has not been deposited / uploaded");
  XB_CHECK(halo_fill(c, x, (op & XB_OP_L) ? 2 : 1));
  const int tiles_x = (g.nx + TX - 1) / TX, tiles_y = (g.ny + TY - 1) / TY, tiles_z = (g.nzl + TZ - 1) / TZ;
  const int grid = tiles_x * tiles_y * tiles_z;
  const bool prof = c->spmv_profile && (op & XB_OP_L);
  if (prof) {
    if (c->spmv_events_used + 2 > c->spmv_events.size())
      for (int i = 0; i < 512; ++i) {
        cudaEvent_t e;
        XB_CUDA(cudaEventCreate(&e));
        c->spmv_events.push_back(e);
      }
    XB_CUDA(cudaEventRecord(c->spmv_events[c->spmv_events_used], c->stream));
  }
  switch (op) {
    case XB_OP_L: XB_LAUNCH(c, k_spmv<XB_OP_L>, grid, TX * TY * TZ, 0, g, c->coef, x, y, tiles_x, tiles_y); break;
    case XB_OP_M: XB_LAUNCH(c, k_spmv<XB_OP_M>, grid, TX * TY * TZ, 0, g, c->coef, x, y, tiles_x, tiles_y); break;
    case XB_OP_A: XB_LAUNCH(c, k_spmv<XB_OP_A>, grid, TX * TY * TZ, 0, g, c->coef, x, y, tiles_x, tiles_y); break;
    default: XB_FAIL("spmv: unknown operator selector");
  }
  if (prof) {
    XB_CUDA(cudaEventRecord(c->spmv_events[c->spmv_events_used + 1], c->stream));
    c->spmv_events_used += 2;
  }
  return 0;
}

}  // namespace xb
