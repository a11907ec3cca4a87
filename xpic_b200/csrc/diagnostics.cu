// diagnostics.cu -- per-step particle diagnostics on the device: the reference walks
// `Particles::storage` on the host every step for these, which would force a particle download per step.
//
// Reference code replaced:
//   ParticlesChargeDensity::Shape::setup / collect   src/diagnostics/charge_conservation.cpp:33-97
//   ChargeConservation::initialize / add_columns     src/diagnostics/charge_conservation.cpp:115-171
//   Divergence (negative Yee shift)                  src/utils/operators.cpp:275-323
// The charge density is a scalar field kept in component 0 of a ghosted grid vector, so the halo
// reduction and the slab layout of the field kernels apply unchanged.
#include "comm.cuh"
#include "common.cuh"
#include "gather.cuh"

namespace xb {

int reduce_finish(xb_ctx* c, int nv, double* host_out);  // fields.cu

__device__ __forceinline__ double spline2_diag(double x)  // interfaces/sort_parameters.cpp:21-30
{
  x = fabs(x);
  if (x <= 0.5) return (0.75 - x * x);
  if (x < 1.5) return 0.5 * (1.5 - x) * (1.5 - x);
  return 0.0;
}

// rho(g) += q * n/Np * S2(x - gx) S2(y - gy) S2(z - gz) over the 3^3 nodes from ceil(p - 1.5)
__global__ void __launch_bounds__(256) k_charge_density(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                       const double* __restrict__ z, double qn_Np, double* __restrict__ rho)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {to_cells(x[i], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(y[i], g.dy, g.inv_dy, g.exact_inv & 2),
                       to_cells(z[i], g.dz, g.inv_dz, g.exact_inv & 4)};
  int start[3];
  double w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    start[a] = (int)ceil(p[a] - 1.5);
#pragma unroll
    for (int j = 0; j < 3; ++j) w[a][j] = spline2_diag(p[a] - (double)(start[a] + j));
  }
#pragma unroll
  for (int kz = 0; kz < 3; ++kz)
#pragma unroll
    for (int jy = 0; jy < 3; ++jy)
#pragma unroll
      for (int ix = 0; ix < 3; ++ix) {
        const double v = qn_Np * (w[0][ix] * w[1][jy] * w[2][kz]);  // q * cache[i] * n_Np, cache = sx * sy * sz
        atomicAdd(&rho[g.vidx(wrapi(start[0] + ix, g.nx), wrapi(start[1] + jy, g.ny), start[2] + kz - g.z0, 0)], v);
      }
}

// diff = (rho_new - rho_old) / dt + div^- J ; partial sums of |diff| and diff^2; optionally sum += (rho_new - rho_old) / dt
__global__ void __launch_bounds__(RED_THREADS) k_continuity(Grid g, const double* __restrict__ rho_new, const double* __restrict__ rho_old,
                                                           const double* __restrict__ J, double* __restrict__ sum, int mode,
                                                           double* __restrict__ partial)
{
  double n1 = 0.0, n2 = 0.0;
  const double ix = 1.0 / g.dx, iy = 1.0 / g.dy, iz = 1.0 / g.dz;
  for (int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; node < g.ncl; node += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(node % g.nx), y = (int)((node / g.nx) % g.ny), zl = (int)(node / g.plane);
    const int64_t o = g.vidx(x, y, zl, 0);
    double diff;
    if (mode == 0) {  // one sort: d rho / dt, accumulated into the running sum of all sorts
      diff = (rho_new[o] - rho_old[o]) / g.dt;  // VecAYPX(diff, -1, rho); VecScale(diff, 1 / dt)
      sum[o] += diff;
    }
    else
      diff = sum[o];  // the total: sum over sorts + div of the total current
    const double div = (ix * J[o + 0] - ix * J[g.vidx(wrapi(x - 1, g.nx), y, zl, 0)]) + (iy * J[o + 1] - iy * J[g.vidx(x, wrapi(y - 1, g.ny), zl, 1)]) +
                       (iz * J[o + 2] - iz * J[g.vidx(x, y, zl - 1, 2)]);
    diff += div;
    n1 += fabs(diff);
    n2 += diff * diff;
  }
  __shared__ double sh[RED_THREADS / 32][2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  n1 = warp_sum(n1);
  n2 = warp_sum(n2);
  if (lane == 0) {
    sh[wid][0] = n1;
    sh[wid][1] = n2;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * RED_MAXV + threadIdx.x] = t;
  }
}

// MomentumConservation::calculate (src/diagnostics/momentum_conservation.cpp:71-126): per particle the node
// weights of the global 2nd-order Shape (Shape::setup, utils/shape.cpp:34-45,81-107: window from
// round(p - 1.5), 3 or 4 nodes per axis; all four are visited here, the extra one carries zero weight):
//   P  += m / Np * v * sum_i ns(i),         ns = No_z No_y No_x
//   QE += q / Np * sum_i E(g_i) . Es(i),    Es = Shape::electric (shape.h:56-63): No No Sh per component
__global__ void __launch_bounds__(RED_THREADS) k_momentum(Grid g, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                         const double* __restrict__ z, const double* __restrict__ vx, const double* __restrict__ vy,
                                                         const double* __restrict__ vz, const double* __restrict__ E, double* __restrict__ partial)
{
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double p[3] = {to_cells(x[i], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(y[i], g.dy, g.inv_dy, g.exact_inv & 2),
                         to_cells(z[i], g.dz, g.inv_dz, g.exact_inv & 4)};
    int start[3];
    double no[3][4], sh[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      start[a] = (int)round(p[a] - 1.5);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double gj = (double)(start[a] + j);
        no[a][j] = spline2_diag(p[a] - gj);
        sh[a][j] = spline2_diag(p[a] - (gj + 0.5));
      }
    }
    double ns = 0.0, e[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int kz = 0; kz < 4; ++kz)
#pragma unroll
      for (int jy = 0; jy < 4; ++jy)
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
          const int64_t o = g.vidx(wrapi(start[0] + ix, g.nx), wrapi(start[1] + jy, g.ny), start[2] + kz - g.z0, 0);
          ns += no[2][kz] * no[1][jy] * no[0][ix];
          e[0] += __ldg(&E[o + 0]) * (no[2][kz] * no[1][jy] * sh[0][ix]);
          e[1] += __ldg(&E[o + 1]) * (no[2][kz] * sh[1][jy] * no[0][ix]);
          e[2] += __ldg(&E[o + 2]) * (sh[2][kz] * no[1][jy] * no[0][ix]);
        }
    acc[0] += vx[i] * ns;
    acc[1] += vy[i] * ns;
    acc[2] += vz[i] * ns;
    acc[3] += e[0];
    acc[4] += e[1];
    acc[5] += e[2];
  }
  __shared__ double shm[RED_THREADS / 32][6];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double t = warp_sum(acc[k]);
    if (lane == 0) shm[wid][k] = t;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += shm[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * RED_MAXV + threadIdx.x] = t;
  }
}

// out = {Px, Py, Pz, QEx, QEy, QEz} of one sort with the present E, summed over all ranks
int momentum(xb_ctx* c, Species& s, double* out)
{
  XB_CHECK(halo_fill(c, c->E, GZ));  // DMGlobalToLocal(E), momentum_conservation.cpp:75-76
  double** p = s.p[s.cur];
  XB_LAUNCH(c, k_momentum, RED_BLOCKS, RED_THREADS, 0, c->g, s.count, p[0], p[1], p[2], p[3], p[4], p[5], c->E, c->red_partial);
  XB_CHECK(reduce_finish(c, 6, out));
  const double Np = (double)s.Np;
  for (int k = 0; k < 3; ++k) {
    out[k] *= s.m / Np;      // :87
    out[3 + k] *= s.q / Np;  // :88
  }
  return 0;
}

// DistributionMoment::collect (src/diagnostics/distribution_moment.cpp:157-210) with the moments of :212-313:
// cell-centred quantities, 1st-order form factor on the two cells per axis from round(p - 1), weight n / Np.
// Only particles whose cell lies in the region contribute (:180-181); a deposit outside the region is dropped
// unless the region spans the whole (periodic) axis (set_local_da, :59-110: DM_BOUNDARY_GHOSTED otherwise).
struct MomentArgs {
  int moment;
  int start[3], size[3];  // region in cells
  double q, m, n_Np;
  double cx, cy;          // box centre for the cylindrical moments
};

__device__ __forceinline__ int moment_values(const MomentArgs& a, double px, double py, const double* v, double* out)
{
  switch (a.moment) {
    case XB_MOMENT_DENSITY: out[0] = 1.0; return 1;
    case XB_MOMENT_CURRENT:
      for (int c = 0; c < 3; ++c) out[c] = a.q * v[c];
      return 3;
    case XB_MOMENT_MOMENTUM_FLUX:
    case XB_MOMENT_MOMENTUM_FLUX_DIAG:
    case XB_MOMENT_MOMENTUM_FLUX_CYL:
    case XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL: {
      double w[3] = {v[0], v[1], v[2]};
      if (a.moment == XB_MOMENT_MOMENTUM_FLUX_CYL || a.moment == XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL) {  // _get_v_cyl, :257-275
        const double x = px - a.cx, y = py - a.cy, r = hypot(x, y);
        if (!isinf(1.0 / r)) {
          w[0] = (+x * v[0] + y * v[1]) / r;
          w[1] = (-y * v[0] + x * v[1]) / r;
        }
      }
      if (a.moment == XB_MOMENT_MOMENTUM_FLUX_DIAG || a.moment == XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL) {
        for (int c = 0; c < 3; ++c) out[c] = a.m * w[c] * w[c];
        return 3;
      }
      out[0] = a.m * w[0] * w[0];
      out[1] = a.m * w[0] * w[1];
      out[2] = a.m * w[0] * w[2];
      out[3] = a.m * w[1] * w[1];
      out[4] = a.m * w[1] * w[2];
      out[5] = a.m * w[2] * w[2];
      return 6;
    }
  }
  return 0;
}

// lo: components 0..2 of the moment, hi: components 3..5 (grid vectors, owned part valid after the halo reduction)
__global__ void __launch_bounds__(256) k_cell_moment(Grid g, MomentArgs a, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                    const double* __restrict__ z, const double* __restrict__ vx, const double* __restrict__ vy,
                                                    const double* __restrict__ vz, double* __restrict__ lo, double* __restrict__ hi)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {to_cells(x[i], g.dx, g.inv_dx, g.exact_inv & 1), to_cells(y[i], g.dy, g.inv_dy, g.exact_inv & 2),
                       to_cells(z[i], g.dz, g.inv_dz, g.exact_inv & 4)};
  const int full[3] = {a.size[0] == g.nx, a.size[1] == g.ny, a.size[2] == g.nz};
  int start[3];
  double w[3][2];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const int cell = (int)floor(p[ax]);
    if (cell < a.start[ax] || cell >= a.start[ax] + a.size[ax]) return;  // is_point_within_bounds(vg, gstart, gsize)
    start[ax] = (int)round(p[ax] - 1.0);  // Shape::make_start(p_r, shr = 1)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double s = fabs(p[ax] - ((double)(start[ax] + j) + 0.5));
      w[ax][j] = s <= 1.0 ? 1.0 - s : 0.0;  // spline_of_1st_order
    }
  }
  const double v[3] = {vx[i], vy[i], vz[i]};
  double mv[6];
  const int ms = moment_values(a, x[i], y[i], v, mv);
#pragma unroll
  for (int kz = 0; kz < 2; ++kz)
#pragma unroll
    for (int jy = 0; jy < 2; ++jy)
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const int c[3] = {start[0] + ix, start[1] + jy, start[2] + kz};
        bool keep = true;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax)
          if (!full[ax] && (c[ax] < a.start[ax] || c[ax] >= a.start[ax] + a.size[ax])) keep = false;
        if (!keep) continue;
        const double si = (w[0][ix] * w[1][jy] * w[2][kz]) * a.n_Np;
        const int64_t o = g.vidx(wrapi(c[0], g.nx), wrapi(c[1], g.ny), c[2] - g.z0, 0);
        for (int j = 0; j < ms; ++j) atomicAdd(j < 3 ? &lo[o + j] : &hi[o + j - 3], mv[j] * si);
      }
}

int moment_size(int moment)
{
  switch (moment) {
    case XB_MOMENT_DENSITY: return 1;
    case XB_MOMENT_CURRENT:
    case XB_MOMENT_MOMENTUM_FLUX_DIAG:
    case XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL: return 3;
    case XB_MOMENT_MOMENTUM_FLUX:
    case XB_MOMENT_MOMENTUM_FLUX_CYL: return 6;
  }
  return 0;
}

// components 0..2 into c->tmp2, 3..5 into c->tmp (owned parts valid after the halo reduction)
int distribution_moment(xb_ctx* c, Species& s, int moment, const int32_t* start, const int32_t* size)
{
  const Grid& g = c->g;
  const int ms = moment_size(moment);
  if (ms == 0) XB_FAIL("distribution_moment: unknown moment");
  MomentArgs a;
  a.moment = moment;
  const int full[3] = {g.nx, g.ny, g.nz};
  for (int ax = 0; ax < 3; ++ax) {
    a.start[ax] = start ? start[ax] : 0;
    a.size[ax] = size ? size[ax] : full[ax];
    if (a.start[ax] < 0 || a.size[ax] < 1 || a.start[ax] + a.size[ax] > full[ax]) XB_FAIL("distribution_moment: the region is not inside the box");
  }
  a.q = s.q;
  a.m = s.m;
  a.n_Np = s.n / (double)s.Np;
  a.cx = 0.5 * g.Lx;
  a.cy = 0.5 * g.Ly;
  XB_CHECK(vec_zero(c, c->tmp2));
  if (ms > 3) XB_CHECK(vec_zero(c, c->tmp));
  if (s.count > 0) {
    double** p = s.p[s.cur];
    XB_LAUNCH(c, k_cell_moment, (int)((s.count + 255) / 256), 256, 0, g, a, s.count, p[0], p[1], p[2], p[3], p[4], p[5], c->tmp2, c->tmp);
  }
  XB_CHECK(halo_reduce(c, c->tmp2, GZ, GZ));
  if (ms > 3) XB_CHECK(halo_reduce(c, c->tmp, GZ, GZ));
  return 0;
}

// ---- VelocityDistribution (src/diagnostics/velocity_distribution.cpp) ------------------------------
struct VelocityArgs {
  int projector;
  Geometry ge;
  int xstart[3], xsize[3];  // axis-aligned box of the geometry in cells (velocity_distribution_builder.cpp:35-69)
  int vstart, vsize;        // the same for both velocity axes, as in set_regions (:53-63)
  double dvx, dvy;
  double cx, cy;            // 0.5 * geom_x, 0.5 * geom_y (get_vr_vphi, :178-194)
};

// one thread per particle: the cell it sits in selects it (centre inside the geometry, :133-145), its projected
// velocity picks the bin (:147-156); counts are integers, the weight n / Np is applied afterwards
__global__ void __launch_bounds__(256) k_velocity_histogram(Grid g, VelocityArgs a, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                           const double* __restrict__ z, const double* __restrict__ vx, const double* __restrict__ vy,
                                                           const double* __restrict__ vz, unsigned long long* __restrict__ hist)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r[3] = {x[i], y[i], z[i]};
  const int cell[3] = {(int)floor(to_cells(r[0], g.dx, g.inv_dx, g.exact_inv & 1)), (int)floor(to_cells(r[1], g.dy, g.inv_dy, g.exact_inv & 2)),
                       (int)floor(to_cells(r[2], g.dz, g.inv_dz, g.exact_inv & 4))};
#pragma unroll
  for (int ax = 0; ax < 3; ++ax)
    if (cell[ax] < a.xstart[ax] || cell[ax] >= a.xstart[ax] + a.xsize[ax]) return;
  if (!within_geometry(a.ge, (cell[0] + 0.5) * g.dx, (cell[1] + 0.5) * g.dy, (cell[2] + 0.5) * g.dz)) return;
  const double v[3] = {vx[i], vy[i], vz[i]};
  double pa, pb;
  if (a.projector == XB_PROJECTOR_VX_VY) {
    pa = v[0];
    pb = v[1];
  }
  else if (a.projector == XB_PROJECTOR_VZ_VXY) {
    pa = v[2];
    pb = sqrt(v[0] * v[0] + v[1] * v[1] + 0.0);  // Vector3R(px, py, 0).length()
  }
  else {
    const double X = r[0] - a.cx, Y = r[1] - a.cy;
    const double rr = hypot(X, Y);
    if (isinf(1.0 / rr)) {  // particles on the axis keep (vx, vy)
      pa = v[0];
      pb = v[1];
    }
    else {
      pa = (+X * v[0] + Y * v[1]) / rr;
      pb = (-Y * v[0] + X * v[1]) / rr;
    }
  }
  const long long ia = (long long)round(pa / a.dvx), ib = (long long)round(pb / a.dvy);  // ROUND_STEP
  if (ia < a.vstart || ia >= (long long)a.vstart + a.vsize || ib < a.vstart || ib >= (long long)a.vstart + a.vsize) return;
  atomicAdd(&hist[(ib - a.vstart) * (long long)a.vsize + (ia - a.vstart)], 1ull);
}

__global__ void __launch_bounds__(256) k_histogram_counts(int64_t n, const unsigned long long* __restrict__ hist, double* __restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)hist[i];  // exact; the weight is applied after the sum over the ranks
}

void velocity_region(const double dv[2], const double vmin[2], const double vmax[2], int32_t* vstart, int32_t* vsize)
{
  // set_regions (velocity_distribution.cpp:53-63) uses vx_min, vx_max and dvx for both axes
  *vstart = (int32_t)std::round(vmin[0] / dv[0]);
  *vsize = (int32_t)std::round((vmax[0] - vmin[0]) / dv[0]);
}

// out: vsize * vsize doubles, [index of the second projection][index of the first], summed over all ranks
int velocity_distribution(xb_ctx* c, Species& s, int projector, const Geometry& ge, const double dv[2], const double vmin[2], const double vmax[2],
                          double* host_out)
{
  const Grid& g = c->g;
  VelocityArgs a;
  a.projector = projector;
  a.ge = ge;
  const double d[3] = {g.dx, g.dy, g.dz};
  for (int ax = 0; ax < 3; ++ax) {  // FLOOR_STEP of the bounding box
    const double lo = ge.kind == XB_GEOMETRY_BOX ? ge.p[ax] : (ax < 2 ? ge.p[ax] - ge.p[3] : ge.p[2] - 0.5 * ge.p[4]);
    const double hi = ge.kind == XB_GEOMETRY_BOX ? ge.p[3 + ax] : (ax < 2 ? ge.p[ax] + ge.p[3] : ge.p[2] + 0.5 * ge.p[4]);
    a.xstart[ax] = (int)std::floor(lo / d[ax]);
    a.xsize[ax] = (int)std::floor(hi / d[ax]) - a.xstart[ax];
  }
  int32_t vstart, vsize;
  velocity_region(dv, vmin, vmax, &vstart, &vsize);
  if (vsize <= 0) XB_FAIL("velocity distribution: empty velocity region");
  a.vstart = vstart;
  a.vsize = vsize;
  a.dvx = dv[0];
  a.dvy = dv[1];
  a.cx = 0.5 * g.Lx;
  a.cy = 0.5 * g.Ly;
  const int64_t bins = (int64_t)vsize * vsize;
  unsigned long long* hist = nullptr;
  double* sum = nullptr;
  XB_CUDA(cudaMalloc(&hist, sizeof(unsigned long long) * bins));
  XB_CUDA(cudaMalloc(&sum, sizeof(double) * bins));
  int rc = 0;
  do {
    if ((rc = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * bins, c->stream) != cudaSuccess)) break;
    if (s.count > 0) {
      double** p = s.p[s.cur];
      k_velocity_histogram<<<(int)((s.count + 255) / 256), 256, 0, c->stream>>>(g, a, s.count, p[0], p[1], p[2], p[3], p[4], p[5], hist);
    }
    k_histogram_counts<<<(int)((bins + 255) / 256), 256, 0, c->stream>>>(bins, hist, sum);
    if ((rc = cudaGetLastError() != cudaSuccess)) break;
    if (g.nranks > 1 && (rc = comm_allreduce_sum(c, sum, (int)bins))) break;  // VecScatter ADD_VALUES over the ranks (:163-164)
    if ((rc = cudaMemcpyAsync(host_out, sum, sizeof(double) * bins, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)) break;
    rc = cudaStreamSynchronize(c->stream) != cudaSuccess;
  } while (false);
  cudaFree(hist);
  cudaFree(sum);
  if (rc) XB_FAIL("velocity distribution: device error");
  const double w = s.n / (double)s.Np;  // n_Np(point)
  for (int64_t i = 0; i < bins; ++i) host_out[i] *= w;
  return 0;
}

static int ensure_rho(xb_ctx* c, Species& s)
{
  for (int b = 0; b < 2; ++b)
    if (!s.rho[b]) {
      XB_CUDA(cudaMalloc(&s.rho[b], sizeof(double) * c->g.ntot));
      XB_CUDA(cudaMemsetAsync(s.rho[b], 0, sizeof(double) * c->g.ntot, c->stream));
    }
  return 0;
}

// ParticlesChargeDensity::collect into s.rho[s.rho_cur]
int charge_density(xb_ctx* c, Species& s)
{
  XB_CHECK(ensure_rho(c, s));
  double* rho = s.rho[s.rho_cur];
  XB_CHECK(vec_zero(c, rho));
  if (s.count > 0) {
    double** p = s.p[s.cur];
    const int blocks = (int)((s.count + 255) / 256);
    XB_LAUNCH(c, k_charge_density, blocks, 256, 0, c->g, s.count, p[0], p[1], p[2], s.q * (s.n / (double)s.Np), rho);
  }
  return halo_reduce(c, rho, GZ, GZ);  // DMLocalToGlobal(ADD), charge_conservation.cpp:95
}

// ChargeConservation::add_columns: norms = {N1dQ_s, N2dQ_s} for every sort, then {N1dQ_tot, N2dQ_tot}
int charge_conservation(xb_ctx* c, int which_current, double* norms)
{
  const Grid& g = c->g;
  if (which_current == 1) XB_CHECK(cap_alloc(c));  // before the first eccapfim step J = 0, as in the reference
  double* total_J = which_current == 0 ? c->currJe : c->cap_J;
  XB_CHECK(vec_zero(c, c->tmp2));  // running sum over the sorts
  size_t k = 0;
  for (auto& s : c->sorts) {
    XB_CHECK(ensure_rho(c, s));
    s.rho_cur ^= 1;  // the previous collection becomes rho_old
    XB_CHECK(charge_density(c, s));
    double* J = which_current == 0 ? s.currJe : s.currI;  // eccapfim keeps Particles::J in currI
    XB_CHECK(halo_fill(c, J, 1));
    XB_LAUNCH(c, k_continuity, RED_BLOCKS, RED_THREADS, 0, g, s.rho[s.rho_cur], s.rho[s.rho_cur ^ 1], J, c->tmp2, 0, c->red_partial);
    double r[2];
    XB_CHECK(reduce_finish(c, 2, r));
    norms[2 * k + 0] = r[0];
    norms[2 * k + 1] = std::sqrt(r[1]);
    ++k;
  }
  XB_CHECK(halo_fill(c, total_J, 1));
  XB_LAUNCH(c, k_continuity, RED_BLOCKS, RED_THREADS, 0, g, nullptr, nullptr, total_J, c->tmp2, 1, c->red_partial);
  double r[2];
  XB_CHECK(reduce_finish(c, 2, r));
  norms[2 * k + 0] = r[0];
  norms[2 * k + 1] = std::sqrt(r[1]);
  return 0;
}

}  // namespace xb
