// fields.cu -- grid-vector kernels: ghost planes, Yee curls, right-hand side, field update,
// BLAS-1 and deterministic reductions.  Every kernel here is HBM-bound streaming work.
//
// Reference call sites replaced (Appendix C of SURVEY.md):
//   DMGlobalToLocal / DMLocalToGlobal      -> halo_fill / halo_reduce
//   Rotor::create_positive/negative + MatMultAdd (utils/operators.cpp:155-215) -> curl kernels
//   advance_fields' right-hand side (ecsim/simulation.cpp:255-270)            -> build_rhs
//   final_update (ecsim/simulation.cpp:241-253)                                -> final_update
//   VecDot / VecNorm / VecAXPY / VecMAXPY / VecScale                           -> dots / axpy_multi / ...
#include "common.cuh"
#include "comm.cuh"

namespace xb {

// ---------------------------------------------------------------------------------------------
// ghost planes
// ---------------------------------------------------------------------------------------------
__global__ void k_halo_fill_local(Grid g, double* __restrict__ v, int width)
{
  const int64_t p3 = g.plane * 3;
  const int64_t total = 2 * (int64_t)width * p3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int which = (int)(i / p3);
    const int64_t e = i % p3;
    const int zl = which < width ? -(which + 1) : g.nzl + (which - width);
    const int src = wrapi(zl, g.nzl);
    // open z: the ghost planes lie outside the box, where DMDA's local vectors hold zeros
    v[(int64_t)(zl + GZ) * p3 + e] = g.open_z ? 0.0 : v[(int64_t)(src + GZ) * p3 + e];
  }
}

// single rank: owned[wrap(zl)] += ghost[zl], ghost planes visited in a fixed order
__global__ void k_halo_reduce_local(Grid g, double* __restrict__ v, int wlo, int whi)
{
  const int64_t p3 = g.plane * 3;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < p3; e += (int64_t)gridDim.x * blockDim.x) {
    for (int k = 1; k <= wlo; ++k) {
      const int zl = -k, dst = wrapi(zl, g.nzl);
      if (!g.open_z) v[(int64_t)(dst + GZ) * p3 + e] += v[(int64_t)(zl + GZ) * p3 + e];  // open z: deposits outside the box are dropped
      v[(int64_t)(zl + GZ) * p3 + e] = 0.0;
    }
    for (int k = 0; k < whi; ++k) {
      const int zl = g.nzl + k, dst = wrapi(zl, g.nzl);
      if (!g.open_z) v[(int64_t)(dst + GZ) * p3 + e] += v[(int64_t)(zl + GZ) * p3 + e];
      v[(int64_t)(zl + GZ) * p3 + e] = 0.0;
    }
  }
}

// multi rank: owned planes += received neighbour ghost planes
__global__ void k_add_planes(double* __restrict__ dst, const double* __restrict__ src, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

static inline int grid_for(int64_t n, int threads = 256)
{
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int halo_fill(xb_ctx* c, double* v, int width)
{
  const Grid& g = c->g;
  if (width > GZ) XB_FAIL("halo_fill: width exceeds GZ");
  if (g.nranks == 1) {
    XB_LAUNCH(c, k_halo_fill_local, grid_for(2 * width * g.plane * 3), 256, 0, g, v, width);
    return 0;
  }
  return comm_halo_fill(c, v, width);
}

// Multi-rank: the exchange of v's ghost planes starts on the copy stream as soon as everything queued so far on the
// main stream is done (the kernel that produced v's edge planes), halo_end() makes the main stream wait for it.  Work
// that needs no ghost planes (the interior planes of a stencil sweep) is launched in between and hides the exchange.
int halo_begin(xb_ctx* c, double* v, int width)
{
  if (c->g.nranks == 1) return halo_fill(c, v, width);
  if (c->halo_pending) XB_FAIL("halo_begin: the previous exchange has not been waited for");
  if (!c->halo_ready) {
    XB_CUDA(cudaEventCreateWithFlags(&c->halo_ready, cudaEventDisableTiming));
    XB_CUDA(cudaEventCreateWithFlags(&c->halo_done, cudaEventDisableTiming));
  }
  XB_CUDA(cudaEventRecord(c->halo_ready, c->stream));
  XB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->halo_ready, 0));
  XB_CHECK(comm_halo_fill(c, v, width, c->copy_stream));
  XB_CUDA(cudaEventRecord(c->halo_done, c->copy_stream));
  c->halo_pending = true;
  return 0;
}

int halo_end(xb_ctx* c)
{
  if (!c->halo_pending) return 0;
  XB_CUDA(cudaStreamWaitEvent(c->stream, c->halo_done, 0));
  c->halo_pending = false;
  return 0;
}

int halo_reduce(xb_ctx* c, double* v, int wlo, int whi)
{
  const Grid& g = c->g;
  if (wlo > GZ || whi > GZ) XB_FAIL("halo_reduce: width exceeds GZ");
  if (g.nranks == 1) {
    XB_LAUNCH(c, k_halo_reduce_local, grid_for(g.plane * 3), 256, 0, g, v, wlo, whi);
    return 0;
  }
  return comm_halo_reduce(c, v, wlo, whi);
}

int add_planes(xb_ctx* c, double* dst, const double* src, int64_t n)
{
  XB_LAUNCH(c, k_add_planes, grid_for(n), 256, 0, dst, src, n);
  return 0;
}

int vec_zero(xb_ctx* c, double* v)
{
  XB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * c->g.ntot, c->stream));
  return 0;
}

int vec_copy_owned(xb_ctx* c, const double* src, double* dst)
{
  XB_CUDA(cudaMemcpyAsync(dst + c->g.own0, src + c->g.own0, sizeof(double) * c->g.nown, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int upload_owned(xb_ctx* c, const double* host, double* dev)
{
  XB_CUDA(cudaMemcpyAsync(dev + c->g.own0, host, sizeof(double) * c->g.nown, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

int download_owned(xb_ctx* c, const double* dev, double* host)
{
  XB_CUDA(cudaMemcpyAsync(host, dev + c->g.own0, sizeof(double) * c->g.nown, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Yee curls (utils/operators.cpp:158-160 values, :175-213 stencils).  One thread per owned node.
// F fetches component comp at (x + ox, y + oy, zl + oz) with x/y wrapped.
// ---------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void curl_node(const Grid& g, bool positive, F&& f, double& cx, double& cy, double& cz)
{
  const double ix = 1.0 / g.dx, iy = 1.0 / g.dy, iz = 1.0 / g.dz;
  if (positive) {
    cx = (+iy * f(0, 1, 0, 2) - iy * f(0, 0, 0, 2)) + (-iz * f(0, 0, 1, 1) + iz * f(0, 0, 0, 1));
    cy = (-ix * f(1, 0, 0, 2) + ix * f(0, 0, 0, 2)) + (+iz * f(0, 0, 1, 0) - iz * f(0, 0, 0, 0));
    cz = (+ix * f(1, 0, 0, 1) - ix * f(0, 0, 0, 1)) + (-iy * f(0, 1, 0, 0) + iy * f(0, 0, 0, 0));
  }
  else {
    cx = (+iy * f(0, 0, 0, 2) - iy * f(0, -1, 0, 2)) + (-iz * f(0, 0, 0, 1) + iz * f(0, 0, -1, 1));
    cy = (-ix * f(0, 0, 0, 2) + ix * f(-1, 0, 0, 2)) + (+iz * f(0, 0, 0, 0) - iz * f(0, 0, -1, 0));
    cz = (+ix * f(0, 0, 0, 1) - ix * f(-1, 0, 0, 1)) + (-iy * f(0, 0, 0, 0) + iy * f(0, -1, 0, 0));
  }
  const double sg = (double)g.curl_sign;
  cx *= sg;
  cy *= sg;
  cz *= sg;
}

#define XB_NODE_LOOP(g, node, x, y, zl)                                                                      \
  for (int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; node < (g).ncl;                       \
       node += (int64_t)gridDim.x * blockDim.x)                                                              \
    if (int x = (int)(node % (g).nx), y = (int)((node / (g).nx) % (g).ny), zl = (int)(node / (g).plane); true)

__global__ void k_curl(Grid g, int positive, const double* __restrict__ f, double* __restrict__ out, double scale, int accumulate)
{
  XB_NODE_LOOP(g, node, x, y, zl)
  {
    auto fetch = [&](int ox, int oy, int oz, int comp) {
      return f[g.vidx(wrapi(x + ox, g.nx), wrapi(y + oy, g.ny), zl + oz, comp)];
    };
    double cx, cy, cz;
    curl_node(g, positive != 0, fetch, cx, cy, cz);
    const int64_t o = g.vidx(x, y, zl, 0);
    if (accumulate) {
      out[o + 0] += scale * cx;
      out[o + 1] += scale * cy;
      out[o + 2] += scale * cz;
    }
    else {
      out[o + 0] = scale * cx;
      out[o + 1] = scale * cy;
      out[o + 2] = scale * cz;
    }
  }
}

int curl_apply(xb_ctx* c, bool positive, const double* f, double* out, double scale, bool accumulate)
{
  XB_LAUNCH(c, k_curl, grid_for(c->g.ncl), 256, 0, c->g, positive ? 1 : 0, f, out, scale, accumulate ? 1 : 0);
  return 0;
}

// rhs = 2 E^n - dt * curr + dt * curl^-(B - B0)     (ecsim/simulation.cpp:260-264)
__global__ void k_build_rhs(Grid g, const double* __restrict__ E, const double* __restrict__ B, const double* __restrict__ B0,
                            const double* __restrict__ curr, double* __restrict__ rhs)
{
  XB_NODE_LOOP(g, node, x, y, zl)
  {
    auto fetch = [&](int ox, int oy, int oz, int comp) {
      const int64_t i = g.vidx(wrapi(x + ox, g.nx), wrapi(y + oy, g.ny), zl + oz, comp);
      return B[i] - B0[i];
    };
    double cx, cy, cz;
    curl_node(g, false, fetch, cx, cy, cz);
    const int64_t o = g.vidx(x, y, zl, 0);
    rhs[o + 0] = (2.0 * E[o + 0] + (-g.dt) * curr[o + 0]) + g.dt * cx;
    rhs[o + 1] = (2.0 * E[o + 1] + (-g.dt) * curr[o + 1]) + g.dt * cy;
    rhs[o + 2] = (2.0 * E[o + 2] + (-g.dt) * curr[o + 2]) + g.dt * cz;
  }
}

int build_rhs(xb_ctx* c, const double* curr, double* rhs)
{
  XB_CHECK(halo_fill(c, c->B, 1));
  XB_CHECK(halo_fill(c, c->B0, 1));
  XB_LAUNCH(c, k_build_rhs, grid_for(c->g.ncl), 256, 0, c->g, c->E, c->B, c->B0, curr, rhs);
  return 0;
}

// E^{n+1} = 2 E^{n+1/2} - E^n ;  B^{n+1} = B^n - dt curl^+ E^{n+1/2}   (ecsim/simulation.cpp:247-248)
__global__ void k_final_update(Grid g, const double* __restrict__ Eh, const double* E, const double* B, double* Eout, double* Bout)
{
  XB_NODE_LOOP(g, node, x, y, zl)
  {
    auto fetch = [&](int ox, int oy, int oz, int comp) {
      return Eh[g.vidx(wrapi(x + ox, g.nx), wrapi(y + oy, g.ny), zl + oz, comp)];
    };
    double cx, cy, cz;
    curl_node(g, true, fetch, cx, cy, cz);
    const int64_t o = g.vidx(x, y, zl, 0);
    Eout[o + 0] = 2.0 * Eh[o + 0] + (-1.0) * E[o + 0];
    Eout[o + 1] = 2.0 * Eh[o + 1] + (-1.0) * E[o + 1];
    Eout[o + 2] = 2.0 * Eh[o + 2] + (-1.0) * E[o + 2];
    Bout[o + 0] = B[o + 0] + (-g.dt) * cx;
    Bout[o + 1] = B[o + 1] + (-g.dt) * cy;
    Bout[o + 2] = B[o + 2] + (-g.dt) * cz;
  }
}

// E^{n+1}, B^{n+1} into other vectors (E and B themselves stay as they are): xb_step_host sends the new fields home
// while the second push still reads B^n
int final_update_into(xb_ctx* c, const double* Ehalf, double* Eout, double* Bout)
{
  XB_CHECK(halo_fill(c, const_cast<double*>(Ehalf), 1));
  XB_LAUNCH(c, k_final_update, grid_for(c->g.ncl), 256, 0, c->g, Ehalf, c->E, c->B, Eout, Bout);
  return 0;
}

int final_update(xb_ctx* c, const double* Ehalf) { return final_update_into(c, Ehalf, c->E, c->B); }

// ---------------------------------------------------------------------------------------------
// FieldsDamping (src/commands/fields_damping.cpp:16-112): a node whose cell centre lies outside the geometry has E and
// B - B0 multiplied by the damping factor of DampForBox (:66-89) / DampForCylinder (:91-112)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double damping_factor(const Grid& g, const Geometry& ge, double coef, const double* r)
{
  if (ge.kind == XB_GEOMETRY_BOX) {
    const double L[3] = {g.Lx, g.Ly, g.Lz};
    double damping = 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double width, delta;
      if (r[i] > ge.p[3 + i]) {
        width = L[i] - ge.p[3 + i];
        delta = r[i] - ge.p[3 + i];
      }
      else if (r[i] < ge.p[i]) {
        width = ge.p[i] - 0;
        delta = r[i] - 0;
      }
      else
        continue;
      const double t = delta / width - 1.0;
      damping *= 1.0 - coef * (t * t);
    }
    return damping;
  }
  const double rr = hypot(r[0] - ge.p[0], r[1] - ge.p[1]);
  if (rr < ge.p[3]) return 1.0;
  const double width = ge.p[0] - ge.p[3], delta = rr - ge.p[3];
  const double delta0 = width * (1.0 + 1.0 / sqrt(coef));
  if (!(delta < delta0)) return 0.0;  // avoids negative factors (:104-109)
  const double t = delta / width - 1.0;
  return 1.0 - coef * (t * t);
}

__global__ void __launch_bounds__(RED_THREADS) k_fields_damping(Grid g, Geometry ge, double coef, double* __restrict__ E, double* __restrict__ B,
                                                               const double* __restrict__ B0, double* __restrict__ partial)
{
  double taken = 0.0;
  XB_NODE_LOOP(g, node, x, y, zl)
  {
    const double r[3] = {(x + 0.5) * g.dx, (y + 0.5) * g.dy, (g.z0 + zl + 0.5) * g.dz};
    if (within_geometry(ge, r[0], r[1], r[2])) continue;
    const double d = damping_factor(g, ge, coef, r);
    const int64_t o = g.vidx(x, y, zl, 0);
    double e2 = 0.0, b2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double e = E[o + c], b = B[o + c] - B0[o + c];
      e2 += e * e;
      b2 += b * b;
      E[o + c] = e * d;
      B[o + c] = b * d + B0[o + c];
    }
    taken += (0.5 * e2) * (1.0 - d * d) + (0.5 * b2) * (1.0 - d * d);  // Energy::get_field(f) * (1 - damping^2)
  }
  __shared__ double sh[RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  taken = warp_sum(taken);
  if (lane == 0) sh[wid] = taken;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q];
    partial[(int64_t)blockIdx.x * RED_MAXV] = t;
  }
}

int reduce_finish(xb_ctx* c, int nv, double* host_out);

int fields_damping(xb_ctx* c, const Geometry& ge, double coefficient, double* damped_energy)
{
  XB_LAUNCH(c, k_fields_damping, RED_BLOCKS, RED_THREADS, 0, c->g, ge, coefficient, c->E, c->B, c->B0, c->red_partial);
  double e = 0.0;
  XB_CHECK(reduce_finish(c, 1, &e));
  if (damped_energy) *damped_energy = e;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Reductions: fixed grid, fixed tree => bit-reproducible run to run.
// ---------------------------------------------------------------------------------------------
struct VecPtrs {
  const double* v[8];
};

template <int NV>
__global__ void __launch_bounds__(RED_THREADS) k_dots(VecPtrs vs, const double* __restrict__ w, int64_t n, double* __restrict__ partial, int slot0)
{
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += vs.v[k][i] * wi;
  }
  __shared__ double sh[RED_THREADS / 32][NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) sh[wid][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) s += sh[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * RED_MAXV + slot0 + threadIdx.x] = s;
  }
}

// one warp per result: lane-strided fixed-order sum over the block partials
__global__ void k_reduce_final(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out)
{
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wid >= nv) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partial[(int64_t)b * RED_MAXV + wid];
  s = warp_sum(s);
  if (lane == 0) out[wid] = s;
}

template <int NV>
static int launch_dots(xb_ctx* c, const double* const* vs, const double* w, int slot0)
{
  VecPtrs p;
  for (int i = 0; i < 8; ++i) p.v[i] = i < NV ? vs[i] + c->g.own0 : nullptr;
  XB_LAUNCH(c, k_dots<NV>, RED_BLOCKS, RED_THREADS, 0, p, w + c->g.own0, c->g.nown, c->red_partial, slot0);
  return 0;
}

int reduce_finish(xb_ctx* c, int nv, double* host_out)
{
  XB_LAUNCH(c, k_reduce_final, 1, 32 * RED_MAXV, 0, c->red_partial, RED_BLOCKS, nv, c->red_out);
  if (c->g.nranks > 1) XB_CHECK(comm_allreduce_sum(c, c->red_out, nv));
  XB_CUDA(cudaMemcpyAsync(c->red_host, c->red_out, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
  XB_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nv; ++i) host_out[i] = c->red_host[i];
  return 0;
}

// sum of every component and of the squares over the owned part (Energy::calculate_energy's VecNorm and
// VecStrideSumAll, src/diagnostics/energy.cpp:43-59), all ranks
__global__ void __launch_bounds__(RED_THREADS) k_field_sums(const double* __restrict__ v, int64_t nodes, double* __restrict__ partial)
{
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nodes; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
    a[0] += x;
    a[1] += y;
    a[2] += z;
    a[3] += (x * x + y * y) + z * z;
  }
  __shared__ double sh[RED_THREADS / 32][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double t = warp_sum(a[k]);
    if (lane == 0) sh[wid][k] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * RED_MAXV + threadIdx.x] = t;
  }
}

int field_sums(xb_ctx* c, const double* v, double* out4)
{
  XB_LAUNCH(c, k_field_sums, RED_BLOCKS, RED_THREADS, 0, v + c->g.own0, c->g.ncl, c->red_partial);
  return reduce_finish(c, 4, out4);
}

int dots(xb_ctx* c, int nv, const double* const* vs, const double* w, double* host_out)
{
  if (nv > RED_MAXV) XB_FAIL("dots: too many vectors");
  for (int s = 0; s < nv; s += 8) {
    const int m = nv - s < 8 ? nv - s : 8;
    switch (m) {
      case 1: XB_CHECK(launch_dots<1>(c, vs + s, w, s)); break;
      case 2: XB_CHECK(launch_dots<2>(c, vs + s, w, s)); break;
      case 3: XB_CHECK(launch_dots<3>(c, vs + s, w, s)); break;
      case 4: XB_CHECK(launch_dots<4>(c, vs + s, w, s)); break;
      case 5: XB_CHECK(launch_dots<5>(c, vs + s, w, s)); break;
      case 6: XB_CHECK(launch_dots<6>(c, vs + s, w, s)); break;
      case 7: XB_CHECK(launch_dots<7>(c, vs + s, w, s)); break;
      default: XB_CHECK(launch_dots<8>(c, vs + s, w, s)); break;
    }
  }
  return reduce_finish(c, nv, host_out);
}

struct Coefs {
  double a[8];
};

// w = alpha * (w + sum_k a_k v_k)
template <int NV>
__global__ void k_axpy_multi(VecPtrs vs, Coefs cf, double* __restrict__ w, int64_t n, double alpha)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = w[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) s += cf.a[k] * vs.v[k][i];
    w[i] = alpha * s;
  }
}

template <int NV>
static int launch_axpy(xb_ctx* c, const double* const* vs, const double* cf, double* w, double alpha)
{
  VecPtrs p;
  Coefs a;
  for (int i = 0; i < 8; ++i) {
    p.v[i] = i < NV ? vs[i] + c->g.own0 : nullptr;
    a.a[i] = i < NV ? cf[i] : 0.0;
  }
  XB_LAUNCH(c, k_axpy_multi<NV>, grid_for(c->g.nown), 256, 0, p, a, w + c->g.own0, c->g.nown, alpha);
  return 0;
}

// w = alpha * (w + sum coef_i vs_i): the projection and the normalisation of a Gram-Schmidt step in one pass
int axpy_multi_scaled(xb_ctx* c, int nv, const double* const* vs, const double* cf, double* w, double alpha)
{
  for (int s = 0; s < nv; s += 8) {
    const int m = nv - s < 8 ? nv - s : 8;
    const double al = s + 8 >= nv ? alpha : 1.0;  // the scale rides on the last chunk
    switch (m) {
      case 1: XB_CHECK(launch_axpy<1>(c, vs + s, cf + s, w, al)); break;
      case 2: XB_CHECK(launch_axpy<2>(c, vs + s, cf + s, w, al)); break;
      case 3: XB_CHECK(launch_axpy<3>(c, vs + s, cf + s, w, al)); break;
      case 4: XB_CHECK(launch_axpy<4>(c, vs + s, cf + s, w, al)); break;
      case 5: XB_CHECK(launch_axpy<5>(c, vs + s, cf + s, w, al)); break;
      case 6: XB_CHECK(launch_axpy<6>(c, vs + s, cf + s, w, al)); break;
      case 7: XB_CHECK(launch_axpy<7>(c, vs + s, cf + s, w, al)); break;
      default: XB_CHECK(launch_axpy<8>(c, vs + s, cf + s, w, al)); break;
    }
  }
  return 0;
}

int axpy_multi(xb_ctx* c, int nv, const double* const* vs, const double* cf, double* w) { return axpy_multi_scaled(c, nv, vs, cf, w, 1.0); }

__global__ void k_scale_into(const double* __restrict__ w, double alpha, double* __restrict__ out, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = alpha * w[i];
}

int scale_into(xb_ctx* c, const double* w, double alpha, double* out)
{
  XB_LAUNCH(c, k_scale_into, grid_for(c->g.nown), 256, 0, w + c->g.own0, alpha, out + c->g.own0, c->g.nown);
  return 0;
}

__global__ void k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}

int axpby(xb_ctx* c, double a, const double* x, double b, double* y)
{
  XB_LAUNCH(c, k_axpby, grid_for(c->g.nown), 256, 0, a, x + c->g.own0, b, y + c->g.own0, c->g.nown);
  return 0;
}

}  // namespace xb
