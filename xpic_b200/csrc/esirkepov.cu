// esirkepov.cu -- ECSIMCorr particle stages: half-moves with charge-conserving Esirkepov current
// deposition (2nd-order spline), the Boris update between them and the predicted field work.
//
// Replaces ecsimcorr::Particles::first_push / second_push (src/impls/ecsimcorr/particles.cpp:27-91),
// Shape::setup(old, new) (src/utils/shape.cpp:12-79), spline_of_2nd_order
// (src/interfaces/sort_parameters.cpp:21-30) and EsirkepovDecomposition::process
// (src/algorithms/esirkepov_decomposition.cpp:20-103).
//
// Round-1 form: one thread per particle, window sums kept in registers, contributions added to
// the ghosted current with fp64 global reductions (RED.ADD.F64).  The deposit is therefore not
// order-deterministic yet; DESIGN.md lists the tile-owned replacement as the next step.
#include "common.cuh"
#include "gather.cuh"

namespace xb {

int reduce_finish(xb_ctx* c, int nv, double* host_out);  // fields.cu

__device__ __forceinline__ double spline2(double s)
{
  s = fabs(s);
  if (s <= 0.5) return (0.75 - s * s);
  if (0.5 < s && s < 1.5) return 0.5 * (1.5 - s) * (1.5 - s);
  return 0.0;
}

// J += Esirkepov current of the straight move old -> new (global coordinates, unwrapped new)
__device__ __forceinline__ void esirkepov_deposit(const Grid& g, const double* ro, const double* rn, double alpha, double* __restrict__ J)
{
  constexpr double radius = 1.5;
  constexpr int SHW = 4;
  const double d[3] = {g.dx, g.dy, g.dz};
  int start[3], size[3];
  double So[3][SHW], Sn[3][SHW];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double po = ro[a] / d[a], pn = rn[a] / d[a];
    start[a] = (int)round(fmin(po, pn) - radius);
    size[a] = (int)floor(fmax(po, pn) + radius) + 1 - start[a];
#pragma unroll
    for (int i = 0; i < SHW; ++i) {
      const double gx = (double)(start[a] + i);
      So[a][i] = i < size[a] ? spline2(po - gx) : 0.0;
      Sn[a][i] = i < size[a] ? spline2(pn - gx) : 0.0;
    }
  }
  const double qx = alpha * g.dx, qy = alpha * g.dy, qz = alpha * g.dz;
  double tjy[SHW][SHW];  // running sums along y, indexed [z][x]
  double tjz[SHW][SHW];  // running sums along z, indexed [y][x]
#pragma unroll
  for (int z = 0; z < SHW; ++z) {
    if (z >= size[2]) break;
    const double nZ = Sn[2][z], oZ = So[2][z];
    const int zl = start[2] + z - g.z0;
#pragma unroll
    for (int y = 0; y < SHW; ++y) {
      if (y >= size[1]) break;
      const double nY = Sn[1][y], oY = So[1][y];
      const int gy = wrapi(start[1] + y, g.ny);
      double jx = 0.0;
#pragma unroll
      for (int x = 0; x < SHW; ++x) {
        if (x >= size[0]) break;
        const double nX = Sn[0][x], oX = So[0][x];
        const double wx = -qx * (nX - oX) * (nY * (2.0 * nZ + oZ) + oY * (2.0 * oZ + nZ));
        const double wy = -qy * (nY - oY) * (nX * (2.0 * nZ + oZ) + oX * (2.0 * oZ + nZ));
        const double wz = -qz * (nZ - oZ) * (nY * (2.0 * nX + oX) + oY * (2.0 * oX + nX));
        jx = ((double)(x > 0) * jx) + wx;
        const double jy = tjy[z][x] = (y > 0 ? tjy[z][x] : 0.0) + wy;
        const double jz = tjz[y][x] = (z > 0 ? tjz[y][x] : 0.0) + wz;
        const int64_t o = g.vidx(wrapi(start[0] + x, g.nx), gy, zl, 0);
        if (jx != 0.0) atomicAdd(&J[o + 0], jx);
        if (jy != 0.0) atomicAdd(&J[o + 1], jy);
        if (jz != 0.0) atomicAdd(&J[o + 2], jz);
      }
    }
  }
}

// first_push: r += v dt/2, Esirkepov(old -> new).  Positions are left unwrapped; the re-binning
// pass that follows wraps them (update_cells, src/impls/ecsim/simulation.cpp:183).
__global__ void k_push_first_corr(Grid g, int64_t n, double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                                  const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz, double alpha,
                                  double* __restrict__ J)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ro[3] = {x[i], y[i], z[i]};
    const double h = 0.5 * g.dt;
    const double rn[3] = {ro[0] + vx[i] * h, ro[1] + vy[i] * h, ro[2] + vz[i] * h};
    esirkepov_deposit(g, ro, rn, alpha, J);
    x[i] = rn[0];
    y[i] = rn[1];
    z[i] = rn[2];
  }
}

// second_push: gather, Boris, half move, Esirkepov, predicted work (ecsimcorr/particles.cpp:59-79)
__global__ void __launch_bounds__(RED_THREADS) k_push_second_corr(Grid g, int64_t n, double* __restrict__ x, double* __restrict__ y,
                                                                 double* __restrict__ z, double* __restrict__ vx, double* __restrict__ vy,
                                                                 double* __restrict__ vz, const double* __restrict__ E,
                                                                 const double* __restrict__ B, double qm, double qn_Np, double alpha,
                                                                 double* __restrict__ J, double* __restrict__ partial)
{
  double work = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ro[3] = {x[i], y[i], z[i]};
    const double vo[3] = {vx[i], vy[i], vz[i]};
    Weights w;
    make_weights(g, ro[0], ro[1], ro[2], 0, w);
    NodeOffsets off;
    make_offsets(g, w, off);
    double Ep[3], Bp[3];
    gather_E(g, E, w, off, Ep);
    gather_B(g, B, w, off, Bp);
    double v[3] = {vo[0], vo[1], vo[2]};
    boris_update_vEB(g.dt, qm, Ep, Bp, v);
    const double h = 0.5 * g.dt;
    const double rn[3] = {ro[0] + v[0] * h, ro[1] + v[1] * h, ro[2] + v[2] * h};
    esirkepov_deposit(g, ro, rn, alpha, J);
    const double vs[3] = {vo[0] + v[0], vo[1] + v[1], vo[2] + v[2]};
    work += qn_Np * 0.5 * dot3(vs, Ep);
    x[i] = rn[0];
    y[i] = rn[1];
    z[i] = rn[2];
    vx[i] = v[0];
    vy[i] = v[1];
    vz[i] = v[2];
  }
  __shared__ double sh[RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  work = warp_sum(work);
  if (lane == 0) sh[wid] = work;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < RED_THREADS / 32; ++q) t += sh[q];
    partial[(int64_t)blockIdx.x * RED_MAXV] = t;
  }
}

int push_first_corr(xb_ctx* c, Species& s)
{
  if (s.count == 0) return 0;
  double** p = s.p[s.cur];
  const double alpha = s.q * s.n / s.Np / (6.0 * c->g.dt);  // ecsimcorr/particles.cpp:130
  int64_t blocks = (s.count + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  XB_LAUNCH(c, k_push_first_corr, (int)blocks, 256, 0, c->g, s.count, p[0], p[1], p[2], p[3], p[4], p[5], alpha, s.currJe);
  s.sorted = false;
  return 0;
}

int push_second_corr(xb_ctx* c, Species& s, const double* Eh, const double* B)
{
  double** p = s.p[s.cur];
  const double qn_Np = s.q * s.n / s.Np;
  const double alpha = qn_Np / (6.0 * c->g.dt);
  XB_LAUNCH(c, k_push_second_corr, RED_BLOCKS, RED_THREADS, 0, c->g, s.count, p[0], p[1], p[2], p[3], p[4], p[5], Eh, B, s.q / s.m, qn_Np, alpha,
            s.currJe, c->red_partial);
  XB_CHECK(reduce_finish(c, 1, &s.pred_w));  // includes the all-reduce of ecsimcorr/particles.cpp:85
  s.sorted = false;
  return 0;
}

}  // namespace xb
