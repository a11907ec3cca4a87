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
int launch_cell_moments(xb_ctx* c, const DepositArgs& a, int zl_off, int occupancy);

}  // namespace xb
