// comm.cuh -- z-slab neighbour exchange and global sums over NCCL (NVLink 5 / NVSwitch).
// Replaces the MPI traffic hidden in DMGlobalToLocal / DMLocalToGlobal / VecDot / MatMult and the
// explicit particle migration of src/interfaces/particles.cpp:118-248 (only the +-z neighbours
// exist in a slab layout).  NCCL is loaded with dlopen on first use, so a single-GPU run does
// not need it at all.
#pragma once
#include "common.cuh"

namespace xb {

int comm_unique_id(void* out128);
int comm_init(xb_ctx* c, const void* uid128);
void comm_free(xb_ctx* c);
int comm_halo_fill(xb_ctx* c, double* v, int width, cudaStream_t stream = nullptr);
int comm_halo_reduce(xb_ctx* c, double* v, int wlo, int whi);
int comm_allreduce_sum(xb_ctx* c, double* dev, int n);
// Exchange byte buffers with the z neighbours: `to_down`/`to_up` are sent, `from_up`/`from_down`
// received (sizes in bytes, known on both sides).
int comm_exchange(xb_ctx* c, const void* to_down, size_t n_to_down, const void* to_up, size_t n_to_up, void* from_up,
                  size_t n_from_up, void* from_down, size_t n_from_down, cudaStream_t stream = nullptr);
// every rank contributes `bytes` bytes at buf + rank * bytes and receives all ranks' contributions (in place)
int comm_allgather(xb_ctx* c, void* buf, size_t bytes, cudaStream_t stream = nullptr);
// The same for lists of buffers (particle SoA segments), all inside one NCCL group.
struct ExchangeList {
  int n = 0;
  const void* to_down[16];
  const void* to_up[16];
  void* from_up[16];
  void* from_down[16];
  size_t n_to_down[16], n_to_up[16], n_from_up[16], n_from_down[16];
};
int comm_exchange_list(xb_ctx* c, const ExchangeList& l, cudaStream_t stream = nullptr);
int add_planes(xb_ctx* c, double* dst, const double* src, int64_t n);  // fields.cu

}  // namespace xb
