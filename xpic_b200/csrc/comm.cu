// comm.cu -- see comm.cuh.
#include <dlfcn.h>

#include <cstdlib>
#include <nccl.h>

#include "comm.cuh"

namespace xb {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi api;

int load_api()
{
  if (api.handle) return 0;
  // a process that already imported torch has its bundled libnccl.so.2 mapped; dlopen by
  // soname returns that copy, otherwise the system one is used.
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) XB_FAIL(std::string("cannot load NCCL: ") + dlerror());
#define XB_SYM(field, name)                                           \
  *(void**)(&api.field) = dlsym(api.handle, name);                    \
  if (!api.field) XB_FAIL(std::string("NCCL symbol missing: ") + name)
  XB_SYM(GetUniqueId, "ncclGetUniqueId");
  XB_SYM(CommInitRank, "ncclCommInitRank");
  XB_SYM(CommDestroy, "ncclCommDestroy");
  XB_SYM(AllReduce, "ncclAllReduce");
  XB_SYM(AllGather, "ncclAllGather");
  XB_SYM(Send, "ncclSend");
  XB_SYM(Recv, "ncclRecv");
  XB_SYM(GroupStart, "ncclGroupStart");
  XB_SYM(GroupEnd, "ncclGroupEnd");
  XB_SYM(GetErrorString, "ncclGetErrorString");
#undef XB_SYM
  return 0;
}

#define XB_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) XB_FAIL(std::string(#call) + " failed: " + api.GetErrorString(r_));     \
  } while (0)
}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int up = 0, down = 0;
  double* recv_buf = nullptr;  // 2 * GZ planes for halo_reduce
};

int comm_unique_id(void* out128)
{
  XB_CHECK(load_api());
  ncclUniqueId id;
  XB_NCCL(api.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init(xb_ctx* c, const void* uid128)
{
  XB_CHECK(load_api());
  if (!uid128) XB_FAIL("xb_create: nranks > 1 needs a communicator id (xb_comm_unique_id)");
  Comm* cm = new Comm();
  ncclUniqueId id;
  memcpy(&id, uid128, 128);
  XB_NCCL(api.CommInitRank(&cm->comm, c->g.nranks, id, c->g.rank));
  cm->up = (c->g.rank + 1) % c->g.nranks;
  cm->down = (c->g.rank - 1 + c->g.nranks) % c->g.nranks;
  XB_CUDA(cudaMalloc(&cm->recv_buf, sizeof(double) * 2 * GZ * c->g.plane * 3));
  c->comm = cm;
  return 0;
}

void comm_free(xb_ctx* c)
{
  if (!c->comm) return;
  if (c->comm->comm) api.CommDestroy(c->comm->comm);
  cudaFree(c->comm->recv_buf);
  delete c->comm;
  c->comm = nullptr;
}

int comm_exchange(xb_ctx* c, const void* to_down, size_t n_to_down, const void* to_up, size_t n_to_up, void* from_up,
                  size_t n_from_up, void* from_down, size_t n_from_down, cudaStream_t stream)
{
  Comm* cm = c->comm;
  if (!stream) stream = c->stream;
  XB_NCCL(api.GroupStart());
  // order matters when up == down (two ranks): first message = low planes -> peer's high side
  if (n_to_down) XB_NCCL(api.Send(to_down, n_to_down, ncclUint8, cm->down, cm->comm, stream));
  if (n_to_up) XB_NCCL(api.Send(to_up, n_to_up, ncclUint8, cm->up, cm->comm, stream));
  if (n_from_up) XB_NCCL(api.Recv(from_up, n_from_up, ncclUint8, cm->up, cm->comm, stream));
  if (n_from_down) XB_NCCL(api.Recv(from_down, n_from_down, ncclUint8, cm->down, cm->comm, stream));
  XB_NCCL(api.GroupEnd());
  return 0;
}

int comm_allgather(xb_ctx* c, void* buf, size_t bytes, cudaStream_t stream)
{
  if (!stream) stream = c->stream;
  XB_NCCL(api.AllGather(static_cast<const char*>(buf) + (size_t)c->g.rank * bytes, buf, bytes, ncclUint8, c->comm->comm, stream));
  return 0;
}

int comm_exchange_list(xb_ctx* c, const ExchangeList& l, cudaStream_t stream)
{
  Comm* cm = c->comm;
  if (!stream) stream = c->stream;
  XB_NCCL(api.GroupStart());
  // same ordering rule as comm_exchange: everything for `down` first, then everything for `up`;
  // receives from `up` first, then from `down` (keeps message order consistent when up == down)
  for (int i = 0; i < l.n; ++i)
    if (l.n_to_down[i]) XB_NCCL(api.Send(l.to_down[i], l.n_to_down[i], ncclUint8, cm->down, cm->comm, stream));
  for (int i = 0; i < l.n; ++i)
    if (l.n_to_up[i]) XB_NCCL(api.Send(l.to_up[i], l.n_to_up[i], ncclUint8, cm->up, cm->comm, stream));
  for (int i = 0; i < l.n; ++i)
    if (l.n_from_up[i]) XB_NCCL(api.Recv(l.from_up[i], l.n_from_up[i], ncclUint8, cm->up, cm->comm, stream));
  for (int i = 0; i < l.n; ++i)
    if (l.n_from_down[i]) XB_NCCL(api.Recv(l.from_down[i], l.n_from_down[i], ncclUint8, cm->down, cm->comm, stream));
  XB_NCCL(api.GroupEnd());
  return 0;
}

// ghost planes [-w, 0) <- down's top planes ; [nzl, nzl + w) <- up's bottom planes
int comm_halo_fill(xb_ctx* c, double* v, int w, cudaStream_t stream)
{
  if (!stream) stream = c->stream;
  const Grid& g = c->g;
  if (g.nzl < w) XB_FAIL("slab thinner than the halo width");
  const int64_t p3 = g.plane * 3;
  const size_t bytes = sizeof(double) * w * p3;
  double* own_lo = v + (int64_t)GZ * p3;
  double* own_hi = v + (int64_t)(GZ + g.nzl - w) * p3;
  double* gh_lo = v + (int64_t)(GZ - w) * p3;
  double* gh_hi = v + (int64_t)(GZ + g.nzl) * p3;
  // open z: nothing crosses the box ends; the ghost planes there hold zeros (DMDA local vectors)
  const bool down = !(g.open_z && g.rank == 0), up = !(g.open_z && g.rank == g.nranks - 1);
  if (!down) XB_CUDA(cudaMemsetAsync(gh_lo, 0, bytes, stream));
  if (!up) XB_CUDA(cudaMemsetAsync(gh_hi, 0, bytes, stream));
  return comm_exchange(c, own_lo, down ? bytes : 0, own_hi, up ? bytes : 0, gh_hi, up ? bytes : 0, gh_lo, down ? bytes : 0, stream);
}

// owned bottom planes += up-neighbour's view of them (its high ghosts go to up's owners) ...
// my low ghosts [-wlo, 0) belong to `down` (its top planes); my high ghosts [nzl, nzl + whi) to `up`.
int comm_halo_reduce(xb_ctx* c, double* v, int wlo, int whi)
{
  const Grid& g = c->g;
  Comm* cm = c->comm;
  if (g.nzl < wlo || g.nzl < whi) XB_FAIL("slab thinner than the halo width");
  const int64_t p3 = g.plane * 3;
  double* gh_lo = v + (int64_t)(GZ - wlo) * p3;
  double* gh_hi = v + (int64_t)(GZ + g.nzl) * p3;
  double* r_from_up = cm->recv_buf;               // up's low ghosts = my top planes [nzl - wlo, nzl)
  double* r_from_down = cm->recv_buf + GZ * p3;   // down's high ghosts = my bottom planes [0, whi)
  // open z: what was deposited outside the box ends is dropped
  const bool down = !(g.open_z && g.rank == 0), up = !(g.open_z && g.rank == g.nranks - 1);
  XB_CHECK(comm_exchange(c, gh_lo, down ? sizeof(double) * wlo * p3 : 0, gh_hi, up ? sizeof(double) * whi * p3 : 0, r_from_up,
                         up ? sizeof(double) * wlo * p3 : 0, r_from_down, down ? sizeof(double) * whi * p3 : 0));
  // fixed order: contribution from below first, then from above
  if (whi && down) XB_CHECK(add_planes(c, v + (int64_t)GZ * p3, r_from_down, whi * p3));
  if (wlo && up) XB_CHECK(add_planes(c, v + (int64_t)(GZ + g.nzl - wlo) * p3, r_from_up, wlo * p3));
  XB_CUDA(cudaMemsetAsync(gh_lo, 0, sizeof(double) * wlo * p3, c->stream));
  XB_CUDA(cudaMemsetAsync(gh_hi, 0, sizeof(double) * whi * p3, c->stream));
  return 0;
}

int comm_allreduce_sum(xb_ctx* c, double* dev, int n)
{
  XB_NCCL(api.AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, c->comm->comm, c->stream));
  return 0;
}

}  // namespace xb
