// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// NCCL is dlopen'ed on first use so that single-GPU users (and the symbol-export test on a CPU box)
// never need it; in a torch process the already-loaded libnccl.so.2 is reused.
//
// The reference never communicates state itself: `dss!` is a user closure and norms/dots are local
// (src/functions.jl:99, newtons_method.jl:6-10).  These entry points are what a distributed state type
// plugs into those hooks: a slab halo exchange for `dss!` and in-stream all-reduces for the Krylov /
// Newton reductions.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "cts_internal.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string load_error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.load_error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return;
    }
#define LOAD(field, sym)                                             \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) api.load_error = std::string("missing NCCL symbol ") + sym;
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  });
  return &api;
}

#define CTS_NCCL(ctx, api, expr)                                                                          \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess)                                                                                \
      return cts_fail((ctx), CTS_ENCCL, std::string(#expr) + ": " + ((api)->GetErrorString ? (api)->GetErrorString(_r) : "?")); \
  } while (0)

}  // namespace

extern "C" {

int cts_comm_get_unique_id(void* out) {
  if (!out) return CTS_EINVAL;
  NcclApi* api = nccl_api();
  if (!api->load_error.empty()) return CTS_ENCCL;
  static_assert(sizeof(ncclUniqueId) == CTS_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return CTS_ENCCL;
  memcpy(out, &id, sizeof(id));
  return CTS_OK;
}

int cts_comm_init(cts_ctx* ctx, int nranks, int rank, const void* unique_id) {
  CTS_CHECK_CTX(ctx);
  if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) return cts_fail(ctx, CTS_EINVAL, "cts_comm_init: bad argument");
  if (ctx->nccl_comm) return cts_fail(ctx, CTS_EINVAL, "communicator already attached");
  NcclApi* api = nccl_api();
  if (!api->load_error.empty()) return cts_fail(ctx, CTS_ENCCL, api->load_error);
  CTS_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  CTS_NCCL(ctx, api, api->CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return CTS_OK;
}

int cts_comm_destroy(cts_ctx* ctx) {
  CTS_CHECK_CTX(ctx);
  if (!ctx->nccl_comm) return CTS_OK;
  NcclApi* api = nccl_api();
  cudaStreamSynchronize(ctx->stream);
  api->CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->rank = 0;
  ctx->nranks = 1;
  return CTS_OK;
}

int cts_comm_rank(cts_ctx* ctx, int* rank, int* nranks) {
  CTS_CHECK_CTX(ctx);
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return CTS_OK;
}

int cts_halo_exchange(cts_ctx* ctx, cts_dtype ft, size_t count, const void* send_lo, const void* send_hi, void* recv_lo,
                      void* recv_hi) {
  CTS_CHECK_CTX(ctx);
  return cts_halo_exchange_on(ctx, ctx->stream, ft, count, send_lo, send_hi, recv_lo, recv_hi);
}

int cts_allreduce_sum_f64(cts_ctx* ctx, double* dev_buf, size_t count) {
  CTS_CHECK_CTX(ctx);
  if (ctx->nranks <= 1) return CTS_OK;
  NcclApi* api = nccl_api();
  CTS_NCCL(ctx, api, api->AllReduce(dev_buf, dev_buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
  ctx->launches++;
  return CTS_OK;
}

int cts_allreduce_max_f64_host(cts_ctx* ctx, double* host_val) {
  CTS_CHECK_CTX(ctx);
  if (!host_val) return CTS_EINVAL;
  if (ctx->nranks <= 1) return CTS_OK;
  NcclApi* api = nccl_api();
  ctx->h_result[0] = *host_val;
  CTS_CUDA(ctx, cudaMemcpyAsync(ctx->d_result, ctx->h_result, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CTS_NCCL(ctx, api, api->AllReduce(ctx->d_result, ctx->d_result, 1, ncclDouble, ncclMax, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
  CTS_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *host_val = ctx->h_result[0];
  return CTS_OK;
}

}  // extern "C"

int cts_halo_exchange_on(cts_ctx* ctx, cudaStream_t stream, cts_dtype ft, size_t count, const void* send_lo,
                         const void* send_hi, void* recv_lo, void* recv_hi) {
  if (!ctx->nccl_comm) return cts_fail(ctx, CTS_EINVAL, "cts_halo_exchange: no communicator attached");
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  const ncclDataType_t dt = ft == CTS_F64 ? ncclDouble : ncclFloat;
  const int lo = (ctx->rank - 1 + ctx->nranks) % ctx->nranks, hi = (ctx->rank + 1) % ctx->nranks;
  // With 2 ranks lo == hi and NCCL matches the two messages of a pair in issue order: the peer's
  // first send (its first row) must land in my recv_hi, its second (its last row) in my recv_lo.
  CTS_NCCL(ctx, api, api->GroupStart());
  CTS_NCCL(ctx, api, api->Send(send_lo, count, dt, lo, comm, stream));
  CTS_NCCL(ctx, api, api->Send(send_hi, count, dt, hi, comm, stream));
  CTS_NCCL(ctx, api, api->Recv(recv_hi, count, dt, hi, comm, stream));
  CTS_NCCL(ctx, api, api->Recv(recv_lo, count, dt, lo, comm, stream));
  CTS_NCCL(ctx, api, api->GroupEnd());
  ctx->launches++;
  return CTS_OK;
}
