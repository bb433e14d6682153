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

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "cts_internal.h"

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
    LOAD(AllGather, "ncclAllGather")
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

// =============================================================================================
// Peer-memory halo exchange over NVLink / NVSwitch (one kernel per dss!, no NCCL on the data path)
//
// Every rank owns a "mailbox" (cudaMalloc + CUDA IPC, mapped by its two ring neighbours): two arrival counters and
// 2 directions x 2 slots x row bytes.  One exchange = one launch of halo_peer_kernel on every rank:
//   push  : the CTAs store this rank's first owned row into the lower neighbour's mailbox[from_hi][slot] and its last
//           owned row into the upper neighbour's mailbox[from_lo][slot] with 16-byte stores over NVLink, then
//           __threadfence_system(); the last CTA of a direction to get there adds 1 (red.release.sys) to the neighbour's
//           arrival counter, which therefore counts exchanges;
//   pull  : the same CTAs spin (ld.acquire.sys) until their own counter has reached exchange #seq and copy
//           mailbox -> ghost row.
// seq lives in device memory (so a captured CUDA graph can be replayed), slot = seq & 1.  Two slots suffice: a rank can
// only start exchange s+1 after it has received its neighbours' exchange-s rows, which they push at the START of their
// exchange s, i.e. after they have finished unpacking exchange s-1 (the previous user of slot (s+1)&1).  Exchanges of one
// context must be ordered with respect to each other (they are: same stream or event chains).  A spin that sees no
// arrival for a minute gives up and raises the error flag (a dead peer must not hang the GPU).
// =============================================================================================
namespace {

constexpr int kHaloThreads = 512;
constexpr unsigned long long kHaloTimeoutNs = 60ull * 1000000000ull;
constexpr size_t kPeerHeaderBytes = 256;  // [0]: arrivals from the lower neighbour, [1]: from the upper neighbour

struct HaloPeerArgs {
  const uint4* send_lo;   // my first owned row  -> lower neighbour's from_hi area
  const uint4* send_hi;   // my last owned row   -> upper neighbour's from_lo area
  uint4* recv_lo;         // my lower ghost row  <- my from_lo area
  uint4* recv_hi;         // my upper ghost row  <- my from_hi area
  char* mine;             // my mailbox (header + data)
  char* peer_lo;          // lower neighbour's mailbox (peer mapping)
  char* peer_hi;          // upper neighbour's mailbox
  size_t nvec;            // 16-byte vectors per row
  size_t slot_bytes;      // capacity of one slot
  int ctas_per_dir;
  unsigned long long* seq;   // local: completed exchanges
  unsigned int* done;        // local: CTA ticket
  unsigned int* pushed;      // local: [2] CTAs of a direction that have stored their chunk
  unsigned int* error;       // local: set when a wait timed out
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// data area of a mailbox: [dir][slot][slot_bytes]; dir 0 = rows that came from the lower neighbour
__device__ __forceinline__ uint4* mailbox_slot(char* base, int dir, int slot, size_t slot_bytes) {
  return reinterpret_cast<uint4*>(base + kPeerHeaderBytes + ((size_t)dir * 2 + slot) * slot_bytes);
}

__global__ void __launch_bounds__(kHaloThreads) halo_peer_kernel(const HaloPeerArgs a) {
  const unsigned long long s = *a.seq + 1;  // this exchange (seq is only advanced by the last CTA of a launch)
  const int slot = (int)(s & 1);
  const int dir = blockIdx.x / a.ctas_per_dir;        // 0: talk to the lower neighbour, 1: to the upper one
  const int part = blockIdx.x % a.ctas_per_dir;
  const size_t per = (a.nvec + a.ctas_per_dir - 1) / a.ctas_per_dir;
  const size_t v0 = (size_t)part * per, v1 = v0 + per < a.nvec ? v0 + per : a.nvec;
  // ---- push: my boundary row -> the neighbour's mailbox (what I send down arrives there "from the upper neighbour")
  {
    const uint4* src = dir == 0 ? a.send_lo : a.send_hi;
    char* peer = dir == 0 ? a.peer_lo : a.peer_hi;
    uint4* dst = mailbox_slot(peer, dir == 0 ? 1 : 0, slot, a.slot_bytes);
    for (size_t v = v0 + threadIdx.x; v < v1; v += kHaloThreads) dst[v] = src[v];
    __threadfence_system();
    __syncthreads();
    // the last CTA of this direction to finish its chunk signals the neighbour: ONE arrival per exchange and direction,
    // so the counter is the exchange number whatever the row size (and with it the CTA count) of each exchange is
    if (threadIdx.x == 0) {
      const unsigned int t = atomicAdd(a.pushed + dir, 1u);
      if (t == (unsigned)a.ctas_per_dir - 1) {
        a.pushed[dir] = 0;
        __threadfence_system();
        red_release_sys_add(reinterpret_cast<unsigned long long*>(peer) + (dir == 0 ? 1 : 0), 1ull);
      }
    }
  }
  // ---- pull: wait for all chunks of exchange s from the neighbour on this side, then unpack my mailbox
  {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(a.mine) + dir;
    const unsigned long long want = s;
    __shared__ int ok;
    if (threadIdx.x == 0) {
      ok = 1;
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(flag) < want) {
        __nanosleep(64);
        // a neighbour may legitimately be seconds behind (host-side work between steps); a minute means it is gone.
        // Once the flag is up every later wait gives up at once, so a dead peer costs one timeout, not one per exchange.
        if (*reinterpret_cast<volatile unsigned int*>(a.error) != 0u || global_timer_ns() - t0 > kHaloTimeoutNs) {
          ok = 0;
          atomicExch(a.error, 1u);
          break;
        }
      }
    }
    __syncthreads();
    if (ok) {
      const uint4* src = mailbox_slot(a.mine, dir, slot, a.slot_bytes);
      uint4* dst = dir == 0 ? a.recv_lo : a.recv_hi;
      for (size_t v = v0 + threadIdx.x; v < v1; v += kHaloThreads) dst[v] = __ldcv(&src[v]);
    }
  }
  // ---- the last CTA advances the exchange counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(a.done, 1u);
    if (t == gridDim.x - 1) {
      *a.done = 0;
      *a.seq = s;
      __threadfence();
    }
  }
}

}  // namespace

struct cts_peer_halo {
  char* mine = nullptr;           // cudaMalloc'ed mailbox
  char* peer_lo = nullptr;        // IPC mappings (peer_hi may equal peer_lo with 2 ranks)
  char* peer_hi = nullptr;
  bool same_peer = false;
  size_t slot_bytes = 0;
  unsigned long long* seq = nullptr;  // [seq][done|error] local control words
  unsigned int* done = nullptr;
  unsigned int* pushed = nullptr;
  unsigned int* error = nullptr;
  bool ok = false;
};

static void peer_halo_free(cts_ctx* ctx) {
  cts_peer_halo* p = ctx->peer;
  if (!p) return;
  if (p->peer_lo) cudaIpcCloseMemHandle(p->peer_lo);
  if (p->peer_hi && !p->same_peer) cudaIpcCloseMemHandle(p->peer_hi);
  if (p->mine) cudaFree(p->mine);
  if (p->seq) cudaFree(p->seq);
  delete p;
  ctx->peer = nullptr;
}

// Collective over the communicator: (re)creates the mailboxes for rows of up to `row_bytes` and maps the neighbours'.
// Every rank ends with the same answer (ok or not): a rank that cannot map its neighbours makes all ranks use NCCL.
static int peer_halo_setup(cts_ctx* ctx, size_t row_bytes) {
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CTS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
  peer_halo_free(ctx);
  auto* p = new cts_peer_halo();
  ctx->peer = p;
  p->slot_bytes = (row_bytes + 255) / 256 * 256;
  const size_t bytes = kPeerHeaderBytes + 4 * p->slot_bytes;
  int good = 1;
  if (cudaMalloc(&p->mine, bytes) != cudaSuccess || cudaMemset(p->mine, 0, kPeerHeaderBytes) != cudaSuccess) good = 0;
  if (good && (cudaMalloc(&p->seq, 64) != cudaSuccess || cudaMemset(p->seq, 0, 64) != cudaSuccess)) good = 0;
  if (good) {
    p->done = reinterpret_cast<unsigned int*>(p->seq + 1);
    p->error = p->done + 1;
    p->pushed = p->done + 2;
  }
  // all-gather the IPC handles (64 bytes each, padded to 128) through the communicator the context already has
  constexpr size_t kSlot = 128;
  static_assert(sizeof(cudaIpcMemHandle_t) <= kSlot - 8, "IPC handle size");
  std::vector<unsigned char> h_all(kSlot * ctx->nranks, 0), h_mine(kSlot, 0);
  if (good) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p->mine) == cudaSuccess) {
      memcpy(h_mine.data(), &h, sizeof(h));
      h_mine[kSlot - 1] = 1;  // "valid"
    } else {
      good = 0;
    }
  }
  cudaGetLastError();
  unsigned char* d_buf = nullptr;
  CTS_CUDA(ctx, cudaMalloc(&d_buf, kSlot * (ctx->nranks + 1)));
  CTS_CUDA(ctx, cudaMemcpyAsync(d_buf, h_mine.data(), kSlot, cudaMemcpyHostToDevice, ctx->stream));
  CTS_NCCL(ctx, api, api->AllGather(d_buf, d_buf + kSlot, kSlot, ncclUint8, comm, ctx->stream));
  CTS_CUDA(ctx, cudaMemcpyAsync(h_all.data(), d_buf + kSlot, kSlot * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int lo = (ctx->rank - 1 + ctx->nranks) % ctx->nranks, hi = (ctx->rank + 1) % ctx->nranks;
  for (int r = 0; r < ctx->nranks; ++r) good = good && h_all[kSlot * r + kSlot - 1] == 1;
  if (good) {
    cudaIpcMemHandle_t h;
    memcpy(&h, &h_all[kSlot * lo], sizeof(h));
    if (cudaIpcOpenMemHandle(reinterpret_cast<void**>(&p->peer_lo), h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) good = 0;
    if (good && hi == lo) {
      p->peer_hi = p->peer_lo;
      p->same_peer = true;
    } else if (good) {
      memcpy(&h, &h_all[kSlot * hi], sizeof(h));
      if (cudaIpcOpenMemHandle(reinterpret_cast<void**>(&p->peer_hi), h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) good = 0;
    }
    if (!good) ctx->err = std::string("peer halo: cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
  }
  cudaGetLastError();
  // agree on the outcome (min over ranks)
  double flag = good ? 1.0 : 0.0;
  CTS_CUDA(ctx, cudaMemcpyAsync(d_buf, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CTS_NCCL(ctx, api, api->AllReduce(d_buf, d_buf, 1, ncclDouble, ncclMin, comm, ctx->stream));
  CTS_CUDA(ctx, cudaMemcpyAsync(&flag, d_buf, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_buf);
  p->ok = flag == 1.0;
  return CTS_OK;
}

static int peer_halo_exchange(cts_ctx* ctx, cudaStream_t stream, size_t bytes, const void* send_lo, const void* send_hi,
                              void* recv_lo, void* recv_hi) {
  cts_peer_halo* p = ctx->peer;
  HaloPeerArgs a;
  a.send_lo = static_cast<const uint4*>(send_lo);
  a.send_hi = static_cast<const uint4*>(send_hi);
  a.recv_lo = static_cast<uint4*>(recv_lo);
  a.recv_hi = static_cast<uint4*>(recv_hi);
  a.mine = p->mine;
  a.peer_lo = p->peer_lo;
  a.peer_hi = p->peer_hi;
  a.nvec = bytes / 16;
  a.slot_bytes = p->slot_bytes;
  int per_dir = (int)((bytes + 32767) / 32768);
  a.ctas_per_dir = per_dir < 1 ? 1 : (per_dir > 32 ? 32 : per_dir);
  a.seq = p->seq;
  a.done = p->done;
  a.pushed = p->pushed;
  a.error = p->error;
  halo_peer_kernel<<<2 * a.ctas_per_dir, kHaloThreads, 0, stream>>>(a);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

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
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  peer_halo_free(ctx);
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

}  // extern "C"

int cts_allreduce_sum_f64_oop(cts_ctx* ctx, const double* send, double* recv, size_t count) {
  if (ctx->nranks <= 1) return cts_fail(ctx, CTS_EINVAL, "cts_allreduce_sum_f64_oop: no communicator");
  NcclApi* api = nccl_api();
  CTS_NCCL(ctx, api, api->AllReduce(send, recv, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
  ctx->launches++;
  return CTS_OK;
}

extern "C" {

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

extern "C" int cts_comm_enable_peer_halo(cts_ctx* ctx, size_t max_row_bytes) {
  CTS_CHECK_CTX(ctx);
  if (!ctx->nccl_comm || ctx->nranks <= 1) return cts_fail(ctx, CTS_EINVAL, "cts_comm_enable_peer_halo: no communicator attached");
  if (getenv("CTS_HALO_NCCL")) {  // developer switch: keep the NCCL send/recv path
    ctx->peer_mode = -1;
    return CTS_OK;
  }
  CTS_TRY(peer_halo_setup(ctx, max_row_bytes));
  ctx->peer_mode = ctx->peer->ok ? 1 : -1;
  return CTS_OK;
}

extern "C" int cts_comm_halo_path(cts_ctx* ctx, int* path) {
  CTS_CHECK_CTX(ctx);
  if (!path) return CTS_EINVAL;
  if (ctx->peer && ctx->peer->ok && ctx->peer->error) {  // surface a timed-out wait
    unsigned int e = 0;
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(&e, ctx->peer->error, sizeof(e), cudaMemcpyDeviceToHost);
    if (e) return cts_fail(ctx, CTS_ECUDA, "peer halo: a wait for the neighbour's rows timed out");
  }
  *path = ctx->nranks <= 1 ? 0 : (ctx->peer_mode == 1 ? 2 : 1);
  return CTS_OK;
}

int cts_halo_exchange_on(cts_ctx* ctx, cudaStream_t stream, cts_dtype ft, size_t count, const void* send_lo,
                         const void* send_hi, void* recv_lo, void* recv_hi) {
  if (!ctx->nccl_comm) return cts_fail(ctx, CTS_EINVAL, "cts_halo_exchange: no communicator attached");
  {
    const size_t bytes = count * cts_sizeof(ft);
    const bool vec_ok = bytes % 16 == 0 && cts_aligned16(send_lo) && cts_aligned16(send_hi) && cts_aligned16(recv_lo) && cts_aligned16(recv_hi);
    if (ctx->peer_mode == 1 && vec_ok && bytes <= ctx->peer->slot_bytes)
      return peer_halo_exchange(ctx, stream, bytes, send_lo, send_hi, recv_lo, recv_hi);
  }
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
