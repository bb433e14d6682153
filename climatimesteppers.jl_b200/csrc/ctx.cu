// Context, memory and synthetic-input helpers of libcts_sm100a.so.
#include <mutex>
#include <unordered_set>

#include "cts_internal.h"

// Live-context registry.  Caches, vectors and operators keep a plain pointer to their context, and garbage-collected hosts
// (Julia finalizers, Python __del__) release objects in no particular order: releasing something whose context is already
// gone must not touch the dead context (cts_free then degrades to a plain cudaFree; *_destroy skip the stream work).
namespace {
std::mutex g_ctx_mu;
std::unordered_set<const cts_ctx*>& live_ctxs() {
  static std::unordered_set<const cts_ctx*> s;
  return s;
}
}  // namespace
bool cts_ctx_alive(const cts_ctx* ctx) {
  if (!ctx) return false;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  return live_ctxs().count(ctx) != 0;
}

extern "C" {

const char* cts_version(void) { return "cts-sm100a 0.1.0 (reference: ClimaTimeSteppers.jl v0.10.0)"; }
int cts_abi_version(void) { return CTS_ABI_VERSION; }

int cts_ctx_create(cts_ctx** out, int device, void* cuda_stream) {
  if (!out) return CTS_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return CTS_ENODEVICE;  // no CPU path by design
  }
  if (device < 0 || device >= ndev) return CTS_EINVAL;
  cts_ctx* ctx = new cts_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete ctx;
    return CTS_ECUDA;
  }
  if (cuda_stream) {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return CTS_ECUDA;
    }
    ctx->own_stream = true;
  }
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    live_ctxs().insert(ctx);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  bool ok = cudaMalloc(&ctx->d_result, sizeof(double) * kMaxMulti) == cudaSuccess &&
            cudaMalloc(&ctx->d_counter, sizeof(unsigned int)) == cudaSuccess &&
            cudaMallocHost(&ctx->h_result, sizeof(double) * kMaxMulti) == cudaSuccess &&
            cudaMemset(ctx->d_counter, 0, sizeof(unsigned int)) == cudaSuccess;
  if (!ok) {
    cts_ctx_destroy(ctx);
    return CTS_ECUDA;
  }
  *out = ctx;
  return CTS_OK;
}

int cts_comm_destroy(cts_ctx* ctx);

int cts_ctx_destroy(cts_ctx* ctx) {
  if (!ctx) return CTS_EINVAL;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!live_ctxs().erase(ctx)) return CTS_EINVAL;  // already destroyed (or never created here)
  }
  cudaSetDevice(ctx->device);
  if (ctx->nccl_comm) cts_comm_destroy(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_partials_g) cudaFree(ctx->d_partials_g);
  if (ctx->d_result) cudaFree(ctx->d_result);
  if (ctx->d_counter) cudaFree(ctx->d_counter);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->comm_stream) {
    cudaStreamSynchronize(ctx->comm_stream);
    cudaStreamDestroy(ctx->comm_stream);
  }
  if (ctx->ev_compute) cudaEventDestroy(ctx->ev_compute);
  if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CTS_OK;
}

int cts_ctx_sync(cts_ctx* ctx) {
  CTS_CHECK_CTX(ctx);
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CTS_OK;
}

const char* cts_last_error(cts_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* cts_ctx_stream(cts_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int cts_launch_count(cts_ctx* ctx, uint64_t* out) {
  CTS_CHECK_CTX(ctx);
  if (!out) return CTS_EINVAL;
  *out = ctx->launches;
  return CTS_OK;
}

int cts_nvtx_range_push(const char* name) {
  if (!name) return CTS_EINVAL;
  nvtxRangePushA(name);
  return CTS_OK;
}

int cts_nvtx_range_pop(void) {
  nvtxRangePop();
  return CTS_OK;
}

int cts_bytes_count(cts_ctx* ctx, uint64_t* out) {
  CTS_CHECK_CTX(ctx);
  if (!out) return CTS_EINVAL;
  *out = ctx->bytes;
  return CTS_OK;
}

int cts_prof_enable(cts_ctx* ctx, int enabled) {
  CTS_CHECK_CTX(ctx);
  ctx->prof_on = enabled != 0;
  return CTS_OK;
}

int cts_prof_read(cts_ctx* ctx, double* ms, uint64_t* launches) {
  CTS_CHECK_CTX(ctx);
  if (!ms || !launches) return CTS_EINVAL;
  for (int k = 0; k < CTS_PROF_NCAT; ++k) {
    ms[k] = 0.0;
    launches[k] = 0;
  }
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cat >= 0 && r.cat < CTS_PROF_NCAT) {
      ms[r.cat] += t;
      launches[r.cat] += 1;
    }
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_recs.clear();
  return CTS_OK;
}

int cts_alloc(cts_ctx* ctx, size_t bytes, void** dptr) {
  CTS_CHECK_CTX(ctx);
  if (!dptr) return CTS_EINVAL;
  *dptr = nullptr;
  if (bytes == 0) return CTS_OK;
  CTS_CUDA(ctx, cudaSetDevice(ctx->device));
  CTS_CUDA(ctx, cudaMalloc(dptr, bytes));
  CTS_CUDA(ctx, cudaMemsetAsync(*dptr, 0, bytes, ctx->stream));
  return CTS_OK;
}

int cts_free(cts_ctx* ctx, void* dptr) {
  CTS_CHECK_CTX(ctx);
  if (!dptr) return CTS_OK;
  if (!cts_ctx_alive(ctx)) {  // released after its context (finalizer order): cudaFree synchronises the device itself
    cudaFree(dptr);
    cudaGetLastError();
    return CTS_OK;
  }
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CTS_CUDA(ctx, cudaFree(dptr));
  return CTS_OK;
}

int cts_memset_zero(cts_ctx* ctx, void* dptr, size_t bytes) {
  CTS_CHECK_CTX(ctx);
  if (bytes) CTS_CUDA(ctx, cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
  return CTS_OK;
}

int cts_memcpy_h2d(cts_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  CTS_CHECK_CTX(ctx);
  if (bytes) CTS_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return CTS_OK;
}

int cts_memcpy_d2h_sync(cts_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  CTS_CHECK_CTX(ctx);
  if (bytes) CTS_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CTS_OK;
}

int cts_memcpy_d2h(cts_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  CTS_CHECK_CTX(ctx);
  if (bytes) CTS_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return CTS_OK;
}

int cts_memcpy_d2d(cts_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
  CTS_CHECK_CTX(ctx);
  if (bytes) CTS_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return CTS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// counter-based synthetic inputs: splitmix64(seed ^ idx) -> uniform [0,1)   
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t cts_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <typename T>
__global__ void fill_hash_kernel(T* out, size_t n, uint64_t seed, uint64_t offset, double scale, double shift) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint64_t h = cts_splitmix64(seed ^ (offset + i));
    double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);  // 53-bit mantissa -> [0,1)
    out[i] = (T)__dmul_rn(scale, __dadd_rn(u, shift));
  }
}

extern "C" int cts_fill_hash(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, uint64_t seed, uint64_t global_offset,
                             double scale, double shift) {
  CTS_CHECK_CTX(ctx);
  if (n == 0) return CTS_OK;
  int threads = 256;
  size_t want = (n + threads - 1) / threads;
  int blocks = (int)(want < (size_t)ctx->num_sms * 16 ? want : (size_t)ctx->num_sms * 16);
  if (ft == CTS_F64)
    fill_hash_kernel<double><<<blocks, threads, 0, ctx->stream>>>((double*)out, n, seed, global_offset, scale, shift);
  else
    fill_hash_kernel<float><<<blocks, threads, 0, ctx->stream>>>((float*)out, n, seed, global_offset, scale, shift);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}
