// Reductions R1-R4: dot, nrm2, sum(1+|x|), and a K-way multi-dot for Gram-Schmidt sweeps.
//
// One launch per reduction: a fixed grid (kReduceBlocks x kReduceThreads, so the summation order -- and
// therefore the result -- is deterministic) streams the operands with 128-bit loads, accumulates in
// Float64, tree-reduces with warp shuffles, and the last block to finish (threadfence + counter)
// folds the per-block partials in a fixed order.  With a communicator attached the K results are
// all-reduced over NCCL in-stream before the single D2H copy.
#include "cts_internal.h"

namespace {

enum RedOp { RED_DOT = 0, RED_SUMSQ = 1, RED_SUM1PABS = 2 };

template <typename T>
struct Vec;
template <>
struct Vec<double> {
  using type = double2;
  static constexpr int N = 2;
};
template <>
struct Vec<float> {
  using type = float4;
  static constexpr int N = 4;
};
__device__ __forceinline__ void unpack(const double2& v, double (&e)[2]) {
  e[0] = v.x;
  e[1] = v.y;
}
__device__ __forceinline__ void unpack(const float4& v, float (&e)[4]) {
  e[0] = v.x;
  e[1] = v.y;
  e[2] = v.z;
  e[3] = v.w;
}

__device__ __forceinline__ void pack(const double (&e)[2], double2& v) { v = make_double2(e[0], e[1]); }
__device__ __forceinline__ void pack(const float (&e)[4], float4& v) { v = make_float4(e[0], e[1], e[2], e[3]); }

template <int K>
struct RedArgs {
  const void* x[K];  // K left operands (multi-dot) -- for K == 1 the single operand
  const void* y;     // shared right operand (dot) or null
  size_t n;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T, int K, int OP>
__device__ __forceinline__ void accumulate(double (&acc)[K], const T (&xv)[K], T yv) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = (double)xv[k];
    if (OP == RED_DOT)
      acc[k] = __dadd_rn(acc[k], __dmul_rn(x, (double)yv));
    else if (OP == RED_SUMSQ)
      acc[k] = __dadd_rn(acc[k], __dmul_rn(x, x));
    else
      acc[k] = __dadd_rn(acc[k], __dadd_rn(1.0, fabs(x)));
  }
}

// Block tree + deterministic last-block fold shared by the plain and the fused (element-wise + reduction) kernels.
template <int K>
__device__ __forceinline__ void block_fold(double (&acc)[K], double* partials, double* result, unsigned int* counter) {
  // block tree
  __shared__ double smem[K][kReduceThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) smem[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double v = lane < kReduceThreads / 32 ? smem[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) partials[(size_t)k * gridDim.x + blockIdx.x] = v;
    }
  }
  // last block folds the partials in a fixed order
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      v = __dadd_rn(v, __ldcg(&partials[(size_t)k * gridDim.x + b]));
    v = warp_sum(v);
    if (lane == 0) smem[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double v = lane < kReduceThreads / 32 ? smem[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) result[k] = v;
    }
    if (lane == 0) *counter = 0;  // ready for the next reduction on this stream
  }
}

template <typename T, int K, int OP, bool VEC>
__global__ void __launch_bounds__(kReduceThreads) reduce_kernel(const RedArgs<K> a, double* partials, double* result,
                                                                unsigned int* counter) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (VEC) {
    const size_t nvec = a.n / VN;
    for (size_t i = tid; i < nvec; i += nthreads) {
      V xv[K];
      V yv;
#pragma unroll
      for (int k = 0; k < K; ++k) xv[k] = reinterpret_cast<const V*>(a.x[k])[i];
      if (OP == RED_DOT) yv = reinterpret_cast<const V*>(a.y)[i];
      T ye[VN];
      if (OP == RED_DOT) unpack(yv, ye);
      T xe[K][VN];
#pragma unroll
      for (int k = 0; k < K; ++k) unpack(xv[k], xe[k]);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        T xs[K];
#pragma unroll
        for (int k = 0; k < K; ++k) xs[k] = xe[k][e];
        accumulate<T, K, OP>(acc, xs, OP == RED_DOT ? ye[e] : T(0));
      }
    }
    done = nvec * VN;
  }
  for (size_t i = done + tid; i < a.n; i += nthreads) {
    T xs[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xs[k] = reinterpret_cast<const T*>(a.x[k])[i];
    accumulate<T, K, OP>(acc, xs, OP == RED_DOT ? reinterpret_cast<const T*>(a.y)[i] : T(0));
  }
  block_fold<K>(acc, partials, result, counter);
}

// Fused producer + reduction: y = a*x + b*y (EW_AXPBY, the MGS update of GMRES) or y = x/d (EW_DIV, the basis-vector
// scaling), and in the same pass dot(y, z) or sum(y^2) of the freshly stored values.  The reduction covers the context's
// window [win_off, win_off + win_n) with exactly the element-to-thread mapping of reduce_kernel, so the result is
// bit-identical to running the element-wise kernel and then cts_dot / cts_nrm2; elements outside the window are only
// updated.
enum EwOp { EW_AXPBY = 0, EW_DIV = 1 };

template <typename T, int EW>
__device__ __forceinline__ T ew_apply(T x, T y, double a, double b) {
  if (EW == EW_AXPBY) return (T)__dadd_rn(__dmul_rn(a, (double)x), __dmul_rn(b, (double)y));
  return (T)__ddiv_rn((double)x, a);
}

template <typename T, int EW, int OP, bool VEC>
__global__ void __launch_bounds__(kReduceThreads) reduce_ew_kernel(T* y, const T* x, const T* z, double a, double b,
                                                                   size_t n_total, size_t win_off, size_t win_n,
                                                                   double* partials, double* result, unsigned int* counter) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  double acc[1] = {0.0};
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  // outside the window: update only
  for (size_t i = tid; i < win_off; i += nthreads) y[i] = ew_apply<T, EW>(x[i], EW == EW_AXPBY ? y[i] : T(0), a, b);
  for (size_t i = win_off + win_n + tid; i < n_total; i += nthreads)
    y[i] = ew_apply<T, EW>(x[i], EW == EW_AXPBY ? y[i] : T(0), a, b);
  T* yw = y + win_off;
  const T* xw = x + win_off;
  const T* zw = OP == RED_DOT ? z + win_off : nullptr;
  size_t done = 0;
  if (VEC) {
    const size_t nvec = win_n / VN;
    for (size_t i = tid; i < nvec; i += nthreads) {
      V xv = reinterpret_cast<const V*>(xw)[i], yv, zv;
      if (EW == EW_AXPBY) yv = reinterpret_cast<const V*>(yw)[i];
      if (OP == RED_DOT) zv = reinterpret_cast<const V*>(zw)[i];
      T xe[VN], ye[VN], ze[VN], oe[VN];
      unpack(xv, xe);
      if (EW == EW_AXPBY) unpack(yv, ye);
      if (OP == RED_DOT) unpack(zv, ze);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        oe[e] = ew_apply<T, EW>(xe[e], EW == EW_AXPBY ? ye[e] : T(0), a, b);
        T xs[1] = {OP == RED_DOT ? ze[e] : oe[e]};
        accumulate<T, 1, OP>(acc, xs, oe[e]);
      }
      V ov;
      pack(oe, ov);
      reinterpret_cast<V*>(yw)[i] = ov;
    }
    done = nvec * VN;
  }
  for (size_t i = done + tid; i < win_n; i += nthreads) {
    const T o = ew_apply<T, EW>(xw[i], EW == EW_AXPBY ? yw[i] : T(0), a, b);
    yw[i] = o;
    T xs[1] = {OP == RED_DOT ? zw[i] : o};
    accumulate<T, 1, OP>(acc, xs, o);
  }
  block_fold<1>(acc, partials, result, counter);
}

template <typename T, int K, int OP>
int launch_reduce(cts_ctx* ctx, size_t n, const void* const* xs, const void* y, double* d_out) {
  RedArgs<K> a;
  size_t off = 0;
  if (ctx->red_window) {
    if (ctx->red_offset + ctx->red_count > n) return cts_fail(ctx, CTS_EINVAL, "reduction window exceeds vector");
    off = ctx->red_offset * sizeof(T);
    n = ctx->red_count;
  }
  bool aligned = true;
  for (int k = 0; k < K; ++k) {
    a.x[k] = static_cast<const char*>(xs[k]) + off;
    aligned = aligned && cts_aligned16(a.x[k]);
  }
  a.y = y ? static_cast<const char*>(y) + off : nullptr;
  if (y) aligned = aligned && cts_aligned16(a.y);
  a.n = n;
  if (aligned)
    reduce_kernel<T, K, OP, true><<<kReduceBlocks, kReduceThreads, 0, ctx->stream>>>(a, ctx->d_partials, d_out, ctx->d_counter);
  else
    reduce_kernel<T, K, OP, false><<<kReduceBlocks, kReduceThreads, 0, ctx->stream>>>(a, ctx->d_partials, d_out, ctx->d_counter);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

int finish(cts_ctx* ctx, int k, double* host_out) {
  if (ctx->nranks > 1) CTS_TRY(cts_allreduce_sum_f64(ctx, ctx->d_result, (size_t)k));
  CTS_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < k; ++j) host_out[j] = ctx->h_result[j];
  return CTS_OK;
}

template <int OP>
int reduce1(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, const void* y, double* host_out) {
  const void* xs[1] = {x};
  if (ft == CTS_F64)
    CTS_TRY((launch_reduce<double, 1, OP>(ctx, n, xs, y, ctx->d_result)));
  else
    CTS_TRY((launch_reduce<float, 1, OP>(ctx, n, xs, y, ctx->d_result)));
  return finish(ctx, 1, host_out);
}

template <typename T>
int multi_dot_t(cts_ctx* ctx, size_t n, int k, const void* const* V, const void* q, double* d_out) {
  switch (k) {
    case 1: return launch_reduce<T, 1, RED_DOT>(ctx, n, V, q, d_out);
    case 2: return launch_reduce<T, 2, RED_DOT>(ctx, n, V, q, d_out);
    case 3: return launch_reduce<T, 3, RED_DOT>(ctx, n, V, q, d_out);
    case 4: return launch_reduce<T, 4, RED_DOT>(ctx, n, V, q, d_out);
    case 5: return launch_reduce<T, 5, RED_DOT>(ctx, n, V, q, d_out);
    case 6: return launch_reduce<T, 6, RED_DOT>(ctx, n, V, q, d_out);
    case 7: return launch_reduce<T, 7, RED_DOT>(ctx, n, V, q, d_out);
    case 8: return launch_reduce<T, 8, RED_DOT>(ctx, n, V, q, d_out);
  }
  return CTS_EINVAL;
}


template <typename T, int EW, int OP>
int launch_reduce_ew(cts_ctx* ctx, size_t n, void* y, const void* x, const void* z, double a, double b) {
  size_t off = 0, cnt = n;
  if (ctx->red_window) {
    if (ctx->red_offset + ctx->red_count > n) return cts_fail(ctx, CTS_EINVAL, "reduction window exceeds vector");
    off = ctx->red_offset;
    cnt = ctx->red_count;
  }
  const bool aligned = cts_aligned16((const T*)y + off) && cts_aligned16((const T*)x + off) && (!z || cts_aligned16((const T*)z + off));
  if (aligned)
    reduce_ew_kernel<T, EW, OP, true><<<kReduceBlocks, kReduceThreads, 0, ctx->stream>>>((T*)y, (const T*)x, (const T*)z, a, b, n, off, cnt, ctx->d_partials, ctx->d_result, ctx->d_counter);
  else
    reduce_ew_kernel<T, EW, OP, false><<<kReduceBlocks, kReduceThreads, 0, ctx->stream>>>((T*)y, (const T*)x, (const T*)z, a, b, n, off, cnt, ctx->d_partials, ctx->d_result, ctx->d_counter);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

}  // namespace

extern "C" {

int cts_axpby_dot(cts_ctx* ctx, cts_dtype ft, size_t n, double a, const void* x, double b, void* y, const void* z,
                  double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (!x || !y || !host_out || x == y) return cts_fail(ctx, CTS_EINVAL, "cts_axpby_dot: bad argument");
  {  // mixed alignment would change the vector/scalar split and with it the summation order: use the separate passes
    const size_t off = ctx->red_window ? ctx->red_offset * cts_sizeof(ft) : 0;
    const bool all16 = cts_aligned16((const char*)x + off) && cts_aligned16((const char*)y + off) && (!z || cts_aligned16((const char*)z + off));
    if (!all16) {
      CTS_TRY(cts_axpby(ctx, ft, n, a, x, b, y));
      if (z) return cts_dot(ctx, ft, n, z, y, host_out);
      double nr = 0;
      CTS_TRY(cts_nrm2(ctx, ft, n, y, &nr));
      *host_out = nr * nr;  // (only the norm is used by callers; exact sum of squares is not recoverable here)
      return CTS_OK;
    }
  }
  if (ft == CTS_F64) {
    if (z) CTS_TRY((launch_reduce_ew<double, EW_AXPBY, RED_DOT>(ctx, n, y, x, z, a, b)));
    else CTS_TRY((launch_reduce_ew<double, EW_AXPBY, RED_SUMSQ>(ctx, n, y, x, nullptr, a, b)));
  } else {
    if (z) CTS_TRY((launch_reduce_ew<float, EW_AXPBY, RED_DOT>(ctx, n, y, x, z, a, b)));
    else CTS_TRY((launch_reduce_ew<float, EW_AXPBY, RED_SUMSQ>(ctx, n, y, x, nullptr, a, b)));
  }
  return finish(ctx, 1, host_out);
}

int cts_div_scalar_nrm2(cts_ctx* ctx, cts_dtype ft, size_t n, void* y, const void* x, double denom, double* host_norm) {
  CTS_CHECK_CTX(ctx);
  if (!x || !y || !host_norm || x == y) return cts_fail(ctx, CTS_EINVAL, "cts_div_scalar_nrm2: bad argument");
  {
    const size_t off = ctx->red_window ? ctx->red_offset * cts_sizeof(ft) : 0;
    if (!(cts_aligned16((const char*)x + off) && cts_aligned16((const char*)y + off))) {
      CTS_TRY(cts_div_scalar(ctx, ft, n, y, x, denom));
      return cts_nrm2(ctx, ft, n, y, host_norm);
    }
  }
  if (ft == CTS_F64) CTS_TRY((launch_reduce_ew<double, EW_DIV, RED_SUMSQ>(ctx, n, y, x, nullptr, denom, 0.0)));
  else CTS_TRY((launch_reduce_ew<float, EW_DIV, RED_SUMSQ>(ctx, n, y, x, nullptr, denom, 0.0)));
  double ss = 0.0;
  CTS_TRY(finish(ctx, 1, &ss));
  *host_norm = sqrt(ss);
  return CTS_OK;
}

int cts_dot(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, const void* y, double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (!x || !y || !host_out) return cts_fail(ctx, CTS_EINVAL, "cts_dot: null argument");
  return reduce1<RED_DOT>(ctx, ft, n, x, y, host_out);
}

int cts_nrm2(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (!x || !host_out) return cts_fail(ctx, CTS_EINVAL, "cts_nrm2: null argument");
  double ss = 0.0;
  CTS_TRY(reduce1<RED_SUMSQ>(ctx, ft, n, x, nullptr, &ss));
  *host_out = sqrt(ss);
  return CTS_OK;
}

int cts_sum1pabs(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (!x || !host_out) return cts_fail(ctx, CTS_EINVAL, "cts_sum1pabs: null argument");
  return reduce1<RED_SUM1PABS>(ctx, ft, n, x, nullptr, host_out);
}

int cts_multi_dot(cts_ctx* ctx, cts_dtype ft, size_t n, int k, const void* const* V, const void* q, double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (k <= 0 || !V || !q || !host_out) return cts_fail(ctx, CTS_EINVAL, "cts_multi_dot: bad argument");
  for (int j0 = 0; j0 < k; j0 += kMaxMulti) {
    int kk = (k - j0 < kMaxMulti) ? k - j0 : kMaxMulti;
    if (ft == CTS_F64)
      CTS_TRY(multi_dot_t<double>(ctx, n, kk, V + j0, q, ctx->d_result));
    else
      CTS_TRY(multi_dot_t<float>(ctx, n, kk, V + j0, q, ctx->d_result));
    CTS_TRY(finish(ctx, kk, host_out + j0));
  }
  return CTS_OK;
}

int cts_set_reduction_window(cts_ctx* ctx, size_t offset, size_t count) {
  CTS_CHECK_CTX(ctx);
  ctx->red_window = count != 0;
  ctx->red_offset = offset;
  ctx->red_count = count;
  return CTS_OK;
}

}  // extern "C"
