// Reductions R1-R4: dot, nrm2, sum(1+|x|), and a K-way multi-dot for Gram-Schmidt sweeps.
//
// The summation order is part of the result's definition (LinearAlgebra.norm's own order is implementation defined,
// SURVEY A.6), and it is defined on GLOBAL element indices so that it does not depend on how the state is sharded:
//   * the reduction domain [0, G) is cut into fixed chunks of CH elements (default kReduceChunk; a power of two);
//   * one CTA reduces one chunk: thread t accumulates (in Float64, from 0) the 16-byte vectors t, t+256, t+512, ... of the
//     chunk, elements of a vector in order, then a warp xor-butterfly and a fold of the 8 warp sums -> chunk partial;
//   * the chunk partials are folded in global chunk order by one CTA: thread t takes partials t, t+256, ..., same tree.
// On one rank the last CTA to finish does that fold in the same launch.  With a communicator and a layout
// (cts_set_reduction_layout: this rank owns whole chunks [g0/CH, (g0+count)/CH) of the global chunk array) the ranks
// exchange chunk partials (an out-of-place sum all-reduce of an array in which every entry has exactly one non-zero
// contributor, i.e. an exact gather) and every rank runs the identical fold: 1, 2, 4 or 8 ranks give the same bits.
// Without a layout the per-rank results are all-reduced (deterministic, but dependent on the rank count).
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
  const void* x[K];  // K left operands (multi-dot) -- for K == 1 the single operand; already offset to the window start
  const void* y;     // shared right operand (dot) or null
  size_t n;          // window length
};

// where the chunk partials go and who folds them
struct FoldArgs {
  double* partials;   // [K][nc_global]
  size_t nc_global;   // chunks of the global domain
  size_t chunk0;      // global index of this rank's first chunk
  unsigned chunk;     // elements per chunk
  double* result;     // [K]
  unsigned int* counter;
  int fold_here;      // 1: the last CTA folds all partials into result (single rank); 0: a separate launch does
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T, int OP>
__device__ __forceinline__ double term(T x, T y) {
  const double xd = (double)x;
  if (OP == RED_DOT) return __dmul_rn(xd, (double)y);
  if (OP == RED_SUMSQ) return __dmul_rn(xd, xd);
  return __dadd_rn(1.0, fabs(xd));
}

// 256 per-thread values -> one value (valid in warp 0, lane 0)
template <int K>
__device__ __forceinline__ void block_tree(double (&acc)[K], double (&out)[K]) {
  __shared__ double smem[K][kReduceThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // smem may still be read by the previous use
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) smem[k][warp] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = lane < kReduceThreads / 32 ? smem[k][lane] : 0.0;
    out[k] = warp_sum(v);
  }
}

// fold of the global chunk partials by one CTA: thread t takes partials t, t+256, ...
template <int K>
__device__ __forceinline__ void fold_partials(const double* partials, size_t nc, double* result) {
  double acc[K], out[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = 0.0;
    for (size_t b = threadIdx.x; b < nc; b += kReduceThreads) v = __dadd_rn(v, __ldcg(&partials[(size_t)k * nc + b]));
    acc[k] = v;
  }
  block_tree<K>(acc, out);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) result[k] = out[k];
  }
}

// Chunk partial -> partials array; on one rank the last CTA to finish folds them.
template <int K>
__device__ __forceinline__ void chunk_finish(double (&acc)[K], const FoldArgs& f) {
  double out[K];
  block_tree<K>(acc, out);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) f.partials[(size_t)k * f.nc_global + f.chunk0 + blockIdx.x] = out[k];
  }
  if (!f.fold_here) return;
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(f.counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  fold_partials<K>(f.partials, f.nc_global, f.result);
  if (threadIdx.x == 0) *f.counter = 0;  // ready for the next reduction on this stream
}

template <int K>
__global__ void __launch_bounds__(kReduceThreads) reduce_fold_kernel(const double* partials, size_t nc, double* result) {
  fold_partials<K>(partials, nc, result);
}

// One CTA per chunk.  VEC: the window start is 16-byte aligned (the accumulation order is the same either way).
template <typename T, int K, int OP, bool VEC>
__global__ void __launch_bounds__(kReduceThreads) reduce_kernel(const RedArgs<K> a, const FoldArgs f) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int UN = K == 1 ? 8 : (K <= 2 ? 4 : 2);  // vectors in flight per thread and operand
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const size_t e0 = (size_t)blockIdx.x * f.chunk;                       // first element of the chunk inside the window
  const size_t len = e0 >= a.n ? 0 : (a.n - e0 < (size_t)f.chunk ? a.n - e0 : (size_t)f.chunk);
  const unsigned nfull = (unsigned)(len / VN);                           // whole vectors of this chunk
  if (VEC) {
    unsigned j = threadIdx.x;
    for (; j + (UN - 1) * kReduceThreads < nfull; j += UN * kReduceThreads) {
      V xv[K][UN], yv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const size_t vi = e0 / VN + j + u * kReduceThreads;
#pragma unroll
        for (int k = 0; k < K; ++k) xv[k][u] = reinterpret_cast<const V*>(a.x[k])[vi];
        if (OP == RED_DOT) yv[u] = reinterpret_cast<const V*>(a.y)[vi];
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        T ye[VN];
        if (OP == RED_DOT) unpack(yv[u], ye);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          T xe[VN];
          unpack(xv[k][u], xe);
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[k] = __dadd_rn(acc[k], term<T, OP>(xe[e], OP == RED_DOT ? ye[e] : T(0)));
        }
      }
    }
    for (; j < nfull; j += kReduceThreads) {
      const size_t vi = e0 / VN + j;
      T ye[VN];
      if (OP == RED_DOT) unpack(reinterpret_cast<const V*>(a.y)[vi], ye);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        T xe[VN];
        unpack(reinterpret_cast<const V*>(a.x[k])[vi], xe);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[k] = __dadd_rn(acc[k], term<T, OP>(xe[e], OP == RED_DOT ? ye[e] : T(0)));
      }
    }
  } else {
    for (unsigned j = threadIdx.x; j < nfull; j += kReduceThreads) {
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const size_t i = e0 + (size_t)j * VN + e;
        const T yv = OP == RED_DOT ? reinterpret_cast<const T*>(a.y)[i] : T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], term<T, OP>(reinterpret_cast<const T*>(a.x[k])[i], yv));
      }
    }
  }
  // the (at most one) partial vector at the end of a short chunk belongs to thread nfull % 256, after its whole vectors
  if (len % VN != 0 && threadIdx.x == nfull % kReduceThreads) {
    for (size_t i = e0 + (size_t)nfull * VN; i < e0 + len; ++i) {
      const T yv = OP == RED_DOT ? reinterpret_cast<const T*>(a.y)[i] : T(0);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], term<T, OP>(reinterpret_cast<const T*>(a.x[k])[i], yv));
    }
  }
  chunk_finish<K>(acc, f);
}

// Fused producer + reduction: y = a*x + b*y (EW_AXPBY, the MGS update of GMRES) or y = x/d (EW_DIV, the basis-vector
// scaling), and in the same pass dot(y, z) or sum(y^2) of the freshly stored values.  The reduction covers the context's
// window [win_off, win_off + win_n) with exactly the element-to-thread mapping of reduce_kernel, so the result is
// bit-identical to running the element-wise kernel and then cts_dot / cts_nrm2; elements outside the window are only
// updated (by the same CTAs, grid-stride).
enum EwOp { EW_AXPBY = 0, EW_DIV = 1 };

template <typename T, int EW>
__device__ __forceinline__ T ew_apply(T x, T y, double a, double b) {
  if (EW == EW_AXPBY) return (T)__dadd_rn(__dmul_rn(a, (double)x), __dmul_rn(b, (double)y));
  return (T)__ddiv_rn((double)x, a);
}

template <typename T, int EW, int OP, bool VEC>
__global__ void __launch_bounds__(kReduceThreads) reduce_ew_kernel(T* y, const T* x, const T* z, double a, double b,
                                                                   size_t n_total, size_t win_off, size_t win_n,
                                                                   const FoldArgs f) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int UN = 4;
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
  const size_t e0 = (size_t)blockIdx.x * f.chunk;
  const size_t len = e0 >= win_n ? 0 : (win_n - e0 < (size_t)f.chunk ? win_n - e0 : (size_t)f.chunk);
  const unsigned nfull = (unsigned)(len / VN);
  auto one = [&](size_t i) {
    const T o = ew_apply<T, EW>(xw[i], EW == EW_AXPBY ? yw[i] : T(0), a, b);
    yw[i] = o;
    acc[0] = __dadd_rn(acc[0], OP == RED_DOT ? term<T, RED_DOT>(zw[i], o) : term<T, RED_SUMSQ>(o, o));
  };
  if (VEC) {
    auto vec_round = [&](const V& xv, const V& yv, const V& zv, size_t vi) {
      T xe[VN], ye[VN], ze[VN], oe[VN];
      unpack(xv, xe);
      if (EW == EW_AXPBY) unpack(yv, ye);
      if (OP == RED_DOT) unpack(zv, ze);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        oe[e] = ew_apply<T, EW>(xe[e], EW == EW_AXPBY ? ye[e] : T(0), a, b);
        acc[0] = __dadd_rn(acc[0], OP == RED_DOT ? term<T, RED_DOT>(ze[e], oe[e]) : term<T, RED_SUMSQ>(oe[e], oe[e]));
      }
      V ov;
      pack(oe, ov);
      reinterpret_cast<V*>(yw)[vi] = ov;
    };
    unsigned j = threadIdx.x;
    for (; j + (UN - 1) * kReduceThreads < nfull; j += UN * kReduceThreads) {
      V xv[UN], yv[UN], zv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const size_t vi = e0 / VN + j + u * kReduceThreads;
        xv[u] = reinterpret_cast<const V*>(xw)[vi];
        if (EW == EW_AXPBY) yv[u] = reinterpret_cast<const V*>(yw)[vi];
        if (OP == RED_DOT) zv[u] = reinterpret_cast<const V*>(zw)[vi];
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) vec_round(xv[u], yv[u], zv[u], e0 / VN + j + u * kReduceThreads);
    }
    for (; j < nfull; j += kReduceThreads) {
      const size_t vi = e0 / VN + j;
      V xv = reinterpret_cast<const V*>(xw)[vi], yv, zv;
      if (EW == EW_AXPBY) yv = reinterpret_cast<const V*>(yw)[vi];
      if (OP == RED_DOT) zv = reinterpret_cast<const V*>(zw)[vi];
      vec_round(xv, yv, zv, vi);
    }
  } else {
    for (unsigned j = threadIdx.x; j < nfull; j += kReduceThreads)
#pragma unroll
      for (int e = 0; e < VN; ++e) one(e0 + (size_t)j * VN + e);
  }
  if (len % VN != 0 && threadIdx.x == nfull % kReduceThreads)
    for (size_t i = e0 + (size_t)nfull * VN; i < e0 + len; ++i) one(i);
  chunk_finish<1>(acc, f);
}

// ---- host side ---------------------------------------------------------------------------------
struct Plan {
  size_t off = 0, n = 0;  // window inside the local vector (elements)
  FoldArgs f{};
  size_t nchunks = 0;     // chunks owned by this rank (= CTAs)
  bool exchange = false;  // multi-rank layout: chunk partials are exchanged, then folded by a second launch
};

int ensure_partials(cts_ctx* ctx, size_t doubles) {
  if (doubles <= ctx->partials_cap) return CTS_OK;
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_partials_g) cudaFree(ctx->d_partials_g);
  ctx->d_partials = ctx->d_partials_g = nullptr;
  ctx->partials_cap = 0;
  size_t cap = doubles < 4096 ? 4096 : doubles + doubles / 2;
  CTS_CUDA(ctx, cudaMalloc(&ctx->d_partials, cap * sizeof(double)));
  CTS_CUDA(ctx, cudaMalloc(&ctx->d_partials_g, cap * sizeof(double)));
  CTS_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, cap * sizeof(double), ctx->stream));
  ctx->partials_cap = cap;
  return CTS_OK;
}

int make_plan(cts_ctx* ctx, size_t n, int K, Plan* p) {
  p->off = 0;
  p->n = n;
  if (ctx->red_window) {
    if (ctx->red_offset + ctx->red_count > n) return cts_fail(ctx, CTS_EINVAL, "reduction window exceeds vector");
    p->off = ctx->red_offset;
    p->n = ctx->red_count;
  }
  const unsigned chunk = ctx->red_chunk ? ctx->red_chunk : (unsigned)kReduceChunk;
  p->nchunks = (p->n + chunk - 1) / chunk;
  if (p->nchunks == 0) p->nchunks = 1;  // an empty window still runs one CTA (result 0)
  p->f.chunk = chunk;
  p->f.result = ctx->d_result;
  p->f.counter = ctx->d_counter;
  const bool layout = ctx->red_layout && ctx->red_window;
  if (layout) {
    p->f.nc_global = (ctx->red_G + chunk - 1) / chunk;
    p->f.chunk0 = ctx->red_g0 / chunk;
  } else {
    p->f.nc_global = p->nchunks;
    p->f.chunk0 = 0;
  }
  p->exchange = layout && ctx->nranks > 1;
  p->f.fold_here = p->exchange ? 0 : 1;
  if (p->nchunks > 0x7fffffffull) return cts_fail(ctx, CTS_EINVAL, "reduction: oversized window");
  if (p->f.nc_global < p->f.chunk0 + p->nchunks) p->f.nc_global = p->f.chunk0 + p->nchunks;
  // (entries of d_partials outside this rank's chunk range are never written: they stay zero from the allocation)
  CTS_TRY(ensure_partials(ctx, (size_t)K * p->f.nc_global));
  p->f.partials = ctx->d_partials;
  return CTS_OK;
}

template <typename T, int K, int OP>
int launch_reduce(cts_ctx* ctx, size_t n, const void* const* xs, const void* y) {
  Plan p;
  CTS_TRY(make_plan(ctx, n, K, &p));
  RedArgs<K> a;
  bool aligned = true;
  for (int k = 0; k < K; ++k) {
    a.x[k] = static_cast<const char*>(xs[k]) + p.off * sizeof(T);
    aligned = aligned && cts_aligned16(a.x[k]);
  }
  a.y = y ? static_cast<const char*>(y) + p.off * sizeof(T) : nullptr;
  if (y) aligned = aligned && cts_aligned16(a.y);
  a.n = p.n;
  CTS_COUNT_BYTES(ctx, K + (y ? 1 : 0), p.n, sizeof(T));  // R1-R4: dot 2w, nrm2 / sum 1w, k-way multi-dot (k+1)w
  if (aligned)
    reduce_kernel<T, K, OP, true><<<(unsigned)p.nchunks, kReduceThreads, 0, ctx->stream>>>(a, p.f);
  else
    reduce_kernel<T, K, OP, false><<<(unsigned)p.nchunks, kReduceThreads, 0, ctx->stream>>>(a, p.f);
  CTS_LAUNCH_CHECK(ctx);
  if (p.exchange) {
    const size_t cnt = (size_t)K * p.f.nc_global;
    CTS_TRY(cts_allreduce_sum_f64_oop(ctx, ctx->d_partials, ctx->d_partials_g, cnt));
    reduce_fold_kernel<K><<<1, kReduceThreads, 0, ctx->stream>>>(ctx->d_partials_g, p.f.nc_global, ctx->d_result);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

int finish(cts_ctx* ctx, int k, double* host_out) {
  // ranks without a layout combine their (rank-local) results; with a layout the exchange already happened
  if (ctx->nranks > 1 && !(ctx->red_layout && ctx->red_window)) CTS_TRY(cts_allreduce_sum_f64(ctx, ctx->d_result, (size_t)k));
  CTS_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < k; ++j) host_out[j] = ctx->h_result[j];
  return CTS_OK;
}

template <int OP>
int reduce1(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, const void* y, double* host_out) {
  const void* xs[1] = {x};
  if (ft == CTS_F64)
    CTS_TRY((launch_reduce<double, 1, OP>(ctx, n, xs, y)));
  else
    CTS_TRY((launch_reduce<float, 1, OP>(ctx, n, xs, y)));
  return finish(ctx, 1, host_out);
}

template <typename T>
int multi_dot_t(cts_ctx* ctx, size_t n, int k, const void* const* V, const void* q) {
  switch (k) {
    case 1: return launch_reduce<T, 1, RED_DOT>(ctx, n, V, q);
    case 2: return launch_reduce<T, 2, RED_DOT>(ctx, n, V, q);
    case 3: return launch_reduce<T, 3, RED_DOT>(ctx, n, V, q);
    case 4: return launch_reduce<T, 4, RED_DOT>(ctx, n, V, q);
    case 5: return launch_reduce<T, 5, RED_DOT>(ctx, n, V, q);
    case 6: return launch_reduce<T, 6, RED_DOT>(ctx, n, V, q);
    case 7: return launch_reduce<T, 7, RED_DOT>(ctx, n, V, q);
    case 8: return launch_reduce<T, 8, RED_DOT>(ctx, n, V, q);
  }
  return CTS_EINVAL;
}

template <typename T, int EW, int OP>
int launch_reduce_ew(cts_ctx* ctx, size_t n, void* y, const void* x, const void* z, double a, double b) {
  Plan p;
  CTS_TRY(make_plan(ctx, n, 1, &p));
  const size_t off = p.off;
  CTS_COUNT_BYTES(ctx, 2 + (EW == EW_AXPBY ? 1 : 0) + (z ? 1 : 0), n, sizeof(T));  // producer + reduction in one pass
  const bool aligned = cts_aligned16((const T*)y + off) && cts_aligned16((const T*)x + off) && (!z || cts_aligned16((const T*)z + off));
  if (aligned)
    reduce_ew_kernel<T, EW, OP, true><<<(unsigned)p.nchunks, kReduceThreads, 0, ctx->stream>>>((T*)y, (const T*)x, (const T*)z, a, b, n, off, p.n, p.f);
  else
    reduce_ew_kernel<T, EW, OP, false><<<(unsigned)p.nchunks, kReduceThreads, 0, ctx->stream>>>((T*)y, (const T*)x, (const T*)z, a, b, n, off, p.n, p.f);
  CTS_LAUNCH_CHECK(ctx);
  if (p.exchange) {
    CTS_TRY(cts_allreduce_sum_f64_oop(ctx, ctx->d_partials, ctx->d_partials_g, p.f.nc_global));
    reduce_fold_kernel<1><<<1, kReduceThreads, 0, ctx->stream>>>(ctx->d_partials_g, p.f.nc_global, ctx->d_result);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

}  // namespace

extern "C" {

int cts_axpby_dot(cts_ctx* ctx, cts_dtype ft, size_t n, double a, const void* x, double b, void* y, const void* z,
                  double* host_out) {
  CTS_CHECK_CTX(ctx);
  if (!x || !y || !host_out || x == y) return cts_fail(ctx, CTS_EINVAL, "cts_axpby_dot: bad argument");
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
      CTS_TRY(multi_dot_t<double>(ctx, n, kk, V + j0, q));
    else
      CTS_TRY(multi_dot_t<float>(ctx, n, kk, V + j0, q));
    CTS_TRY(finish(ctx, kk, host_out + j0));
  }
  return CTS_OK;
}

int cts_set_reduction_window(cts_ctx* ctx, size_t offset, size_t count) {
  CTS_CHECK_CTX(ctx);
  ctx->red_window = count != 0;
  ctx->red_offset = offset;
  ctx->red_count = count;
  ctx->red_layout = false;
  ctx->red_chunk = 0;
  return CTS_OK;
}

int cts_set_reduction_layout(cts_ctx* ctx, size_t local_offset, size_t count, size_t global_offset, size_t global_count,
                             size_t chunk) {
  CTS_CHECK_CTX(ctx);
  if (chunk == 0) chunk = kReduceChunk;
  if (count == 0 || chunk < 4 || (chunk & (chunk - 1)) != 0 || chunk > (1u << 30))
    return cts_fail(ctx, CTS_EINVAL, "cts_set_reduction_layout: chunk must be a power of two >= 4, count > 0");
  if (global_offset + count > global_count || global_offset % chunk != 0 ||
      (count % chunk != 0 && global_offset + count != global_count))
    return cts_fail(ctx, CTS_EINVAL, "cts_set_reduction_layout: the owned range must consist of whole chunks of the global domain");
  ctx->red_window = true;
  ctx->red_offset = local_offset;
  ctx->red_count = count;
  ctx->red_layout = true;
  ctx->red_g0 = global_offset;
  ctx->red_G = global_count;
  ctx->red_chunk = (unsigned)chunk;
  // partial sums of chunks owned by other ranks must read as zero in the exchange
  if (ctx->d_partials) CTS_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, ctx->partials_cap * sizeof(double), ctx->stream));
  return CTS_OK;
}

int cts_reduction_global_length(cts_ctx* ctx, size_t n, size_t* out) {
  CTS_CHECK_CTX(ctx);
  if (!out) return CTS_EINVAL;
  *out = cts_global_length(ctx, n);
  return CTS_OK;
}

}  // extern "C"
