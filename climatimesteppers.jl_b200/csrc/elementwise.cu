// Element-wise state-update kernels (K1-K11 of SURVEY §2c).
//
// All of these are pure HBM streaming: (1+nnz) reads + 1 write per DOF, <= 2 flop/byte, so the design
// rules are the memory ones: 128-bit coalesced accesses, every operand's load issued before the first
// use (fully unrolled operand loop -> NT*UNROLL independent 16-byte requests in flight per thread),
// grids sized from the problem (>= several waves of 148 SMs at the 2^27-DOF target size), no shared
// memory.  Arithmetic uses __dmul_rn/__dadd_rn so nothing is contracted into FMAs: the results are
// bit-identical to the reference's broadcast evaluation order.
#include <algorithm>
#include "cts_internal.h"
#include "ew_device.cuh"

using namespace cts_ew;

namespace {

template <typename T, typename C, int NT, int UNROLL>
__global__ void __launch_bounds__(256) axpy_n_vec_kernel(const AxpyArgs<NT> a, size_t nvec) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  const size_t tile = (size_t)blockIdx.x * (blockDim.x * UNROLL);
  V vb[UNROLL];
  V vt[NTS][UNROLL];
  bool ok[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    ok[u] = i < nvec;
    if (ok[u]) {
      vb[u] = reinterpret_cast<const V*>(a.base)[i];
#pragma unroll
      for (int k = 0; k < NT; ++k) vt[k][u] = reinterpret_cast<const V*>(a.terms[k])[i];
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (!ok[u]) continue;
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    T eb[VN], eo[VN];
    vec_unpack<T>(vb[u], eb);
    T et[NTS][VN];
#pragma unroll
    for (int k = 0; k < NT; ++k) vec_unpack<T>(vt[k][u], et[k]);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      T x[NTS];
#pragma unroll
      for (int k = 0; k < NT; ++k) x[k] = et[k][e];
      eo[e] = axpy_combine<T, C, NT>(eb[e], x, a);
    }
    V vo = vec_pack<T>(eo);
    reinterpret_cast<V*>(a.out)[i] = vo;
    if (a.out2) reinterpret_cast<V*>(a.out2)[i] = vo;
  }
}

// scalar variant for unaligned pointers and tails
template <typename T, typename C, int NT>
__global__ void __launch_bounds__(256) axpy_n_scalar_kernel(const AxpyArgs<NT> a, size_t begin, size_t n) {
  constexpr int NTS = NT > 0 ? NT : 1;
  size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T b = reinterpret_cast<const T*>(a.base)[i];
  T x[NTS];
#pragma unroll
  for (int k = 0; k < NT; ++k) x[k] = reinterpret_cast<const T*>(a.terms[k])[i];
  T o = axpy_combine<T, C, NT>(b, x, a);
  reinterpret_cast<T*>(a.out)[i] = o;
  if (a.out2) reinterpret_cast<T*>(a.out2)[i] = o;
}

template <typename T, typename C, int NT>
int launch_axpy(cts_ctx* ctx, size_t n, void* out, void* out2, double c0, const void* base, int n1, const double* coef,
                const void* const* terms) {
  AxpyArgs<NT> a;
  a.base = base;
  a.out = out;
  a.out2 = out2;
  a.c0 = c0;
  a.n1 = n1;
  a.scale_base = (c0 != 1.0);
  bool aligned = cts_aligned16(base) && cts_aligned16(out) && (!out2 || cts_aligned16(out2));
  for (int k = 0; k < NT; ++k) {
    a.terms[k] = terms[k];
    a.coef[k] = coef[k];
    aligned = aligned && cts_aligned16(terms[k]);
  }
  constexpr int VN = Vec<T>::N;
  CTS_COUNT_BYTES(ctx, 1 + NT + 1 + (out2 ? 1 : 0), n, sizeof(T));  // K1/K2: base + terms in, out (+ out2)
  // more independent requests per thread when there are few operands
  constexpr int UNROLL = (NT <= 1) ? 4 : (NT <= 4 ? 2 : 1);
  size_t nvec = aligned ? n / VN : 0;
  if (nvec > 0) {
    size_t per_block = (size_t)256 * UNROLL;
    size_t blocks = (nvec + per_block - 1) / per_block;
    axpy_n_vec_kernel<T, C, NT, UNROLL><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, nvec);
    CTS_LAUNCH_CHECK(ctx);
  }
  size_t done = nvec * VN;
  if (done < n) {
    size_t rem = n - done;
    size_t blocks = (rem + 255) / 256;
    axpy_n_scalar_kernel<T, C, NT><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, done, n);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

template <typename T, typename C>
int dispatch_axpy(cts_ctx* ctx, size_t n, void* out, void* out2, double c0, const void* base, int n1, int nt,
                  const double* coef, const void* const* terms) {
  switch (nt) {
#define CASE(N) \
  case N:       \
    return launch_axpy<T, C, N>(ctx, n, out, out2, c0, base, n1, coef, terms);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
        CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default:
      return cts_fail(ctx, CTS_EINVAL, "cts_axpy_n: more than CTS_MAX_TERMS operands");
  }
}

// ---------------------------------------------------------------------------------------------
// K11: one Shu-Osher stage assembly of the IMEX-SSPRK scheme in a single pass (imex_ssprk.jl:112-122,162-169; the
// reference runs these as 3-4 separate broadcasts):
//   e1    = U_exp + dt*T_exp                 (rounded to T, as the stand-alone pass stores it)
//   blend = (1-β)*u + β*e1                   -> U_exp   (rounded to T unless this is the final update)
//   U     = blend + Σ coef_k*T_imp_k         -> U, temp
// Every operation is individually rounded in the same order as the chained passes, so results are bit-identical.
template <int NT>
struct SspArgs {
  AxpyArgs<NT> ax;  // coef/terms of the implicit increment; ax.out = U (nullable), ax.out2 = temp (nullable)
  const void* xexp;
  const void* texp;  // nullable
  const void* u;
  void* out_exp;     // nullable
  double dt, beta, omb;
  int dt_native, final_update;
};

template <typename T, typename C, int NT>
__device__ __forceinline__ void ssp_combine(T xe, T te, T uu, const T (&x)[NT > 0 ? NT : 1], const SspArgs<NT>& a, T& o_exp, T& o_u) {
  T e1 = xe;
  if (a.texp) {
    if (a.dt_native)
      e1 = (T)op_add((float)xe, op_mul((float)a.dt, (float)te));
    else
      e1 = (T)op_add((double)xe, op_mul(a.dt, (double)te));
  }
  // β is Float64 (imex_ssprk.jl:48-49), so the blend is a Float64 broadcast for every state type; the implicit increment
  // is folded in C = promote(typeof(dt), eltype(tableau)) and promoted when it meets the blend
  double blend = op_add(op_mul(a.omb, (double)uu), op_mul(a.beta, (double)e1));
  o_exp = (T)blend;
  if (!a.final_update) blend = (double)o_exp;
  if (NT > 0) {
    C g = 0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      C p = op_mul((C)a.ax.coef[k], (C)x[k]);
      g = (k == 0) ? p : op_add(g, p);
    }
    if (!a.final_update && sizeof(C) == 4)
      o_u = (T)op_add((C)o_exp, g);   // `@. U = U_exp + inc_imp` (:169): both Float32 -> Float32 broadcast
    else
      o_u = (T)op_add(blend, (double)g);
  } else {
    o_u = o_exp;
  }
}

template <typename T, typename C, int NT, int UNROLL>
__global__ void __launch_bounds__(256) ssprk_stage_vec_kernel(const SspArgs<NT> a, size_t nvec) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  const size_t tile = (size_t)blockIdx.x * (blockDim.x * UNROLL);
  V vx[UNROLL], vT[UNROLL], vu[UNROLL], vt[NTS][UNROLL];
  bool ok[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    ok[u] = i < nvec;
    if (ok[u]) {
      vx[u] = reinterpret_cast<const V*>(a.xexp)[i];
      vT[u] = a.texp ? reinterpret_cast<const V*>(a.texp)[i] : vx[u];
      vu[u] = reinterpret_cast<const V*>(a.u)[i];
#pragma unroll
      for (int k = 0; k < NT; ++k) vt[k][u] = reinterpret_cast<const V*>(a.ax.terms[k])[i];
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (!ok[u]) continue;
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    T ex[VN], eT[VN], eu[VN], oe[VN], oU[VN], et[NTS][VN];
    vec_unpack<T>(vx[u], ex);
    vec_unpack<T>(vT[u], eT);
    vec_unpack<T>(vu[u], eu);
#pragma unroll
    for (int k = 0; k < NT; ++k) vec_unpack<T>(vt[k][u], et[k]);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      T x[NTS];
#pragma unroll
      for (int k = 0; k < NT; ++k) x[k] = et[k][e];
      ssp_combine<T, C, NT>(ex[e], eT[e], eu[e], x, a, oe[e], oU[e]);
    }
    if (a.out_exp) reinterpret_cast<V*>(a.out_exp)[i] = vec_pack<T>(oe);
    V vo = vec_pack<T>(oU);
    if (a.ax.out) reinterpret_cast<V*>(a.ax.out)[i] = vo;
    if (a.ax.out2) reinterpret_cast<V*>(a.ax.out2)[i] = vo;
  }
}

template <typename T, typename C, int NT>
__global__ void __launch_bounds__(256) ssprk_stage_scalar_kernel(const SspArgs<NT> a, size_t begin, size_t n) {
  constexpr int NTS = NT > 0 ? NT : 1;
  size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T xe = reinterpret_cast<const T*>(a.xexp)[i];
  T te = a.texp ? reinterpret_cast<const T*>(a.texp)[i] : xe;
  T uu = reinterpret_cast<const T*>(a.u)[i];
  T x[NTS];
#pragma unroll
  for (int k = 0; k < NT; ++k) x[k] = reinterpret_cast<const T*>(a.ax.terms[k])[i];
  T oe, oU;
  ssp_combine<T, C, NT>(xe, te, uu, x, a, oe, oU);
  if (a.out_exp) reinterpret_cast<T*>(a.out_exp)[i] = oe;
  if (a.ax.out) reinterpret_cast<T*>(a.ax.out)[i] = oU;
  if (a.ax.out2) reinterpret_cast<T*>(a.ax.out2)[i] = oU;
}

template <typename T, typename C, int NT>
int launch_ssprk(cts_ctx* ctx, size_t n, SspArgs<NT> a, const double* coef, const void* const* terms) {
  bool aligned = cts_aligned16(a.xexp) && cts_aligned16(a.u) && (!a.texp || cts_aligned16(a.texp)) &&
                 (!a.out_exp || cts_aligned16(a.out_exp)) && (!a.ax.out || cts_aligned16(a.ax.out)) &&
                 (!a.ax.out2 || cts_aligned16(a.ax.out2));
  for (int k = 0; k < NT; ++k) {
    a.ax.terms[k] = terms[k];
    a.ax.coef[k] = coef[k];
    aligned = aligned && cts_aligned16(terms[k]);
  }
  constexpr int VN = Vec<T>::N;
  constexpr int UNROLL = (NT <= 2) ? 2 : 1;
  CTS_COUNT_BYTES(ctx, 2 + (a.texp ? 1 : 0) + NT + (a.out_exp ? 1 : 0) + (a.ax.out ? 1 : 0) + (a.ax.out2 ? 1 : 0), n, sizeof(T));  // K11
  size_t nvec = aligned ? n / VN : 0;
  if (nvec > 0) {
    size_t per_block = (size_t)256 * UNROLL;
    ssprk_stage_vec_kernel<T, C, NT, UNROLL><<<(unsigned)((nvec + per_block - 1) / per_block), 256, 0, ctx->stream>>>(a, nvec);
    CTS_LAUNCH_CHECK(ctx);
  }
  size_t done = nvec * VN;
  if (done < n) {
    ssprk_stage_scalar_kernel<T, C, NT><<<(unsigned)((n - done + 255) / 256), 256, 0, ctx->stream>>>(a, done, n);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

template <typename T, typename C>
int dispatch_ssprk(cts_ctx* ctx, size_t n, int nt, void* U_exp_out, void* U_out, void* temp_out, const void* U_exp,
                   const void* T_exp, const void* u, double dt, int dt_native, double beta, double omb,
                   const double* coef, const void* const* terms, int final_update) {
  switch (nt) {
#define CASE(N)                                                                                  \
  case N: {                                                                                      \
    SspArgs<N> a{};                                                                              \
    a.ax.out = U_out; a.ax.out2 = temp_out; a.ax.base = nullptr; a.ax.c0 = 1.0; a.ax.n1 = N; a.ax.scale_base = 0; \
    a.xexp = U_exp; a.texp = T_exp; a.u = u; a.out_exp = U_exp_out;                              \
    a.dt = dt; a.beta = beta; a.omb = omb; a.dt_native = dt_native; a.final_update = final_update; \
    return launch_ssprk<T, C, N>(ctx, n, a, coef, terms);                                        \
  }
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default:
      return cts_fail(ctx, CTS_EINVAL, "cts_ssprk_stage: more than 8 implicit terms");
  }
}

// ---------------------------------------------------------------------------------------------
// Chained accumulation  out = ((((base + c1*x1) + c2*x2) + ...) + cn*xn) * scale,  every operation rounded in Float64
// and the result rounded to T once: what a sequence of `y .+= c .* x` broadcasts followed by `y .*= s` computes when the
// intermediate array is Float64 -- or T == double.  (For a Float32 state the reference rounds to Float32 after each
// pass; that case is handled by `round_each`.)  base == nullptr starts from a literal zero (`fill!(y, 0)`).
// Used by the Rosenbrock step (rosenbrock.jl:178-214,231-233).
template <typename T, int NT, int UNROLL>
__global__ void __launch_bounds__(256) axpy_chain_vec_kernel(const AxpyArgs<NT> a, double scale, int has_scale, int round_each,
                                                             size_t nvec) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  const size_t tile = (size_t)blockIdx.x * (blockDim.x * UNROLL);
  V vb[UNROLL], vt[NTS][UNROLL];
  bool ok[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    ok[u] = i < nvec;
    if (ok[u]) {
      if (a.base) vb[u] = reinterpret_cast<const V*>(a.base)[i];
#pragma unroll
      for (int k = 0; k < NT; ++k) vt[k][u] = reinterpret_cast<const V*>(a.terms[k])[i];
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (!ok[u]) continue;
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    T eb[VN], eo[VN], et[NTS][VN];
    if (a.base) vec_unpack<T>(vb[u], eb);
#pragma unroll
    for (int k = 0; k < NT; ++k) vec_unpack<T>(vt[k][u], et[k]);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      double r = a.base ? (double)eb[e] : 0.0;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        r = op_add(r, op_mul(a.coef[k], (double)et[k][e]));
        if (round_each) r = (double)(T)r;
      }
      if (has_scale) r = op_mul(r, scale);
      eo[e] = (T)r;
    }
    reinterpret_cast<V*>(a.out)[i] = vec_pack<T>(eo);
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(256) axpy_chain_scalar_kernel(const AxpyArgs<NT> a, double scale, int has_scale,
                                                                int round_each, size_t begin, size_t n) {
  size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = a.base ? (double)reinterpret_cast<const T*>(a.base)[i] : 0.0;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    r = op_add(r, op_mul(a.coef[k], (double)reinterpret_cast<const T*>(a.terms[k])[i]));
    if (round_each) r = (double)(T)r;
  }
  if (has_scale) r = op_mul(r, scale);
  reinterpret_cast<T*>(a.out)[i] = (T)r;
}

template <typename T, int NT>
int launch_chain(cts_ctx* ctx, size_t n, void* out, const void* base, const double* coef, const void* const* terms,
                 double scale, int has_scale) {
  AxpyArgs<NT> a{};
  a.base = base;
  a.out = out;
  bool aligned = (!base || cts_aligned16(base)) && cts_aligned16(out);
  for (int k = 0; k < NT; ++k) {
    a.terms[k] = terms[k];
    a.coef[k] = coef[k];
    aligned = aligned && cts_aligned16(terms[k]);
  }
  constexpr int VN = Vec<T>::N;
  constexpr int UNROLL = (NT <= 1) ? 4 : (NT <= 4 ? 2 : 1);
  CTS_COUNT_BYTES(ctx, (base ? 1 : 0) + NT + 1, n, sizeof(T));
  const int round_each = sizeof(T) == 4;
  size_t nvec = aligned ? n / VN : 0;
  if (nvec > 0) {
    size_t per_block = (size_t)256 * UNROLL;
    axpy_chain_vec_kernel<T, NT, UNROLL><<<(unsigned)((nvec + per_block - 1) / per_block), 256, 0, ctx->stream>>>(a, scale, has_scale, round_each, nvec);
    CTS_LAUNCH_CHECK(ctx);
  }
  size_t done = nvec * VN;
  if (done < n) {
    axpy_chain_scalar_kernel<T, NT><<<(unsigned)((n - done + 255) / 256), 256, 0, ctx->stream>>>(a, scale, has_scale, round_each, done, n);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

template <typename T>
int dispatch_chain(cts_ctx* ctx, size_t n, int nt, void* out, const void* base, const double* coef,
                   const void* const* terms, double scale, int has_scale) {
  switch (nt) {
#define CASE(N) \
  case N:       \
    return launch_chain<T, N>(ctx, n, out, base, coef, terms, scale, has_scale);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
    default:
      return cts_fail(ctx, CTS_EINVAL, "cts_axpy_chain: more than 12 operands");
  }
}

// ---------------------------------------------------------------------------------------------
// Component-wise ConvergenceCondition trees (convergence_checker.jl:62-73,97-105): the tree travels by value in postfix
// order; node k's per-element cache is caches[k].
struct CondTree {
  int n;
  cts_cond_node node[CTS_MAX_COND_NODES];
  void* cache[CTS_MAX_COND_NODES];
};

template <typename T>
__device__ __forceinline__ T cond_pow(T err, double order) {
  // Julia's x^p for the Integer orders 1, 2, 3 is x, x*x, x*x*x; other orders go through pow
  if (order == 1.0) return err;
  if (order == 2.0) return op_mul(err, err);
  if (order == 3.0) return op_mul(op_mul(err, err), err);
  return (T)pow((double)err, order);
}

template <typename T>
__device__ __forceinline__ bool cond_eval(const CondTree& t, size_t i, T val, T err, int iter) {
  bool st[CTS_MAX_COND_NODES];
  int sp = 0;
  for (int k = 0; k < t.n; ++k) {
    const cts_cond_node& nd = t.node[k];
    bool r;
    switch (nd.kind) {
      case CTS_COND_MAX_ERROR: r = (double)err <= nd.p0; break;
      case CTS_COND_MAX_REL_ERROR: r = (double)err <= __dmul_rn(nd.p0, (double)val); break;
      case CTS_COND_MAX_ERROR_REDUCTION: r = iter >= 1 && err <= reinterpret_cast<const T*>(t.cache[k])[i]; break;
      case CTS_COND_MIN_RATE: r = iter >= 1 && err >= reinterpret_cast<const T*>(t.cache[k])[i]; break;
      default: {
        r = nd.kind == CTS_COND_ALL;
        for (int c = 0; c < nd.nchildren; ++c) {
          const bool v = st[--sp];
          r = nd.kind == CTS_COND_ALL ? (r && v) : (r || v);
        }
      }
    }
    st[sp++] = r;
  }
  return st[sp - 1];
}

template <typename T>
__global__ void __launch_bounds__(256) component_check_kernel(const CondTree t, const T* __restrict__ val,
                                                              const T* __restrict__ err, size_t n, int iter,
                                                              double* __restrict__ nfail) {
  unsigned local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    local += cond_eval<T>(t, i, (T)fabs((double)val[i]), (T)fabs((double)err[i]), iter) ? 0u : 1u;
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(nfail, (double)local);  // a count: any order gives the same sum
}

template <typename T>
__global__ void __launch_bounds__(256) component_update_kernel(const CondTree t, const T* __restrict__ val,
                                                               const T* __restrict__ err, size_t n, int iter) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const T e = (T)fabs((double)err[i]);
    for (int k = 0; k < t.n; ++k) {
      const cts_cond_node& nd = t.node[k];
      if (nd.kind == CTS_COND_MAX_ERROR_REDUCTION && iter == 0)
        reinterpret_cast<T*>(t.cache[k])[i] = (T)__dmul_rn(nd.p0, (double)e);
      else if (nd.kind == CTS_COND_MIN_RATE)
        reinterpret_cast<T*>(t.cache[k])[i] = (T)__dmul_rn(nd.p0, (double)cond_pow<T>(e, nd.p1));
    }
  }
}

int make_tree(cts_ctx* ctx, int nnodes, const cts_cond_node* nodes, void* const* caches, CondTree& t) {
  if (nnodes < 1 || nnodes > CTS_MAX_COND_NODES || !nodes) return cts_fail(ctx, CTS_EINVAL, "condition tree: bad node count");
  t.n = nnodes;
  for (int k = 0; k < nnodes; ++k) {
    t.node[k] = nodes[k];
    t.cache[k] = caches ? caches[k] : nullptr;
    const bool needs = nodes[k].kind == CTS_COND_MAX_ERROR_REDUCTION || nodes[k].kind == CTS_COND_MIN_RATE;
    if (needs && !t.cache[k]) return cts_fail(ctx, CTS_EINVAL, "condition tree: missing cache array");
  }
  return CTS_OK;
}

// ---------------------------------------------------------------------------------------------
// generic small element-wise kernels: out = f(a, b, c)
// ---------------------------------------------------------------------------------------------
template <typename T, int NIN, typename F>
__global__ void __launch_bounds__(256) ew_vec_kernel(T* out, const T* a, const T* b, const T* c, size_t nvec, F f) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int UNROLL = 2;
  const size_t tile = (size_t)blockIdx.x * (blockDim.x * UNROLL);
  V va[UNROLL], vb[UNROLL], vc[UNROLL];
  bool ok[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    ok[u] = i < nvec;
    if (ok[u]) {
      va[u] = reinterpret_cast<const V*>(a)[i];
      if (NIN > 1) vb[u] = reinterpret_cast<const V*>(b)[i];
      if (NIN > 2) vc[u] = reinterpret_cast<const V*>(c)[i];
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (!ok[u]) continue;
    size_t i = tile + (size_t)u * blockDim.x + threadIdx.x;
    T ea[VN], eb[VN], ec[VN], eo[VN];
    vec_unpack<T>(va[u], ea);
    if (NIN > 1) vec_unpack<T>(vb[u], eb);
    if (NIN > 2) vec_unpack<T>(vc[u], ec);
#pragma unroll
    for (int e = 0; e < VN; ++e) eo[e] = f(ea[e], NIN > 1 ? eb[e] : T(0), NIN > 2 ? ec[e] : T(0));
    reinterpret_cast<V*>(out)[i] = vec_pack<T>(eo);
  }
}

template <typename T, int NIN, typename F>
__global__ void __launch_bounds__(256) ew_scalar_kernel(T* out, const T* a, const T* b, const T* c, size_t begin,
                                                         size_t n, F f) {
  size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = f(a[i], NIN > 1 ? b[i] : T(0), NIN > 2 ? c[i] : T(0));
}

template <typename T, int NIN, typename F>
int launch_ew(cts_ctx* ctx, size_t n, void* out, const void* a, const void* b, const void* c, F f) {
  if (n == 0) return CTS_OK;
  constexpr int VN = Vec<T>::N;
  CTS_COUNT_BYTES(ctx, NIN + 1, n, sizeof(T));  // K3 4w, K5/K6/K7/K8/K10 3w, ...
  bool aligned = cts_aligned16(out) && cts_aligned16(a) && (NIN < 2 || cts_aligned16(b)) && (NIN < 3 || cts_aligned16(c));
  size_t nvec = aligned ? n / VN : 0;
  if (nvec > 0) {
    size_t per_block = 256 * 2;
    size_t blocks = (nvec + per_block - 1) / per_block;
    ew_vec_kernel<T, NIN, F><<<(unsigned)blocks, 256, 0, ctx->stream>>>((T*)out, (const T*)a, (const T*)b, (const T*)c, nvec, f);
    CTS_LAUNCH_CHECK(ctx);
  }
  size_t done = nvec * VN;
  if (done < n) {
    size_t blocks = (n - done + 255) / 256;
    ew_scalar_kernel<T, NIN, F><<<(unsigned)blocks, 256, 0, ctx->stream>>>((T*)out, (const T*)a, (const T*)b, (const T*)c, done, n, f);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}

// K5 + K6 of a one-iteration Newton solve in one pass (newtons_method.jl:651 and imex_ark.jl:269):
//   U = x - Δx ;  T_imp = (U - x) / dtγ          (x == temp bit for bit: nothing has touched U since `temp = U`)
// 4w B/DOF instead of 3w + 3w.  NATIVE: dtγ has the state precision (Float32 dt and a downcast tableau).
template <typename T, bool NATIVE>
__global__ void __launch_bounds__(256) newton_update_timp_kernel(T* __restrict__ U, T* __restrict__ Timp,
                                                                 const T* __restrict__ dx, double dtg, size_t nvec,
                                                                 size_t n) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto one = [&](T x, T d, T& u, T& t) {
    u = op_sub(x, d);
    const T diff = op_sub(u, x);
    t = NATIVE ? op_div(diff, (T)dtg) : (T)__ddiv_rn((double)diff, dtg);
  };
  if (i < nvec) {
    T ex[VN], ed[VN], eu[VN], et[VN];
    vec_unpack<T>(reinterpret_cast<const V*>(U)[i], ex);
    vec_unpack<T>(reinterpret_cast<const V*>(dx)[i], ed);
#pragma unroll
    for (int e = 0; e < VN; ++e) one(ex[e], ed[e], eu[e], et[e]);
    reinterpret_cast<V*>(U)[i] = vec_pack<T>(eu);
    reinterpret_cast<V*>(Timp)[i] = vec_pack<T>(et);
  }
  // scalar tail (and the whole range when the operands are not 16-byte aligned: nvec == 0)
  const size_t j = nvec * VN + i;
  if (j < n) {
    T u, t;
    one(U[j], dx[j], u, t);
    U[j] = u;
    Timp[j] = t;
  }
}

// functors ------------------------------------------------------------------------------------
template <typename T>
struct ResidualF {  // K3: (U0 + dtγ*r) - U'   evaluated in Float64 (dtγ is Float64, imex_ark.jl:122,297)
  double dtg;
  __device__ T operator()(T r, T U0, T Up) const {
    return (T)__dsub_rn(__dadd_rn((double)U0, __dmul_rn(dtg, (double)r)), (double)Up);
  }
};
template <typename T>
struct ResidualNativeF {  // K3 with dtγ in the state precision (Float32 dt and a downcast tableau): a native broadcast
  T dtg;
  __device__ T operator()(T r, T U0, T Up) const { return op_sub(op_add(U0, op_mul(dtg, r)), Up); }
};
template <typename T>
struct TimpNativeF {  // K6 with dtγ in the state precision
  T dtg;
  __device__ T operator()(T U, T temp, T) const { return op_div(op_sub(U, temp), dtg); }
};
template <typename T>
struct SubF {  // K5: x - dx in the state precision
  __device__ T operator()(T x, T dx, T) const { return op_sub(x, dx); }
};
template <typename T>
struct TimpF {  // K6: (U - temp) / dtγ : subtraction in the state precision, division in Float64
  double dtg;
  __device__ T operator()(T U, T temp, T) const { return (T)__ddiv_rn((double)op_sub(U, temp), dtg); }
};
template <typename T>
struct PerturbF {  // K7: x + ε*dx  (ε has the state eltype, newtons_method.jl:177-178)
  T eps;
  __device__ T operator()(T x, T dx, T) const { return op_add(x, op_mul(eps, dx)); }
};
template <typename T>
struct QuotientF {  // K8: (f2 - f)/ε
  T eps;
  __device__ T operator()(T f2, T f, T) const { return op_div(op_sub(f2, f), eps); }
};
template <typename T, typename C>
struct LsrkF {  // K10: u + (dt*B)*du
  C dtB;
  __device__ T operator()(T u, T du, T) const { return (T)op_add((C)u, op_mul(dtB, (C)du)); }
};
template <typename T>
struct AxpbyF {  // y = a*x + b*y in Float64
  double a, b;
  __device__ T operator()(T x, T y, T) const { return (T)__dadd_rn(__dmul_rn(a, (double)x), __dmul_rn(b, (double)y)); }
};
template <typename T>
struct ScaleF {  // y = a*x in Float64
  double a;
  __device__ T operator()(T x, T, T) const { return (T)__dmul_rn(a, (double)x); }
};

}  // namespace

extern "C" {

int cts_axpy_n(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, void* out2, double c0, const void* base, int n_grp1,
               int n_grp2, const double* coef, const void* const* terms, int compute_native) {
  CTS_CHECK_CTX(ctx);
  if (n == 0) return CTS_OK;
  if (!out || !base || n_grp1 < 0 || n_grp2 < 0) return cts_fail(ctx, CTS_EINVAL, "cts_axpy_n: bad arguments");
  int nt = n_grp1 + n_grp2;
  if (nt > 0 && (!coef || !terms)) return cts_fail(ctx, CTS_EINVAL, "cts_axpy_n: null coef/terms");
  if (ft == CTS_F64) return dispatch_axpy<double, double>(ctx, n, out, out2, c0, base, n_grp1, nt, coef, terms);
  if (compute_native) return dispatch_axpy<float, float>(ctx, n, out, out2, c0, base, n_grp1, nt, coef, terms);
  return dispatch_axpy<float, double>(ctx, n, out, out2, c0, base, n_grp1, nt, coef, terms);
}

int cts_ssprk_stage(cts_ctx* ctx, cts_dtype ft, size_t n, void* U_exp_out, void* U_out, void* temp_out, const void* U_exp,
                    const void* T_exp, const void* u, double dt, int dt_native, double beta, double one_minus_beta,
                    int n_imp, const double* coef, const void* const* terms, int compute_native, int final_update) {
  CTS_CHECK_CTX(ctx);
  if (n == 0) return CTS_OK;
  if (!U_exp || !u || n_imp < 0 || (n_imp > 0 && (!coef || !terms)) || (!U_exp_out && !U_out))
    return cts_fail(ctx, CTS_EINVAL, "cts_ssprk_stage: bad arguments");
  if (ft == CTS_F64)
    return dispatch_ssprk<double, double>(ctx, n, n_imp, U_exp_out, U_out, temp_out, U_exp, T_exp, u, dt, 0, beta,
                                          one_minus_beta, coef, terms, final_update);
  if (compute_native)
    return dispatch_ssprk<float, float>(ctx, n, n_imp, U_exp_out, U_out, temp_out, U_exp, T_exp, u, dt, dt_native, beta,
                                        one_minus_beta, coef, terms, final_update);
  return dispatch_ssprk<float, double>(ctx, n, n_imp, U_exp_out, U_out, temp_out, U_exp, T_exp, u, dt, dt_native, beta,
                                       one_minus_beta, coef, terms, final_update);
}

int cts_axpy_chain(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, const void* base, int nterms, const double* coef,
                   const void* const* terms, double scale, int has_scale) {
  CTS_CHECK_CTX(ctx);
  if (n == 0) return CTS_OK;
  if (!out || nterms < 0 || (nterms > 0 && (!coef || !terms))) return cts_fail(ctx, CTS_EINVAL, "cts_axpy_chain: bad arguments");
  if (ft == CTS_F64) return dispatch_chain<double>(ctx, n, nterms, out, base, coef, terms, scale, has_scale);
  return dispatch_chain<float>(ctx, n, nterms, out, base, coef, terms, scale, has_scale);
}

int cts_component_check(cts_ctx* ctx, cts_dtype ft, size_t n, const void* val, const void* err, int nnodes,
                        const cts_cond_node* nodes, void* const* caches, int iter, double* dev_scratch, int* all_converged) {
  CTS_CHECK_CTX(ctx);
  if (!val || !err || !dev_scratch || !all_converged) return cts_fail(ctx, CTS_EINVAL, "cts_component_check: bad arguments");
  CondTree t;
  CTS_TRY(make_tree(ctx, nnodes, nodes, caches, t));
  CTS_CUDA(ctx, cudaMemsetAsync(dev_scratch, 0, sizeof(double), ctx->stream));
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  if (ft == CTS_F64)
    component_check_kernel<double><<<blocks, 256, 0, ctx->stream>>>(t, (const double*)val, (const double*)err, n, iter, dev_scratch);
  else
    component_check_kernel<float><<<blocks, 256, 0, ctx->stream>>>(t, (const float*)val, (const float*)err, n, iter, dev_scratch);
  CTS_LAUNCH_CHECK(ctx);
  if (ctx->nranks > 1) CTS_TRY(cts_allreduce_sum_f64(ctx, dev_scratch, 1));  // every rank must take the same branch
  double nfail = 0;
  CTS_CUDA(ctx, cudaMemcpyAsync(&nfail, dev_scratch, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *all_converged = nfail == 0.0;
  return CTS_OK;
}

int cts_component_update(cts_ctx* ctx, cts_dtype ft, size_t n, const void* val, const void* err, int nnodes,
                         const cts_cond_node* nodes, void* const* caches, int iter) {
  CTS_CHECK_CTX(ctx);
  if (!val || !err) return cts_fail(ctx, CTS_EINVAL, "cts_component_update: bad arguments");
  CondTree t;
  CTS_TRY(make_tree(ctx, nnodes, nodes, caches, t));
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  if (ft == CTS_F64)
    component_update_kernel<double><<<blocks, 256, 0, ctx->stream>>>(t, (const double*)val, (const double*)err, n, iter);
  else
    component_update_kernel<float><<<blocks, 256, 0, ctx->stream>>>(t, (const float*)val, (const float*)err, n, iter);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

int cts_newton_residual(cts_ctx* ctx, cts_dtype ft, size_t n, void* r, const void* U0, double dtgamma, const void* Up) {
  return cts_newton_residual_ex(ctx, ft, n, r, U0, dtgamma, Up, 0);
}

int cts_sub_inplace(cts_ctx* ctx, cts_dtype ft, size_t n, void* x, const void* dx) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, x, x, dx, nullptr, SubF<double>{});
  return launch_ew<float, 2>(ctx, n, x, x, dx, nullptr, SubF<float>{});
}

int cts_timp_from_stage(cts_ctx* ctx, cts_dtype ft, size_t n, void* T, const void* U, const void* temp, double dtgamma) {
  return cts_timp_from_stage_ex(ctx, ft, n, T, U, temp, dtgamma, 0);
}

int cts_jvp_perturb(cts_ctx* ctx, cts_dtype ft, size_t n, void* x2, const void* x, double eps, const void* dx) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, x2, x, dx, nullptr, PerturbF<double>{eps});
  return launch_ew<float, 2>(ctx, n, x2, x, dx, nullptr, PerturbF<float>{(float)eps});
}

int cts_jvp_quotient(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, const void* f2, const void* f, double eps) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, out, f2, f, nullptr, QuotientF<double>{eps});
  return launch_ew<float, 2>(ctx, n, out, f2, f, nullptr, QuotientF<float>{(float)eps});
}

int cts_lsrk_update(cts_ctx* ctx, cts_dtype ft, size_t n, void* u, double dtB, const void* du, int compute_native) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, u, u, du, nullptr, LsrkF<double, double>{dtB});
  if (compute_native) return launch_ew<float, 2>(ctx, n, u, u, du, nullptr, LsrkF<float, float>{(float)dtB});
  return launch_ew<float, 2>(ctx, n, u, u, du, nullptr, LsrkF<float, double>{dtB});
}

int cts_axpby(cts_ctx* ctx, cts_dtype ft, size_t n, double a, const void* x, double b, void* y) {
  CTS_CHECK_CTX(ctx);
  if (b == 0.0) {
    if (ft == CTS_F64) return launch_ew<double, 1>(ctx, n, y, x, nullptr, nullptr, ScaleF<double>{a});
    return launch_ew<float, 1>(ctx, n, y, x, nullptr, nullptr, ScaleF<float>{a});
  }
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, y, x, y, nullptr, AxpbyF<double>{a, b});
  return launch_ew<float, 2>(ctx, n, y, x, y, nullptr, AxpbyF<float>{a, b});
}

}  // extern "C"

int cts_newton_update_timp(cts_ctx* ctx, cts_dtype ft, size_t n, void* U, void* T_imp, const void* dx, double dtgamma,
                           int native) {
  CTS_CHECK_CTX(ctx);
  if (!U || !T_imp || !dx || U == T_imp) return cts_fail(ctx, CTS_EINVAL, "cts_newton_update_timp: bad argument");
  const bool aligned = cts_aligned16(U) && cts_aligned16(T_imp) && cts_aligned16(dx);
  CTS_COUNT_BYTES(ctx, 4, n, cts_sizeof(ft));
  if (ft == CTS_F64) {
    const size_t nvec = aligned ? n / 2 : 0;
    const size_t work = nvec > n - nvec * 2 ? nvec : n - nvec * 2;
    newton_update_timp_kernel<double, false><<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>((double*)U, (double*)T_imp, (const double*)dx, dtgamma, nvec, n);
  } else {
    const size_t nvec = aligned ? n / 4 : 0;
    const size_t work = nvec > n - nvec * 4 ? nvec : n - nvec * 4;
    if (native)
      newton_update_timp_kernel<float, true><<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>((float*)U, (float*)T_imp, (const float*)dx, dtgamma, nvec, n);
    else
      newton_update_timp_kernel<float, false><<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>((float*)U, (float*)T_imp, (const float*)dx, dtgamma, nvec, n);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

// K3 / K6 with the type of dtγ made explicit: native != 0 (Float32 state only) evaluates the broadcast in Float32
int cts_newton_residual_ex(cts_ctx* ctx, cts_dtype ft, size_t n, void* r, const void* U0, double dtgamma, const void* Up,
                           int native) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 3>(ctx, n, r, r, U0, Up, ResidualF<double>{dtgamma});
  if (native) return launch_ew<float, 3>(ctx, n, r, r, U0, Up, ResidualNativeF<float>{(float)dtgamma});
  return launch_ew<float, 3>(ctx, n, r, r, U0, Up, ResidualF<float>{dtgamma});
}
int cts_timp_from_stage_ex(cts_ctx* ctx, cts_dtype ft, size_t n, void* T, const void* U, const void* temp, double dtgamma,
                           int native) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 2>(ctx, n, T, U, temp, nullptr, TimpF<double>{dtgamma});
  if (native) return launch_ew<float, 2>(ctx, n, T, U, temp, nullptr, TimpNativeF<float>{(float)dtgamma});
  return launch_ew<float, 2>(ctx, n, T, U, temp, nullptr, TimpF<float>{dtgamma});
}

// ---------------------------------------------------------------------------------------------
// GMRES helpers: y = x / denom, out = sum_j coef[j] V[j]
// ---------------------------------------------------------------------------------------------
namespace {
template <typename T>
struct DivF {
  double denom;
  __device__ T operator()(T x, T, T) const { return (T)__ddiv_rn((double)x, denom); }
};

constexpr int kLincombMax = 32;
struct LincombArgs {
  const void* V[kLincombMax];
  double coef[kLincombMax];
  int k;
  int accumulate;
};
template <typename T>
__global__ void __launch_bounds__(256) lincomb_kernel(T* out, const LincombArgs a, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = a.accumulate ? (double)out[i] : 0.0;
  for (int j = 0; j < a.k; ++j) acc = __dadd_rn(acc, __dmul_rn(a.coef[j], (double)reinterpret_cast<const T*>(a.V[j])[i]));
  out[i] = (T)acc;
}
// K9 + K5 in one pass: d = sum_j coef[j]*V[j] (the fold of lincomb_kernel, rounded to T as its store would), x -= d in the
// state precision (SubF); d itself is stored only when somebody reads it afterwards.  VN elements per thread (16-byte
// vectors when the arrays allow it).
template <typename T, int VN>
__global__ void __launch_bounds__(256) lincomb_sub_kernel(T* __restrict__ x, T* __restrict__ dx, const LincombArgs a, size_t nvec) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  using V = typename cts_ew::Vec<T>::type;
  double acc[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) acc[e] = 0.0;
  for (int j = 0; j < a.k; ++j) {
    T v[VN];
    if constexpr (VN == 1)
      v[0] = reinterpret_cast<const T*>(a.V[j])[i];
    else
      cts_ew::vec_unpack<T>(reinterpret_cast<const V*>(a.V[j])[i], v);
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = __dadd_rn(acc[e], __dmul_rn(a.coef[j], (double)v[e]));
  }
  T xe[VN], de[VN];
  if constexpr (VN == 1)
    xe[0] = x[i];
  else
    cts_ew::vec_unpack<T>(reinterpret_cast<const V*>(x)[i], xe);
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    de[e] = (T)acc[e];
    xe[e] = cts_ew::op_sub(xe[e], de[e]);
  }
  if constexpr (VN == 1) {
    x[i] = xe[0];
    if (dx) dx[i] = de[0];
  } else {
    reinterpret_cast<V*>(x)[i] = cts_ew::vec_pack<T>(xe);
    if (dx) reinterpret_cast<V*>(dx)[i] = cts_ew::vec_pack<T>(de);
  }
}
}  // namespace

extern "C" int cts_lincomb_sub(cts_ctx* ctx, cts_dtype ft, size_t n, void* x, void* dx, int k, const double* coef,
                               const void* const* V) {
  CTS_CHECK_CTX(ctx);
  if (!x || k < 1 || k > kLincombMax || !coef || !V) return cts_fail(ctx, CTS_EINVAL, "cts_lincomb_sub: bad argument");
  if (n == 0) return CTS_OK;
  LincombArgs a;
  a.k = k;
  a.accumulate = 0;
  const int vn = ft == CTS_F64 ? 2 : 4;
  bool vec = (n % vn) == 0 && cts_aligned16(x) && (!dx || cts_aligned16(dx));
  for (int j = 0; j < k; ++j) {
    if (!V[j]) return cts_fail(ctx, CTS_EINVAL, "cts_lincomb_sub: null basis vector");
    a.V[j] = V[j];
    a.coef[j] = coef[j];
    vec = vec && cts_aligned16(V[j]);
  }
  CTS_COUNT_BYTES(ctx, k + 2 + (dx ? 1 : 0), n, cts_sizeof(ft));
  const size_t nvec = vec ? n / vn : n;
  const unsigned blocks = (unsigned)((nvec + 255) / 256);
  if (ft == CTS_F64) {
    if (vec) lincomb_sub_kernel<double, 2><<<blocks, 256, 0, ctx->stream>>>((double*)x, (double*)dx, a, nvec);
    else lincomb_sub_kernel<double, 1><<<blocks, 256, 0, ctx->stream>>>((double*)x, (double*)dx, a, nvec);
  } else {
    if (vec) lincomb_sub_kernel<float, 4><<<blocks, 256, 0, ctx->stream>>>((float*)x, (float*)dx, a, nvec);
    else lincomb_sub_kernel<float, 1><<<blocks, 256, 0, ctx->stream>>>((float*)x, (float*)dx, a, nvec);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

extern "C" int cts_div_scalar(cts_ctx* ctx, cts_dtype ft, size_t n, void* y, const void* x, double denom) {
  CTS_CHECK_CTX(ctx);
  if (ft == CTS_F64) return launch_ew<double, 1>(ctx, n, y, x, nullptr, nullptr, DivF<double>{denom});
  return launch_ew<float, 1>(ctx, n, y, x, nullptr, nullptr, DivF<float>{denom});
}

extern "C" int cts_lincomb(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, int k, const double* coef,
                           const void* const* V, int accumulate) {
  CTS_CHECK_CTX(ctx);
  if (!out || k < 0 || (k > 0 && (!coef || !V))) return cts_fail(ctx, CTS_EINVAL, "cts_lincomb: bad argument");
  if (n == 0) return CTS_OK;
  if (k == 0 && !accumulate) return cts_memset_zero(ctx, out, n * cts_sizeof(ft));
  size_t blocks = (n + 255) / 256;
  for (int j0 = 0; j0 < k; j0 += kLincombMax) {
    LincombArgs a;
    a.k = (k - j0 < kLincombMax) ? k - j0 : kLincombMax;
    a.accumulate = (j0 > 0) || accumulate;
    CTS_COUNT_BYTES(ctx, a.k + 1 + (a.accumulate ? 1 : 0), n, cts_sizeof(ft));
    for (int j = 0; j < a.k; ++j) {
      a.V[j] = V[j0 + j];
      a.coef[j] = coef[j0 + j];
    }
    if (ft == CTS_F64)
      lincomb_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)out, a, n);
    else
      lincomb_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)out, a, n);
    CTS_LAUNCH_CHECK(ctx);
  }
  return CTS_OK;
}
