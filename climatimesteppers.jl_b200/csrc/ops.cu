// Library-native model operators (layer L5 of the reference: the user's T_exp!/T_imp!/Wfact/dss!) for the
// synthetic benchmark problems of SURVEY §8d, and the fused kernels the solvers may substitute for a
// sequence of hook calls when they recognise these operators:
//   * cts_advection_fused_stage : K10f + K10 in one pass  (du = rhs(u) + A du ; u += dtB du), 4w B/DOF
//   * cts_column_fused_stage    : K1 + K2 + T_imp + K3 + K4 + K5 + K6 of one implicit ARK/SSPRK stage
//                                 in one pass over the column tile, (2+nnz)w [+3w for stored W] B/DOF
// Both produce results bit-identical to the un-fused hook sequence (tests/test_gpu_parity.py).
#include <algorithm>
#include <type_traits>
#include <cmath>
#include <mutex>
#include <unordered_map>
#include <cstdlib>

#include "cts_internal.h"
#include "ew_device.cuh"
#include "tile_device.cuh"
#include "fused_args.cuh"

using namespace cts_tile;
using cts_ew::AxpyArgs;
using cts_ew::axpy_combine;
using cts_ew::axpy_combine_n1;
using cts_ew::Vec;
using cts_ew::vec_pack;
using cts_ew::vec_unpack;
using cts_ew::op_add;
using cts_ew::op_mul;
using cts_fused::FusedArgs;
using cts_fused::GroupSplit;
using cts_fused::timp_linear;

// =============================================================================================
// 1-D periodic upwind advection (config C2)
//   rhs_i = (-c * (u_i - u_{i-1})) / dx            (state precision)
//   du_i  = alpha * rhs_i + beta * du_i            (IncrementingODEFunction, problems.jl:77-85)
// =============================================================================================
struct cts_op_advection {
  cts_ctx* ctx;
  cts_dtype ft;
  size_t n;
  double c, dx;
  void* edge = nullptr;  // in-place fused stage: the upwind neighbour of every CTA's first element (lazy, n/(256*VN) values)
};

namespace {

template <typename T>
__device__ __forceinline__ T adv_rhs(T uc, T ul, T negc, T dx) {
  return rdiv(rmul(negc, rsub(uc, ul)), dx);
}

// FUSED: also performs the 2N register update u_out = u + dtB*du_new (lsrk.jl:65) in the compute type C.
template <typename T, typename C, bool FUSED, bool VEC>
__global__ void __launch_bounds__(256) advection_kernel(T* du, const T* u, T* u_out, size_t n, T negc, T dx, T alpha,
                                                        T beta, C dtB) {
  using V = typename Vec<T>::type;
  constexpr int VN = VEC ? Vec<T>::N : 1;
  const size_t nv = n / VN;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  T eu[VN], ed[VN];
  if (VEC) {
    T tu[Vec<T>::N], td[Vec<T>::N];
    vec_unpack<T>(reinterpret_cast<const V*>(u)[i], tu);
    vec_unpack<T>(reinterpret_cast<const V*>(du)[i], td);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      eu[e] = tu[e];
      ed[e] = td[e];
    }
  } else {
    eu[0] = u[i];
    ed[0] = du[i];
  }
  const size_t first = i * VN;
  T left = u[first == 0 ? n - 1 : first - 1];
  T od[VN], ou[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    T rhs = adv_rhs<T>(eu[e], left, negc, dx);
    od[e] = radd(rmul(alpha, rhs), rmul(beta, ed[e]));
    if (FUSED) ou[e] = (T)cts_ew::op_add((C)eu[e], cts_ew::op_mul(dtB, (C)od[e]));
    left = eu[e];
  }
  if (VEC) {
    T td[Vec<T>::N], tu[Vec<T>::N];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      td[e] = od[e];
      tu[e] = FUSED ? ou[e] : T(0);
    }
    reinterpret_cast<V*>(du)[i] = vec_pack<T>(td);
    if (FUSED) reinterpret_cast<V*>(u_out)[i] = vec_pack<T>(tu);
  } else {
    du[i] = od[0];
    if (FUSED) u_out[i] = ou[0];
  }
}

// In-place fused stage (u_out == u): the stencil reads u[i-1], which another thread overwrites.  Inside a CTA every load is
// issued before the barrier and every store after it; across CTAs the one value that crosses the boundary (the upwind
// neighbour of the CTA's first element) is gathered into `edge` by a tiny kernel before the stage runs.  Same arithmetic
// as advection_kernel<FUSED>: bit-identical.  With it an odd stage count (LSRK54: 5) needs no two-pass stage to end in
// the caller's array: 5 x 4w = 160 B/DOF per step instead of 176.
template <typename T>
__global__ void __launch_bounds__(256) advection_edges_kernel(T* __restrict__ edge, const T* __restrict__ u, size_t n, size_t chunk,
                                                              size_t nblocks) {
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const size_t first = b * chunk;
  edge[b] = u[first == 0 ? n - 1 : first - 1];
}

template <typename T, typename C>
__global__ void __launch_bounds__(256) advection_inplace_kernel(T* du, T* u, const T* __restrict__ edge, size_t n, T negc, T dx,
                                                                T alpha, T beta, C dtB) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  const size_t nv = n / VN;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < nv;
  T eu[VN], ed[VN];
  T left = T(0);
  if (ok) {
    vec_unpack<T>(reinterpret_cast<const V*>(u)[i], eu);
    vec_unpack<T>(reinterpret_cast<const V*>(du)[i], ed);
    left = threadIdx.x == 0 ? edge[blockIdx.x] : u[i * VN - 1];
  }
  __syncthreads();  // every load of this CTA precedes every store
  if (!ok) return;
  T od[VN], ou[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    T rhs = adv_rhs<T>(eu[e], left, negc, dx);
    od[e] = radd(rmul(alpha, rhs), rmul(beta, ed[e]));
    ou[e] = (T)cts_ew::op_add((C)eu[e], cts_ew::op_mul(dtB, (C)od[e]));
    left = eu[e];
  }
  reinterpret_cast<V*>(du)[i] = vec_pack<T>(od);
  reinterpret_cast<V*>(u)[i] = vec_pack<T>(ou);
}

template <typename T, typename C>
int launch_advection_inplace(cts_op_advection* op, void* du, void* u, double A, double dtB) {
  cts_ctx* ctx = op->ctx;
  const size_t n = op->n;
  constexpr int VN = Vec<T>::N;
  const size_t chunk = (size_t)256 * VN;
  const size_t blocks = (n / VN + 255) / 256;
  if (!op->edge) CTS_TRY(cts_alloc(ctx, blocks * sizeof(T), &op->edge));
  CTS_COUNT_BYTES(ctx, 4, n, sizeof(T));  // u, du in; du, u out (the edge gather moves n/(256 VN) elements)
  advection_edges_kernel<T><<<(unsigned)((blocks + 255) / 256), 256, 0, ctx->stream>>>((T*)op->edge, (const T*)u, n, chunk, blocks);
  CTS_LAUNCH_CHECK(ctx);
  advection_inplace_kernel<T, C><<<(unsigned)blocks, 256, 0, ctx->stream>>>((T*)du, (T*)u, (const T*)op->edge, n, (T)(-op->c),
                                                                           (T)op->dx, (T)1.0, (T)A, (C)dtB);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

template <typename T, typename C, bool FUSED>
int launch_advection(cts_op_advection* op, void* du, const void* u, void* u_out, double alpha, double beta, double dtB) {
  cts_ctx* ctx = op->ctx;
  const size_t n = op->n;
  constexpr int VN = Vec<T>::N;
  bool vec = (n % VN) == 0 && cts_aligned16(du) && cts_aligned16(u) && (!FUSED || cts_aligned16(u_out));
  CTS_COUNT_BYTES(ctx, FUSED ? 4 : 3, n, sizeof(T));  // K10f (+ K10): u, du in; du (, u) out
  T negc = (T)(-op->c), dx = (T)op->dx;
  if (vec) {
    size_t blocks = (n / VN + 255) / 256;
    advection_kernel<T, C, FUSED, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>((T*)du, (const T*)u, (T*)u_out, n, negc, dx,
                                                                                   (T)alpha, (T)beta, (C)dtB);
  } else {
    size_t blocks = (n + 255) / 256;
    advection_kernel<T, C, FUSED, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>((T*)du, (const T*)u, (T*)u_out, n, negc, dx,
                                                                                    (T)alpha, (T)beta, (C)dtB);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

}  // namespace

extern "C" {

int cts_op_advection_create(cts_ctx* ctx, cts_dtype ft, size_t n, double c, double dx, cts_op_advection** out) {
  CTS_CHECK_CTX(ctx);
  if (!out || n < 2 || dx == 0.0) return cts_fail(ctx, CTS_EINVAL, "cts_op_advection_create: bad argument");
  *out = new cts_op_advection{ctx, ft, n, c, dx};
  return CTS_OK;
}
int cts_op_advection_destroy(cts_op_advection* op) {
  if (op && op->edge) cts_free(op->ctx, op->edge);
  delete op;
  return CTS_OK;
}
int cts_op_advection_incr(void* vop, void* du, const void* u, double t, double alpha, double beta) {
  auto* op = static_cast<cts_op_advection*>(vop);
  if (!op || !du || !u || du == u) return CTS_EINVAL;
  if (op->ft == CTS_F64) return launch_advection<double, double, false>(op, du, u, nullptr, alpha, beta, 0.0);
  return launch_advection<float, float, false>(op, du, u, nullptr, alpha, beta, 0.0);
}

}  // extern "C"

bool cts_advection_inplace_supported(const cts_op_advection* op, const void* u, const void* du) {
  const size_t vn = op->ft == CTS_F64 ? 2 : 4;
  return (op->n % vn) == 0 && cts_aligned16(u) && cts_aligned16(du);
}

int cts_advection_fused_stage(cts_op_advection* op, void* u_out, const void* u_in, void* du, double A, double dtB,
                              int compute_native) {
  if (!op) return CTS_EINVAL;
  if (u_out == u_in) {  // in place
    if (!cts_advection_inplace_supported(op, u_out, du)) return CTS_EUNSUPPORTED;
    if (op->ft == CTS_F64) return launch_advection_inplace<double, double>(op, du, u_out, A, dtB);
    if (compute_native) return launch_advection_inplace<float, float>(op, du, u_out, A, dtB);
    return launch_advection_inplace<float, double>(op, du, u_out, A, dtB);
  }
  if (op->ft == CTS_F64) return launch_advection<double, double, true>(op, du, u_in, u_out, 1.0, A, dtB);
  if (compute_native) return launch_advection<float, float, true>(op, du, u_in, u_out, 1.0, A, dtB);
  return launch_advection<float, double, true>(op, du, u_in, u_out, 1.0, A, dtB);
}
cts_ctx* cts_op_advection_ctx(const cts_op_advection* op) { return op->ctx; }
cts_dtype cts_op_advection_dtype(const cts_op_advection* op) { return op->ft; }
size_t cts_op_advection_n(const cts_op_advection* op) { return op->n; }

// =============================================================================================
// Column model (configs C3/C4/C5)
// =============================================================================================
struct cts_op_column {
  cts_ctx* ctx;
  cts_dtype ft;
  cts_column_desc d;
  size_t rows;   // ny_local + 2*halo
  size_t ncols;  // rows*nx
  bool matrix_free;
  void* a_col;   // K_col / dz^2 per local column (state precision), computed once at creation
};

namespace {

// per-column diffusion number a = K / dz^2 (state precision)
template <typename T>
__device__ __forceinline__ T col_a(T K, T dz) {
  return rdiv(K, rmul(dz, dz));
}
// nonlinear face diffusivity a*(1 + ubar^2), ubar = (ua+ub)/2
template <typename T>
__device__ __forceinline__ T kface(T a, T ua, T ub) {
  T ubar = rmul(T(0.5), radd(ua, ub));
  return rmul(a, radd(T(1), rmul(ubar, ubar)));
}
template <typename T>
__device__ __forceinline__ T timp_nonlinear(T a, T um, T uc, T up) {
  T fp = rmul(kface(a, uc, up), rsub(up, uc));
  T fm = rmul(kface(a, um, uc), rsub(uc, um));
  return rsub(fp, fm);
}

// T_imp: one thread per DOF vector; neighbours along the (contiguous) level axis.
template <typename T>
__global__ void column_acol_kernel(T* __restrict__ a, const T* __restrict__ Kcol, T dz, size_t ncols) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) a[c] = col_a<T>(Kcol[c], dz);
}

template <typename T, bool NONLINEAR>
__global__ void __launch_bounds__(256) column_timp_kernel(T* __restrict__ out, const T* __restrict__ u,
                                                          const T* __restrict__ acol, int nlev, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t col = i / (size_t)nlev;
  int lev = (int)(i - col * (size_t)nlev);
  T a = acol[col];
  T uc = u[i];
  T um = lev > 0 ? u[i - 1] : T(0);
  T up = lev < nlev - 1 ? u[i + 1] : T(0);
  out[i] = NONLINEAR ? timp_nonlinear<T>(a, um, uc, up) : timp_linear<T>(a, um, uc, up);
}

// Newton residual and Jacobian-free JVP of the implicit stage equation for the library's column operator, in one pass:
//   x2  = x + ε*dx                      K7 (newtons_method.jl:177-178)          [JVP only]
//   r   = T_imp(x2)                     the operator (level stencil inside the column)
//   f2  = (U0 + dtγ*r) - x2             K3 (imex_ark.jl:297)
//   out = (f2 - f)/ε                    K8 (newtons_method.jl:181)              [JVP only]
// instead of four passes over HBM (12w -> 5w per JVP, 6w -> 3w per residual).  Every operation is rounded exactly as
// in the separate kernels (PerturbF, column_timp_kernel, ResidualF, QuotientF), so results are bit-identical.
// One thread per 16-byte level vector; the two neighbouring levels across the vector boundary are scalar loads (L1).
template <typename T, bool NONLINEAR, bool JVP>
__global__ void __launch_bounds__(256) column_residual_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                              const T* __restrict__ dx, const T* __restrict__ U0,
                                                              const T* __restrict__ f, const T* __restrict__ acol,
                                                              T eps, double dtg, int nlev, size_t nvec) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  const size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= nvec) return;
  const size_t i0 = iv * VN;
  const size_t col = i0 / (size_t)nlev;
  const int l0 = (int)(i0 - col * (size_t)nlev);
  const bool has_m = l0 > 0, has_p = l0 + VN < nlev;
  const V vx = reinterpret_cast<const V*>(x)[iv];
  const V vu = reinterpret_cast<const V*>(U0)[iv];
  V vd, vf;
  T xm = T(0), xp = T(0), dm = T(0), dp = T(0);
  if (JVP) {
    vd = reinterpret_cast<const V*>(dx)[iv];
    vf = reinterpret_cast<const V*>(f)[iv];
  }
  if (has_m) {
    xm = x[i0 - 1];
    if (JVP) dm = dx[i0 - 1];
  }
  if (has_p) {
    xp = x[i0 + VN];
    if (JVP) dp = dx[i0 + VN];
  }
  const T a = acol[col];
  T ex[VN], ed[VN], eu[VN], ef[VN], x2[VN], eo[VN];
  vec_unpack<T>(vx, ex);
  vec_unpack<T>(vu, eu);
  if (JVP) {
    vec_unpack<T>(vd, ed);
    vec_unpack<T>(vf, ef);
  }
  auto perturb = [&](T xv, T dv) -> T { return JVP ? radd(xv, rmul(eps, dv)) : xv; };
#pragma unroll
  for (int e = 0; e < VN; ++e) x2[e] = perturb(ex[e], JVP ? ed[e] : T(0));
  const T x2m = has_m ? perturb(xm, dm) : T(0);
  const T x2p = has_p ? perturb(xp, dp) : T(0);
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const T um = e > 0 ? x2[e > 0 ? e - 1 : 0] : x2m;
    const T up = e + 1 < VN ? x2[e + 1 < VN ? e + 1 : e] : x2p;
    const T r = NONLINEAR ? timp_nonlinear<T>(a, um, x2[e], up) : timp_linear<T>(a, um, x2[e], up);
    const T f2 = (T)__dsub_rn(__dadd_rn((double)eu[e], __dmul_rn(dtg, (double)r)), (double)x2[e]);
    eo[e] = JVP ? rdiv(rsub(f2, ef[e]), eps) : f2;
  }
  reinterpret_cast<V*>(out)[iv] = vec_pack<T>(eo);
}

// Wfact: W = dtγ J - I   (tridiagonal; frozen coefficients at u for the nonlinear operator)
template <typename T, bool NONLINEAR>
__global__ void __launch_bounds__(256) column_wfact_kernel(T* __restrict__ dl, T* __restrict__ d, T* __restrict__ du,
                                                           const T* __restrict__ u, const T* __restrict__ acol,
                                                           double dtg, int nlev, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t col = i / (size_t)nlev;
  int lev = (int)(i - col * (size_t)nlev);
  T a = acol[col];
  if (!NONLINEAR) {
    double o = __dmul_rn(dtg, (double)a);
    dl[i] = (T)o;
    du[i] = (T)o;
    d[i] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
  } else {
    T uc = u[i];
    T um = lev > 0 ? u[i - 1] : T(0);
    T up = lev < nlev - 1 ? u[i + 1] : T(0);
    double lo = __dmul_rn(dtg, (double)kface(a, um, uc));
    double hi = __dmul_rn(dtg, (double)kface(a, uc, up));
    dl[i] = (T)lo;
    du[i] = (T)hi;
    d[i] = (T)__dsub_rn(-__dadd_rn(lo, hi), 1.0);
  }
}

// T_exp = S[lev]*g(t) + nu * (((uw + ue) + us) + un - 4 uc), owned rows only.  One thread per VN consecutive
// levels (VN = 16 bytes when the level count allows it): the four horizontal neighbours of a level vector are
// whole columns away, so every access is a full-width coalesced 16-byte load.
template <typename T, int VN>
__global__ void __launch_bounds__(256) column_texp_kernel(T* __restrict__ out, const T* __restrict__ u,
                                                          const T* __restrict__ S, T g, T nu, int nlev, size_t nx,
                                                          size_t ny_local, int halo, int has_src) {
  const size_t row_elems = nx * (size_t)nlev;
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VN;  // first element, within the owned rows
  if (i >= ny_local * row_elems) return;
  const size_t jr = i / row_elems;      // owned row
  const size_t r = i - jr * row_elems;  // offset in the row
  const size_t ix = r / (size_t)nlev;
  const int lev = (int)(r - ix * (size_t)nlev);
  const size_t rows = ny_local + 2 * (size_t)halo;
  const size_t j = jr + (size_t)halo;   // row in the local array
  const size_t idx = j * row_elems + r;
  T val[VN];
  auto ld = [&](const T* p, T (&e)[VN]) {
    if (VN == 1) {
      e[0] = p[0];
    } else {
      typename Vec<T>::type v = *reinterpret_cast<const typename Vec<T>::type*>(p);
      T t[Vec<T>::N];
      vec_unpack<T>(v, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) e[k] = t[k];
    }
  };
  if (has_src) {
    T s[VN];
    ld(S + lev, s);
#pragma unroll
    for (int k = 0; k < VN; ++k) val[k] = rmul(s[k], g);
  } else {
#pragma unroll
    for (int k = 0; k < VN; ++k) val[k] = T(0);
  }
  if (nu != T(0)) {
    const size_t ixw = ix == 0 ? nx - 1 : ix - 1, ixe = ix == nx - 1 ? 0 : ix + 1;
    size_t js, jn;
    if (halo) {
      js = j - 1;
      jn = j + 1;
    } else {  // single-rank periodic wrap without ghost rows
      js = j == 0 ? rows - 1 : j - 1;
      jn = j == rows - 1 ? 0 : j + 1;
    }
    T uc[VN], uw[VN], ue[VN], us[VN], un[VN];
    ld(u + idx, uc);
    ld(u + j * row_elems + ixw * (size_t)nlev + lev, uw);
    ld(u + j * row_elems + ixe * (size_t)nlev + lev, ue);
    ld(u + js * row_elems + r, us);
    ld(u + jn * row_elems + r, un);
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const T lap = rsub(radd(radd(radd(uw[k], ue[k]), us[k]), un[k]), rmul(T(4), uc[k]));
      val[k] = radd(val[k], rmul(nu, lap));
    }
  }
  if (VN == 1) {
    out[idx] = val[0];
  } else {
    T t[Vec<T>::N];
#pragma unroll
    for (int k = 0; k < VN; ++k) t[k] = val[k];
    *reinterpret_cast<typename Vec<T>::type*>(out + idx) = vec_pack<T>(t);
  }
}

// Marching variant (used whenever the level count is a multiple of the 16-byte vector): a thread owns one 16-byte level vector of a row chunk and walks ROWS rows of the slow horizontal
// axis with a sliding (south, centre, north) register window, so each element is fetched once per strip instead of
// three times; west/east come from lines the CTA touched one iteration earlier (L1).  Same arithmetic.
template <typename T, int VN>
__global__ void __launch_bounds__(256) column_texp_march_kernel(T* __restrict__ out, const T* __restrict__ u,
                                                                const T* __restrict__ S, T g, T nu, unsigned nlev,
                                                                unsigned nx, unsigned row_elems, size_t ny_local,
                                                                int halo, int has_src, size_t jr_begin, size_t jr_end,
                                                                int ROWS) {
  using V = typename Vec<T>::type;
  const unsigned r = (blockIdx.x * 256u + threadIdx.x) * VN;
  if (r >= row_elems) return;
  const size_t rows = ny_local + 2 * (size_t)halo;
  const unsigned ix = r / nlev, lev = r - ix * nlev;
  const unsigned wrap = (nx - 1) * nlev;
  const unsigned rw = ix == 0 ? r + wrap : r - nlev;
  const unsigned re = ix == nx - 1 ? r - wrap : r + nlev;
  T src[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) src[e] = T(0);
  if (has_src) {
    T t[VN];
    vec_unpack<T>(*reinterpret_cast<const V*>(S + lev), t);
#pragma unroll
    for (int e = 0; e < VN; ++e) src[e] = rmul(t[e], g);
  }
  const size_t jr0 = jr_begin + (size_t)blockIdx.y * ROWS;  // owned rows [jr_begin, jr_end) of this launch
  const size_t jr1 = jr0 + ROWS < jr_end ? jr0 + ROWS : jr_end;
  // row of the local array that holds owned row jr (jr = -1 and jr = ny_local are the neighbours across the slab
  // edge: ghost rows with a halo, the periodic wrap without)
  auto row_of = [&](long long jr) -> const T* {
    size_t j;
    if (halo)
      j = (size_t)(jr + 1);
    else
      j = jr < 0 ? rows - 1 : ((size_t)jr == rows ? 0 : (size_t)jr);
    return u + j * (size_t)row_elems;
  };
  V vs = *reinterpret_cast<const V*>(row_of((long long)jr0 - 1) + r);
  V vc = *reinterpret_cast<const V*>(row_of((long long)jr0) + r);
  auto emit = [&](size_t jr, const V& vw, const V& ve, const V& vn) {
    T c[VN], w[VN], e_[VN], s_[VN], n_[VN], val[VN];
    vec_unpack<T>(vc, c);
    vec_unpack<T>(vw, w);
    vec_unpack<T>(ve, e_);
    vec_unpack<T>(vs, s_);
    vec_unpack<T>(vn, n_);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const T lap = rsub(radd(radd(radd(w[e], e_[e]), s_[e]), n_[e]), rmul(T(4), c[e]));
      val[e] = radd(src[e], rmul(nu, lap));
    }
    *reinterpret_cast<V*>(out + (jr + (size_t)halo) * (size_t)row_elems + r) = vec_pack<T>(val);
    vs = vc;
    vc = vn;
  };
  constexpr int B = 4;  // rows per batch: 3*B independent 16-byte loads in flight per thread
  size_t jr = jr0;
  for (; jr + B <= jr1; jr += B) {
    V vn[B], vw[B], ve[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const T* crow = row_of((long long)(jr + k));
      vn[k] = *reinterpret_cast<const V*>(row_of((long long)(jr + k) + 1) + r);
      vw[k] = *reinterpret_cast<const V*>(crow + rw);
      ve[k] = *reinterpret_cast<const V*>(crow + re);
    }
#pragma unroll
    for (int k = 0; k < B; ++k) emit(jr + k, vw[k], ve[k], vn[k]);
  }
  for (; jr < jr1; ++jr) {
    const T* crow = row_of((long long)jr);
    const V vn = *reinterpret_cast<const V*>(row_of((long long)jr + 1) + r);
    const V vw = *reinterpret_cast<const V*>(crow + rw);
    const V ve = *reinterpret_cast<const V*>(crow + re);
    emit(jr, vw, ve, vn);
  }
}

// Rows a CTA of the marching kernels walks through.  8 is the measured optimum on a full 1024-row slab (more rows per strip
// = fewer re-fetched rows, fewer CTAs in flight).  A rank's share of a sharded problem has few strips: 112 interior rows in
// strips of 8 are 3584 CTAs = 4.04 waves of 888, i.e. a fifth round for 4 % of the work -- so on small launches the strip
// height is the one of 4..16 that minimises (rounds of CTA waves) x (rows + the two extra rows a strip fetches).
template <typename K>
int march_strip_rows(K kernel, cts_ctx* ctx, size_t nrows, size_t xblocks) {
  static std::mutex mu;  // (K is a function-pointer TYPE shared by several instantiations: key the cache by the pointer)
  static std::unordered_map<const void*, int> cache;
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find((const void*)kernel);
    if (it == cache.end()) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
      }
      cache.emplace((const void*)kernel, per_sm);
    } else {
      per_sm = it->second;
    }
  }
  const double wave = (double)per_sm * ctx->num_sms;
  auto rounds = [&](int r) { return std::ceil((double)((nrows + r - 1) / r) * xblocks / wave); };
  if ((double)((nrows + 7) / 8) * xblocks / wave >= 16.0) return 8;
  int best = 8;
  double best_cost = rounds(8) * (8 + 0.6);
  for (int r = 4; r <= 16; ++r) {
    const double c = rounds(r) * (r + 0.6);
    if (c < best_cost - 1e-9) {
      best_cost = c;
      best = r;
    }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// Final update of an IMEX-ARK step fused with the LAST stage's explicit tendency (imex_ark.jl:79-101 + :256-258): the last
// T_exp! output is only ever read by the final update (a_exp[:, s] == 0), so instead of writing it and reading it back
//   T_exp[s] = S*g(t) + nu*lap(U_s)                                   (column_texp_march_kernel)
//   u        = (u + (Σ_{j<s} c_j T_exp[j] + c_s T_exp[s])) + Σ c'_j T_imp[j]      (axpy_n, fused_increment.jl:220-236)
// run as one marching pass: 2w B/DOF less.  The tendency is rounded to T exactly as the stored array would be and the
// left folds keep their order, so u is bit-identical.  Owned rows only (the halo dss! that follows rewrites the ghosts).
template <int NT>
struct FinalArgs {
  AxpyArgs<NT> ax;     // base = out = u; terms = the OTHER tendencies (group 1: explicit j < s, then group 2: implicit)
  const void* U;       // last stage value (stencil source)
  const void* S;
  double coef_last;    // dt*b_exp[s]
  double g, nu;
  unsigned nlev, nx, row_elems;
  size_t ny_local;
  int halo, has_src;
  size_t jr_begin, jr_end;  // owned rows [jr_begin, jr_end) of this launch
  int strip_rows;           // rows a CTA marches through (march_strip_rows)
};

template <typename T, typename C, int NT>
__global__ void __launch_bounds__(256) column_final_update_kernel(const FinalArgs<NT> a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  const unsigned r = (blockIdx.x * 256u + threadIdx.x) * VN;
  if (r >= a.row_elems) return;
  const T* __restrict__ U = reinterpret_cast<const T*>(a.U);
  const size_t rows = a.ny_local + 2 * (size_t)a.halo;
  const unsigned ix = r / a.nlev, lev = r - ix * a.nlev;
  const unsigned wrap = (a.nx - 1) * a.nlev;
  const unsigned rw = ix == 0 ? r + wrap : r - a.nlev;
  const unsigned re = ix == a.nx - 1 ? r - wrap : r + a.nlev;
  const T g = (T)a.g, nu = (T)a.nu;
  T src[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) src[e] = T(0);
  if (a.has_src) {
    T t[VN];
    vec_unpack<T>(*reinterpret_cast<const V*>(reinterpret_cast<const T*>(a.S) + lev), t);
#pragma unroll
    for (int e = 0; e < VN; ++e) src[e] = rmul(t[e], g);
  }
  const size_t jr0 = a.jr_begin + (size_t)blockIdx.y * a.strip_rows;
  const size_t jr1 = jr0 + a.strip_rows < a.jr_end ? jr0 + a.strip_rows : a.jr_end;
  auto local_row = [&](long long jr) -> size_t {
    if (a.halo) return (size_t)(jr + 1);
    return jr < 0 ? rows - 1 : ((size_t)jr == rows ? 0 : (size_t)jr);
  };
  V vs = *reinterpret_cast<const V*>(U + local_row((long long)jr0 - 1) * (size_t)a.row_elems + r);
  V vc = *reinterpret_cast<const V*>(U + local_row((long long)jr0) * (size_t)a.row_elems + r);
  constexpr int B = NT <= 2 ? 4 : 2;  // rows per batch: (3 + 1 + NT)*B independent 16-byte loads in flight per thread
  auto row_update = [&](size_t jr, const V& vw, const V& ve, const V& vn, const V& vb, const V (&vt)[NTS]) {
    T c[VN], w[VN], e_[VN], s_[VN], n_[VN], eb[VN], eo[VN], et[NTS][VN];
    vec_unpack<T>(vc, c);
    vec_unpack<T>(vw, w);
    vec_unpack<T>(ve, e_);
    vec_unpack<T>(vs, s_);
    vec_unpack<T>(vn, n_);
    vec_unpack<T>(vb, eb);
#pragma unroll
    for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t], et[t]);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const T lap = rsub(radd(radd(radd(w[e], e_[e]), s_[e]), n_[e]), rmul(T(4), c[e]));
      const T te = radd(src[e], rmul(nu, lap));  // T_exp[s], rounded to T as the stored array would be
      C rr = (C)eb[e];
      C g1 = 0, g2 = 0;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const C p = op_mul((C)a.ax.coef[k], (C)et[k][e]);
        if (k < a.ax.n1)
          g1 = (k == 0) ? p : op_add(g1, p);
        else
          g2 = (k == a.ax.n1) ? p : op_add(g2, p);
      }
      const C pl = op_mul((C)a.coef_last, (C)te);
      g1 = a.ax.n1 > 0 ? op_add(g1, pl) : pl;
      rr = op_add(rr, g1);
      if (a.ax.n1 < NT) rr = op_add(rr, g2);
      eo[e] = (T)rr;
    }
    const size_t gi = local_row((long long)jr) * (size_t)a.row_elems + r;
    *reinterpret_cast<V*>(reinterpret_cast<T*>(a.ax.out) + gi) = vec_pack<T>(eo);
    vs = vc;
    vc = vn;
  };
  size_t jr = jr0;
  for (; jr + B <= jr1; jr += B) {
    V vn[B], vw[B], ve[B], vb[B], vt[B][NTS];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const size_t rowc = local_row((long long)(jr + k)) * (size_t)a.row_elems;
      vn[k] = *reinterpret_cast<const V*>(U + local_row((long long)(jr + k) + 1) * (size_t)a.row_elems + r);
      vw[k] = *reinterpret_cast<const V*>(U + rowc + rw);
      ve[k] = *reinterpret_cast<const V*>(U + rowc + re);
      vb[k] = *reinterpret_cast<const V*>(reinterpret_cast<const T*>(a.ax.base) + rowc + r);
#pragma unroll
      for (int t = 0; t < NT; ++t) vt[k][t] = *reinterpret_cast<const V*>(reinterpret_cast<const T*>(a.ax.terms[t]) + rowc + r);
    }
#pragma unroll
    for (int k = 0; k < B; ++k) row_update(jr + k, vw[k], ve[k], vn[k], vb[k], vt[k]);
  }
  for (; jr < jr1; ++jr) {
    const size_t rowc = local_row((long long)jr) * (size_t)a.row_elems;
    V vt[NTS];
    const V vn = *reinterpret_cast<const V*>(U + local_row((long long)jr + 1) * (size_t)a.row_elems + r);
    const V vw = *reinterpret_cast<const V*>(U + rowc + rw);
    const V ve = *reinterpret_cast<const V*>(U + rowc + re);
    const V vb = *reinterpret_cast<const V*>(reinterpret_cast<const T*>(a.ax.base) + rowc + r);
#pragma unroll
    for (int t = 0; t < NT; ++t) vt[t] = *reinterpret_cast<const V*>(reinterpret_cast<const T*>(a.ax.terms[t]) + rowc + r);
    row_update(jr, vw, ve, vn, vb, vt);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused implicit stage for the LINEAR column operator, Newton max_iters = 1, direct solve:
//   U0   = (c0*base + G1) + G2                        K1  (fused_increment.jl:220-236)
//   r    = T_imp(U0)                                  user hook (linear diffusion)
//   f    = (U0 + dtγ r) - U0                          K3  (imex_ark.jl:297, evaluated in Float64)
//   dx   = W \ f                                      K4  (two-sided Thomas; W stored or regenerated)
//   U    = U0 - dx                                    K5  (newtons_method.jl:651)
//   Timp = (U - U0)/dtγ                               K6  (imex_ark.jl:269)
// One CTA per 32-column tile; U0 is assembled straight from global memory with 128-bit loads into a
// padded shared tile, the elimination runs out of shared memory, U / T_imp leave with 128-bit stores.
// ---------------------------------------------------------------------------------------------

#ifndef CTS_FUSED_THREADS
#define CTS_FUSED_THREADS 128
#endif
constexpr int kFusedThreads = CTS_FUSED_THREADS;  // the solver warps (a lane pair per column) + warps that only move data

template <typename T, typename C, int NT, bool MATRIX_FREE>
#ifndef CTS_FUSED_MIN_BLOCKS
#define CTS_FUSED_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(kFusedThreads, CTS_FUSED_MIN_BLOCKS) column_fused_stage_kernel(const FusedArgs<NT> a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  extern __shared__ __align__(128) char smem[];
  const int pitch = a.pitch;
  const int tile_bytes = kTileCols * pitch;
  char* sU0 = smem;
  char* sr = smem + tile_bytes;
  char* sfac = smem + 2 * tile_bytes;  // matrix-free: cp/lp scratch; stored: dl tile (d and du follow)
  char* sdl = sfac;
  char* sd = smem + 3 * tile_bytes;
  char* sdu = smem + 4 * tile_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (MATRIX_FREE ? 3 : 5) * tile_bytes);
  T* sAcol = reinterpret_cast<T*>(smem + (MATRIX_FREE ? 3 : 5) * tile_bytes + 16);  // K/dz^2 of the tile's columns

  const int nlev = a.nlev;
  const size_t col0 = (size_t)blockIdx.x * kTileCols;
  const int nvalid = (int)((a.ncols - col0) < (size_t)kTileCols ? (a.ncols - col0) : (size_t)kTileCols);
  const int col_bytes = nlev * (int)sizeof(T);
  const size_t goff = col0 * (size_t)nlev;
  long long tstamp[5];
  tstamp[0] = clock64();
  if ((int)threadIdx.x < nvalid) sAcol[threadIdx.x] = ((const T*)a.Kcol)[col0 + threadIdx.x];  // visible after phase A's barrier

  if (!MATRIX_FREE) {  // W tiles arrive by bulk TMA while U0 is being assembled
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(bar, (uint32_t)(3 * nvalid * col_bytes));
    }
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
      const int c = threadIdx.x;
      const size_t g = goff + (size_t)c * nlev;
      bulk_g2s(sdl + c * pitch, (const T*)a.dl + g, col_bytes, bar);
      bulk_g2s(sd + c * pitch, (const T*)a.d + g, col_bytes, bar);
      bulk_g2s(sdu + c * pitch, (const T*)a.du + g, col_bytes, bar);
    }
  }

  // ---- phase A: U0 tile = K1(base, terms).  Two 16-byte vectors per thread and round, every operand's load
  // issued before the first use: (1+NT)*2 independent requests per thread, 256 threads per CTA. ---------------
  const int cpc = col_bytes / 16;
  const int total = nvalid * cpc;
  constexpr int UN = 2;
  for (int q0 = threadIdx.x; q0 < total; q0 += UN * kFusedThreads) {
    V vb[UN], vt[NTS][UN];
    size_t gi[UN];
    int so[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int q = q0 + u * kFusedThreads;
      ok[u] = q < total;
      const int col = q / cpc, k = q - col * cpc;
      so[u] = col * pitch + k * 16;
      gi[u] = (goff * sizeof(T) + (size_t)col * col_bytes + (size_t)k * 16) / 16;  // 16-byte vector index
      if (ok[u]) {
        vb[u] = reinterpret_cast<const V*>(a.ax.base)[gi[u]];
#pragma unroll
        for (int t = 0; t < NT; ++t) vt[t][u] = reinterpret_cast<const V*>(a.ax.terms[t])[gi[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!ok[u]) continue;
      T eb[VN], eo[VN], et[NTS][VN];
      vec_unpack<T>(vb[u], eb);
#pragma unroll
      for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t][u], et[t]);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        T x[NTS];
#pragma unroll
        for (int t = 0; t < NT; ++t) x[t] = et[t][e];
        eo[e] = axpy_combine<T, C, NT>(eb[e], x, a.ax);
      }
      V vo = vec_pack<T>(eo);
      *reinterpret_cast<V*>(sU0 + so[u]) = vo;
      if (a.temp) reinterpret_cast<V*>(a.temp)[gi[u]] = vo;
    }
  }
  __syncthreads();
  tstamp[1] = clock64();

  // ---- phase A2: Newton residual f = (U0 + dtγ T_imp(U0)) - U0 (K3), element-parallel over the tile: every thread
  // takes 16-byte level vectors of U0 plus the two neighbouring levels and writes the rhs tile. ------------------
  for (int q = threadIdx.x; q < total; q += kFusedThreads) {
    const int c = q / cpc, k = q - c * cpc;
    const T acol = sAcol[c];
    const T* U0c = reinterpret_cast<const T*>(sU0 + c * pitch);
    const int l0 = k * VN;
    T uu[VN], rr[VN];
    vec_unpack<T>(*reinterpret_cast<const V*>(sU0 + c * pitch + k * 16), uu);
    T um = l0 > 0 ? U0c[l0 - 1] : T(0);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const int l = l0 + e;
      const T up = (e + 1 < VN) ? uu[e + 1 < VN ? e + 1 : e] : (l < nlev - 1 ? U0c[l + 1] : T(0));
      const T ti = timp_linear<T>(acol, um, uu[e], up);
      rr[e] = (T)__dsub_rn(__dadd_rn((double)uu[e], __dmul_rn(a.dtg, (double)ti)), (double)uu[e]);
      um = uu[e];
    }
    *reinterpret_cast<V*>(sr + c * pitch + k * 16) = vec_pack<T>(rr);
  }
  __syncthreads();
  tstamp[2] = clock64();

  // ---- phase B (solver warps): two-sided product-form Thomas, a lane pair per column ---------------------
  if (threadIdx.x < kTileThreads) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = warp * 16 + lane_column(lane);
    const bool is_up = lane_is_up(lane);
    const bool valid = col < nvalid;
    const int col_off = col * pitch;
    if (MATRIX_FREE) {
      const T acol = valid ? sAcol[col] : T(0);
      const double o = __dmul_rn(a.dtg, (double)acol);
      ConstRows<T> rows{(T)o, (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0), sfac};
      babe_solve_lanepair<T>(rows, RhsFromTile<T>{}, sr, col_off, nlev, is_up, valid);
    } else {
      mbar_wait(bar, 0);
      StoredRows<T> rows{sdl, sd, sdu};
      babe_solve_lanepair<T>(rows, RhsFromTile<T>{}, sr, col_off, nlev, is_up, valid);
    }
  }
  __syncthreads();
  tstamp[3] = clock64();

  // ---- phase C: U = U0 - dx ; Timp = (U - U0)/dtγ ; coalesced 128-bit stores -----------------------
  const DivConst dc = div_const_prepare(a.dtg);
  for (int q = threadIdx.x; q < total; q += kFusedThreads) {
    const int c = q / cpc, k = q - c * cpc;
    const size_t gi = (goff * sizeof(T) + (size_t)c * col_bytes + (size_t)k * 16) / 16;
    T e0[VN], ex[VN], eU[VN], eT[VN];
    vec_unpack<T>(*reinterpret_cast<const V*>(sU0 + c * pitch + k * 16), e0);
    vec_unpack<T>(*reinterpret_cast<const V*>(sr + c * pitch + k * 16), ex);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      eU[e] = rsub(e0[e], ex[e]);
      eT[e] = (T)div_const((double)rsub(eU[e], e0[e]), dc);
    }
    reinterpret_cast<V*>(a.U)[gi] = vec_pack<T>(eU);
    if (a.Timp) reinterpret_cast<V*>(a.Timp)[gi] = vec_pack<T>(eT);
    if (MATRIX_FREE && a.emit_w) {  // same expressions as column_wfact_kernel
      const double o = __dmul_rn(a.dtg, (double)sAcol[c]);
      T eo[VN], ed[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        eo[e] = (T)o;
        ed[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
      }
      const V vo = vec_pack<T>(eo);
      reinterpret_cast<V*>(const_cast<void*>(a.dl))[gi] = vo;
      reinterpret_cast<V*>(const_cast<void*>(a.d))[gi] = vec_pack<T>(ed);
      reinterpret_cast<V*>(const_cast<void*>(a.du))[gi] = vo;
    }
  }
  if (a.dbg && threadIdx.x == 0) {
    tstamp[4] = clock64();
#pragma unroll
    for (int k = 0; k < 5; ++k) a.dbg[(size_t)blockIdx.x * 5 + k] = tstamp[k];
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant of the fused stage (the one the engine normally runs).  When a column is a power-of-two
// number (<= 32) of 16-byte vectors, consecutive lanes hold consecutive level vectors of one column, so the K3
// residual's level neighbours come from warp shuffles and U0 never has to live in shared memory: every thread keeps its
// (at most kKeep) U0 vectors in registers across the solve and reuses them in the epilogue.  Shared memory per tile
// drops from 5 to 4 column tiles (stored W) or from 3 to 2 (matrix free), and the separate residual phase with its
// barrier disappears.  Arithmetic is identical to the kernel above.  (Measured on B200: capping registers at 80 to fit
// 6 CTAs per SM costs the solver more than the extra CTA gives: 6.1 ms vs 5.4 ms per step; 64/96-thread CTAs have too
// few loads in flight: 6.7 ms.  So: 128 threads, <= 96 registers, 5 CTAs per SM.)
constexpr int kKeep = (CTS_FUSED_TILE_COLS * 32 + kFusedThreads - 1) / kFusedThreads;  // U0 vectors per thread (f64: 16 cols x <= 32 vectors, f32: 32 x <= 16)
// ... per element type: a Float32 tile of CTS_FUSED_TILE_COLS_F32 columns holds <= 16 vectors per column at 64 levels
template <typename T>
struct KeepOf {
  static constexpr int value = sizeof(T) == 4 ? (TileShape<float>::kCols * 16 + kFusedThreads - 1) / kFusedThreads : kKeep;
};

#ifndef CTS_FUSED_PARK
#define CTS_FUSED_PARK 1
#endif
#ifndef CTS_FUSED_PINGPONG
#define CTS_FUSED_PINGPONG false
#endif
#ifndef CTS_FUSED_REG_MIN_BLOCKS
#define CTS_FUSED_REG_MIN_BLOCKS 5
#endif
// N1SPEC: how the kernel learns the split of the NT operands into the explicit group (first n1) and the implicit group:
// 0 = from the arguments at run time, 1 = all in group 1 (n1 == NT), 2 = n1 == ceil(NT/2) (stage i of an IMEX-ARK tableau
// with dense lower triangles: i-1 explicit and i-2 implicit tendencies).  A compile-time split needs no selects.

// Resident CTAs per SM the register allocation aims at.  The stage is bound by the number of tiles in flight (a tile lives
// ~19 000 cycles: load -> solve -> store), so a sixth CTA is worth more than registers: stored-W launches with at most
// CTS_FUSED_SIX_BLOCKS_NT_MAX operands (Float64; all Float32 ones) and all matrix-free launches are held to 80 registers --
// spill-free since U0 is parked in shared memory during the solve -- and shared memory allows six stored-W tiles of 64
// levels (6 x (35 984 + 1024) B <= 228 KB).  Measured (ncu, 2^21 columns): 3 operands 1.77 -> 1.51 ms per launch, 1 operand
// matrix-free 1.20 -> 1.14 ms; 5 operands 1.87 -> 1.91 ms (the 12 loads of a round need the 96 registers), so it stays at 5.
#ifndef CTS_FUSED_SIX_BLOCKS_NT_MAX
#define CTS_FUSED_SIX_BLOCKS_NT_MAX 4
#endif
#ifndef CTS_FUSED_PARK_F32
#define CTS_FUSED_PARK_F32 2
#endif
#ifndef CTS_FUSED_F32_BLOCKS
#define CTS_FUSED_F32_BLOCKS 6
#endif
#ifndef CTS_FUSED_MF_BLOCKS
#define CTS_FUSED_MF_BLOCKS 6
#endif
#ifndef CTS_FUSED_MF_BLOCKS_NT_MAX
#define CTS_FUSED_MF_BLOCKS_NT_MAX 5
#endif
// U0 vectors a solver thread parks in shared memory during the solve (the rest stay in registers).  A stored-W tile of
// Float32 (32 columns) has twice the solver threads of a Float64 tile: parking all four vectors (4 KiB) would leave room for
// only five CTAs per SM, parking two of them (2 KiB) lets the sixth in.
template <typename T, bool MATRIX_FREE>
struct FusedPark {
  static constexpr int value = (sizeof(T) == 4 && !MATRIX_FREE && CTS_FUSED_PARK_F32 < KeepOf<T>::value) ? CTS_FUSED_PARK_F32 : KeepOf<T>::value;
};
template <typename T, int NT, bool MATRIX_FREE>
struct FusedMinBlocks {
  static constexpr int value = MATRIX_FREE ? (NT <= CTS_FUSED_MF_BLOCKS_NT_MAX ? CTS_FUSED_MF_BLOCKS : CTS_FUSED_REG_MIN_BLOCKS)
                                           : ((sizeof(T) == 4 && NT <= 5) ? CTS_FUSED_F32_BLOCKS : (NT <= CTS_FUSED_SIX_BLOCKS_NT_MAX ? 6 : CTS_FUSED_REG_MIN_BLOCKS));
};

template <typename T, typename C, int NT, bool MATRIX_FREE, int N1SPEC>
__global__ void __launch_bounds__(kFusedThreads, FusedMinBlocks<T, NT, MATRIX_FREE>::value) column_fused_stage_reg_kernel(const FusedArgs<NT> a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  constexpr int kCols = TileShape<T>::kCols;       // 16 columns of Float64, 32 of Float32
  constexpr int kSolvers = TileShape<T>::kThreads;  // solver threads: a lane pair per column
  constexpr int kPark = FusedPark<T, MATRIX_FREE>::value;
  constexpr int KEEP = KeepOf<T>::value;  // 16-byte vectors per thread and tile
  extern __shared__ __align__(128) char smem[];
  const int pitch = a.pitch;
  const int tile_bytes = kCols * pitch;
  char* sx = smem;                   // residual -> rp -> dx
  char* sfac = smem + tile_bytes;    // matrix-free: cp scratch; stored: dl tile (d and du follow)
  char* sdl = sfac;
  char* sd = smem + 2 * tile_bytes;
  char* sdu = smem + 3 * tile_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (MATRIX_FREE ? 2 : 4) * tile_bytes);
  T* sAcol = reinterpret_cast<T*>(smem + (MATRIX_FREE ? 2 : 4) * tile_bytes + 16);
  // the solver threads park their U0 vectors here during the solve, so that those registers are free for the sweeps
  V* sPark = reinterpret_cast<V*>(smem + (MATRIX_FREE ? 2 : 4) * tile_bytes + 16 + ((kCols * (int)sizeof(T) + 15) / 16) * 16);

  const int nlev = a.nlev;
  const size_t col0 = (size_t)blockIdx.x * kCols;
  const int nvalid = (int)((a.ncols - col0) < (size_t)kCols ? (a.ncols - col0) : (size_t)kCols);
  const int col_bytes = nlev * (int)sizeof(T);
  const size_t goff = col0 * (size_t)nlev;
  const size_t gvec0 = goff * sizeof(T) / 16;  // first 16-byte vector of the tile in every operand array
  if (MATRIX_FREE && (int)threadIdx.x < nvalid) sAcol[threadIdx.x] = ((const T*)a.Kcol)[col0 + threadIdx.x];

  if (!MATRIX_FREE) {  // W tiles arrive by bulk TMA while U0 is being assembled
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(bar, (uint32_t)(3 * nvalid * col_bytes));
    }
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
      const int c = threadIdx.x;
      const size_t g = goff + (size_t)c * nlev;
      bulk_g2s(sdl + c * pitch, (const T*)a.dl + g, col_bytes, bar);
      bulk_g2s(sd + c * pitch, (const T*)a.d + g, col_bytes, bar);
      bulk_g2s(sdu + c * pitch, (const T*)a.du + g, col_bytes, bar);
    }
  }

  // ---- phase A: U0 = K1(base, terms) in registers, residual f = (U0 + dtγ T_imp(U0)) - U0 (K3) through warp
  // shuffles, residual tile to shared memory.  cpc = vectors per column (power of two <= 32). ----------------
  const int cpc = col_bytes / 16;
  const int cshift = 31 - __clz(cpc);
  const int total = nvalid * cpc;
  V keep[KEEP];
  constexpr int UN = 2;  // (one round at a time so that the 5-operand stage fits 80 registers / six CTAs: within noise, 6.94-6.95 vs 6.98 ms/step)
  static_assert(KEEP % UN == 0, "rounds are processed in pairs");
  // N1c: the number of group-1 terms as a compile-time constant (-1: take it from the arguments at run time)
  auto phase_a = [&](auto N1c) {
  constexpr int N1 = decltype(N1c)::value;
#pragma unroll
  for (int r0 = 0; r0 < KEEP; r0 += UN) {
    V vb[UN], vt[NTS][UN];
    size_t gi[UN];
    int so[UN], kk[UN];
    bool ok[UN];
    T acol[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int q = (int)threadIdx.x + (r0 + u) * kFusedThreads;
      ok[u] = q < total;
      const int col = q >> cshift;
      kk[u] = q & (cpc - 1);
      so[u] = col * pitch + kk[u] * 16;
      gi[u] = gvec0 + (size_t)q;  // 16-byte vector index: the tile's columns are contiguous, so it is simply tile base + q
      acol[u] = T(0);
      if (ok[u]) {
        vb[u] = reinterpret_cast<const V*>(a.ax.base)[gi[u]];
#pragma unroll
        for (int t = 0; t < NT; ++t) vt[t][u] = reinterpret_cast<const V*>(a.ax.terms[t])[gi[u]];
        acol[u] = ((const T*)a.Kcol)[col0 + col];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      T eb[VN], uu[VN], rr[VN], et[NTS][VN];
      vec_unpack<T>(vb[u], eb);
#pragma unroll
      for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t][u], et[t]);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        T x[NTS];
#pragma unroll
        for (int t = 0; t < NT; ++t) x[t] = et[t][e];
        if constexpr (N1 >= 0)
          uu[e] = axpy_combine_n1<T, C, NT, (N1 >= 0 ? N1 : 0)>(eb[e], x, a.ax);
        else
          uu[e] = axpy_combine<T, C, NT>(eb[e], x, a.ax);
      }
      keep[r0 + u] = vec_pack<T>(uu);
      // level neighbours across the vector boundary: the previous / next lane of the same column segment
      T below = __shfl_up_sync(0xffffffffu, uu[VN - 1], 1, cpc);
      T above = __shfl_down_sync(0xffffffffu, uu[0], 1, cpc);
      if (kk[u] == 0) below = T(0);        // Dirichlet 0 below level 0
      if (kk[u] == cpc - 1) above = T(0);  // ... and above the top level
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const T um = e > 0 ? uu[e > 0 ? e - 1 : 0] : below;
        const T up = e + 1 < VN ? uu[e + 1 < VN ? e + 1 : e] : above;
        const T ti = timp_linear<T>(acol[u], um, uu[e], up);
        rr[e] = (T)__dsub_rn(__dadd_rn((double)uu[e], __dmul_rn(a.dtg, (double)ti)), (double)uu[e]);
      }
      if (ok[u]) {
        *reinterpret_cast<V*>(sx + so[u]) = vec_pack<T>(rr);
        if (a.temp) reinterpret_cast<V*>(a.temp)[gi[u]] = keep[r0 + u];
      }
    }
  }
  };
  phase_a(std::integral_constant<int, GroupSplit<NT, N1SPEC>::value>{});
  __syncthreads();

  // ---- phase B (solver warps): two-sided product-form Thomas, a lane pair per column ---------------------
  if (threadIdx.x < kSolvers) {
#if CTS_FUSED_PARK
#pragma unroll
    for (int r = 0; r < kPark; ++r) sPark[r * kSolvers + threadIdx.x] = keep[r];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = warp * 16 + lane_column(lane);
    const bool is_up = lane_is_up(lane);
    const bool valid = col < nvalid;
    const int col_off = col * pitch;
    if (MATRIX_FREE) {
      const T acol = valid ? sAcol[col] : T(0);
      const double o = __dmul_rn(a.dtg, (double)acol);
      ConstRows<T> rows{(T)o, (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0), sfac};
      babe_solve_lanepair<T, ConstRows<T>, RhsFromTile<T>, CTS_FUSED_PINGPONG>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
    } else {
      mbar_wait(bar, 0);
      StoredRows<T> rows{sdl, sd, sdu};
      babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, CTS_FUSED_PINGPONG>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
    }
    // every warp meets at named barrier 1 (the branches are warp-uniform); the solver threads then take their U0 back
    asm volatile("bar.sync 1, %0;" ::"n"(kFusedThreads) : "memory");
#if CTS_FUSED_PARK
#pragma unroll
    for (int r = 0; r < kPark; ++r) keep[r] = sPark[r * kSolvers + threadIdx.x];
#endif
  } else {
    if (MATRIX_FREE && a.emit_w) {
    // the warps that only move data write W = dtγJ - I (same expressions as column_wfact_kernel) while the solver
    // warp works: the stores overlap the elimination instead of lengthening the epilogue
    for (int q = (int)threadIdx.x - kSolvers; q < total; q += kFusedThreads - kSolvers) {
      const int c = q >> cshift, k = q & (cpc - 1);
      const size_t gi = gvec0 + (size_t)q;
      const double o = __dmul_rn(a.dtg, (double)sAcol[c]);
      T eo[VN], ed[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        eo[e] = (T)o;
        ed[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
      }
      const V vo = vec_pack<T>(eo);
      reinterpret_cast<V*>(const_cast<void*>(a.dl))[gi] = vo;
      reinterpret_cast<V*>(const_cast<void*>(a.d))[gi] = vec_pack<T>(ed);
      reinterpret_cast<V*>(const_cast<void*>(a.du))[gi] = vo;
    }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kFusedThreads) : "memory");
  }

  // ---- phase C: U = U0 - dx ; Timp = (U - U0)/dtγ ; coalesced 128-bit stores; U0 from registers ----------
  const DivConst dc = div_const_prepare(a.dtg);
#pragma unroll
  for (int r = 0; r < KEEP; ++r) {
    const int q = (int)threadIdx.x + r * kFusedThreads;
    if (q >= total) continue;
    const int c = q >> cshift, k = q & (cpc - 1);
    const size_t gi = gvec0 + (size_t)q;
    T e0[VN], ex[VN], eU[VN], eT[VN];
    vec_unpack<T>(keep[r], e0);
    vec_unpack<T>(*reinterpret_cast<const V*>(sx + c * pitch + k * 16), ex);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      eU[e] = rsub(e0[e], ex[e]);
      eT[e] = (T)div_const((double)rsub(eU[e], e0[e]), dc);
    }
    reinterpret_cast<V*>(a.U)[gi] = vec_pack<T>(eU);
    if (a.Timp) reinterpret_cast<V*>(a.Timp)[gi] = vec_pack<T>(eT);
  }
}

// ---------------------------------------------------------------------------------------------
// Pipelined, warp-specialised, persistent variant of the fused stage (one CTA per SM loops over its tiles).
//
// The kernels above give every 16-column tile to its own short-lived CTA whose load -> solve -> store phases are serial;
// five such CTAs per SM are the only overlap, three of four warps idle during a solve, and the loads in flight come in
// bursts (ncu: 30 % issue utilisation, DRAM at 67-80 %).  Here the phases of DIFFERENT tiles run concurrently in
// dedicated warps of one resident CTA, coupled by mbarriers over shared-memory rings:
//
//   producer  (1 warp)   bulk-TMA (cp.async.bulk + mbarrier complete_tx) of ALL operands of tile t+1, t+2: base and the NT
//                        tendencies as one contiguous 8 KiB copy each into a 2-deep staging ring, dl/d/du column by column
//                        into the padded tiles of a slot ring.  No registers, whole tiles in flight (~150 KiB per SM).
//   assemblers (4 warps) phase A of tile t: staging -> U0 = K1(base, terms) (LDS.128, conflict free) -> Newton residual via
//                        warp shuffles -> slot {U0, r};  frees the staging buffer.
//   solvers   (4 warps)  phase B: two-sided product-form Thomas on 16 columns per warp, in place in the slot.
//   epilogue  (4 warps)  phase C of tile t-1..t-2: U = U0 - dx, T_imp = (U - U0)/dtγ (and W when the stage folds Wfact in),
//                        128-bit coalesced stores;  frees the slot.
//
// Arithmetic, operation order and rounding are those of the kernels above (same device functions), so results are
// bit-identical; only the schedule differs.  A tile is always 512 16-byte vectors (16 columns of 64 Float64 levels,
// 32 columns of 64 Float32 levels, ...), so one 128-thread group covers it with 4 vectors per thread and a column is a
// power-of-two lane group of one warp.
// ---------------------------------------------------------------------------------------------
constexpr int kPipeAsm = 128;            // assembler threads (phase A)
constexpr int kPipeMaxSolverWarps = 6;   // solver warps (phase B): teams of `jobs` warps, one team per tile in flight
constexpr int kPipeStore = 128;          // epilogue threads (phase C)
constexpr int kPipeThreads = 64 + kPipeAsm + 32 * kPipeMaxSolverWarps + kPipeStore;  // 512
constexpr int kPipeTileVecs = 512;
constexpr int kPipeMaxDepth = 8;         // upper bound of every ring depth
constexpr int kPipeBarBytes = 512;       // 7 barrier arrays x 8 x 8 B = 448 B
constexpr int kPipeOpBytes = kPipeTileVecs * 16;  // one operand tile (8 KiB) of the operand ring

struct PipeCfg {
  int n_op;     // depth of the operand ring   (base + NT tendencies per stage, flat 8 KiB tiles, bulk TMA)
  int n_w;      // depth of the W ring         (stored W: dl|d|du padded tiles by bulk TMA; matrix-free: one factor tile per team)
  int n_x;      // depth of the x ring         (U0 flat + residual/x padded)
  int n_teams;  // solver teams (a team = `jobs` warps = one tile); stored W: == n_w
};

// mbarrier wait of the pipelined kernel with a watchdog: a protocol error must surface as a launch failure (trap), never
// as a hung GPU.  ~2 s at 2 GHz; the check only runs while the barrier is NOT ready.
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t ok;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {
        printf("cts: fused-stage pipeline watchdog: block %d thread %d barrier tag %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, tag, parity);
        __trap();
      }
    }
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Rings and their barriers (tile i uses buffer i % depth in phase i / depth of that buffer's barriers):
//   operand ring  op_full  (TMA complete_tx)        -> assemblers;   op_empty (128 assembler arrivals)  -> operand producer
//   W ring        w_full   (TMA complete_tx)        -> solver team;  w_empty  (32*jobs solver arrivals) -> W producer
//   x ring        x_full   (128 assembler arrivals) -> solver team;  x_solved (32*jobs)                 -> epilogue;
//                 x_empty  (128 epilogue arrivals)  -> assemblers
// A parity wait is only meaningful for the current or the immediately preceding phase of a barrier, so every waiter must
// meet the phases of a buffer in sequence: producers, assemblers and the epilogue walk ALL tiles in order; solver team k
// takes tiles k, k + n_teams, ... with n_teams <= n_x (x_full completes in tile order) and n_teams == n_w for stored W
// (a team always meets the same W buffer, whatever order the bulk copies complete in).
template <typename T, typename C, int NT, bool MATRIX_FREE, int N1SPEC>
__global__ void __launch_bounds__(kPipeThreads, 1) column_fused_stage_pipe_kernel(const FusedArgs<NT> a, const PipeCfg cfg) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  constexpr int kRounds = kPipeTileVecs / kPipeAsm;  // vectors per thread and tile
  extern __shared__ __align__(128) char smem[];
  const int nlev = a.nlev;
  const int pitch = a.pitch;
  const int col_bytes = nlev * (int)sizeof(T);
  const int cpc = col_bytes / 16;               // vectors per column: a power of two <= 32
  const int cshift = 31 - __clz(cpc);
  const int kCols = kPipeTileVecs / cpc;        // columns per tile
  const int jobs = kCols / 16;                  // solver warps per tile (16 columns each)
  const int tile_bytes = kCols * pitch;         // a padded tile (conflict-free column pitch)
  const int op_bytes = (1 + NT) * kPipeOpBytes;
  const int w_bytes = (MATRIX_FREE ? 1 : 3) * tile_bytes;
  const int x_bytes = kPipeOpBytes + tile_bytes;  // U0 (flat) | residual -> x (padded)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* op_full = bars;
  uint64_t* op_empty = bars + kPipeMaxDepth;
  uint64_t* w_full = bars + 2 * kPipeMaxDepth;
  uint64_t* w_empty = bars + 3 * kPipeMaxDepth;
  uint64_t* x_full = bars + 4 * kPipeMaxDepth;
  uint64_t* x_solved = bars + 5 * kPipeMaxDepth;
  uint64_t* x_empty = bars + 6 * kPipeMaxDepth;
  char* op_ring = smem + kPipeBarBytes;
  char* w_ring = op_ring + (size_t)cfg.n_op * op_bytes;
  char* x_ring = w_ring + (size_t)cfg.n_w * w_bytes;

  const size_t ntiles = (a.ncols + kCols - 1) / kCols;
  const int count = ntiles > blockIdx.x ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int b = 0; b < cfg.n_op; ++b) {
      mbar_init(op_full + b, 1);
      mbar_init(op_empty + b, kPipeAsm);
    }
    for (int b = 0; b < cfg.n_w; ++b) {
      mbar_init(w_full + b, 1);
      mbar_init(w_empty + b, 32 * jobs);
    }
    for (int s = 0; s < cfg.n_x; ++s) {
      mbar_init(x_full + s, kPipeAsm);
      mbar_init(x_solved + s, 32 * jobs);
      mbar_init(x_empty + s, kPipeStore);
    }
    fence_mbar_init();
  }
  __syncthreads();

  auto tile_of = [&](int i, size_t& col0, int& nvalid) {
    const size_t tile = blockIdx.x + (size_t)i * gridDim.x;
    col0 = tile * kCols;
    nvalid = (int)((a.ncols - col0) < (size_t)kCols ? (a.ncols - col0) : (size_t)kCols);
  };
  // optional wait-time profile (CTS_FUSED_TIMELINE): lane 0 of each role's first warp sums the cycles it spends in each wait
  long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = a.dbg ? clock64() : 0;
  auto wait = [&](uint64_t* bar, uint32_t parity, int tag) {
    if (a.dbg) {
      const long long t0 = clock64();
      mbar_wait_wd(bar, parity, tag);
      acc[tag] += clock64() - t0;
    } else {
      mbar_wait_wd(bar, parity, tag);
    }
  };
  int role = -1;  // profile row of this thread (only one thread per role reports)

  if (warp == 0) {
    // ------------------------------------------------------------------ operand producer -------------------------
    if (lane == 0) {
      role = 0;
      for (int i = 0; i < count; ++i) {
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const size_t goff = col0 * (size_t)nlev;
        const uint32_t flat = (uint32_t)(nvalid * col_bytes);
        const int b = i % cfg.n_op;
        char* st = op_ring + (size_t)b * op_bytes;
        wait(op_empty + b, ((i / cfg.n_op) & 1) ^ 1, 1);
        mbar_arrive_expect_tx(op_full + b, (uint32_t)(1 + NT) * flat);
        bulk_g2s(st, (const T*)a.ax.base + goff, flat, op_full + b);
#pragma unroll
        for (int k = 0; k < NT; ++k) bulk_g2s(st + (k + 1) * kPipeOpBytes, (const T*)a.ax.terms[k] + goff, flat, op_full + b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ W producer (stored W) --------------------
    if (!MATRIX_FREE) {
      if (lane == 0) role = 1;
      for (int i = 0; i < count; ++i) {
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const size_t goff = col0 * (size_t)nlev;
        const uint32_t flat = (uint32_t)(nvalid * col_bytes);
        const int b = i % cfg.n_w;
        char* sl = w_ring + (size_t)b * w_bytes;  // dl | d | du padded tiles
        if (lane == 0) {
          wait(w_empty + b, ((i / cfg.n_w) & 1) ^ 1, 3);
          mbar_arrive_expect_tx(w_full + b, 3u * flat);
        }
        __syncwarp();
        for (int c = lane; c < nvalid; c += 32) {
          const size_t g = goff + (size_t)c * nlev;
          bulk_g2s(sl + c * pitch, (const T*)a.dl + g, col_bytes, w_full + b);
          bulk_g2s(sl + tile_bytes + c * pitch, (const T*)a.d + g, col_bytes, w_full + b);
          bulk_g2s(sl + 2 * tile_bytes + c * pitch, (const T*)a.du + g, col_bytes, w_full + b);
        }
      }
    }
  } else if (warp < 2 + kPipeAsm / 32) {
    // ------------------------------------------------------------------ assemblers (phase A) ----------------------
    const int at = threadIdx.x - 64;
    if (at == 0) role = 2;
    // K/dz^2 of this thread's columns, fetched one tile ahead: under a saturated memory system a dependent global load
    // inside the tile loop costs microseconds (it queues behind the bulk copies)
    T acol_next[kRounds];
    auto fetch_acol = [&](int i) {
      size_t c0;
      int nv;
      tile_of(i, c0, nv);
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int col = (at + r * kPipeAsm) >> cshift;
        acol_next[r] = (i < count && col < nv) ? ((const T*)a.Kcol)[c0 + col] : T(0);
      }
    };
    fetch_acol(0);
    for (int i = 0; i < count; ++i) {
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const int total = nvalid * cpc;
      T acol_cur[kRounds];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) acol_cur[r] = acol_next[r];
      fetch_acol(i + 1);
      const size_t gvec0 = col0 * (size_t)nlev * sizeof(T) / 16;
      const int b = i % cfg.n_op, s = i % cfg.n_x;
      const char* st = op_ring + (size_t)b * op_bytes;
      char* xs = x_ring + (size_t)s * x_bytes;
      V* sU0 = reinterpret_cast<V*>(xs);
      char* sx = xs + kPipeOpBytes;
      wait(x_empty + s, ((i / cfg.n_x) & 1) ^ 1, 5);
      wait(op_full + b, (i / cfg.n_op) & 1, 2);
      // (the operands sit in shared memory: every round loads what it needs, no register staging across rounds)
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = at + r * kPipeAsm;
        const bool ok = q < total;
        const int col = q >> cshift, kk = q & (cpc - 1);
        V vb, vt[NTS];
        const T acol = acol_cur[r];
        if (ok) {
          vb = reinterpret_cast<const V*>(st)[q];
#pragma unroll
          for (int t = 0; t < NT; ++t) vt[t] = reinterpret_cast<const V*>(st + (t + 1) * kPipeOpBytes)[q];
        }
        T eb[VN], uu[VN], rr[VN], et[NTS][VN];
        vec_unpack<T>(vb, eb);
#pragma unroll
        for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t], et[t]);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          T x[NTS];
#pragma unroll
          for (int t = 0; t < NT; ++t) x[t] = et[t][e];
          constexpr int N1 = GroupSplit<NT, N1SPEC>::value;
          if constexpr (N1 >= 0)
            uu[e] = axpy_combine_n1<T, C, NT, (N1 >= 0 ? N1 : 0)>(eb[e], x, a.ax);
          else
            uu[e] = axpy_combine<T, C, NT>(eb[e], x, a.ax);
        }
        const V vu = vec_pack<T>(uu);
        T below = __shfl_up_sync(0xffffffffu, uu[VN - 1], 1, cpc);
        T above = __shfl_down_sync(0xffffffffu, uu[0], 1, cpc);
        if (kk == 0) below = T(0);
        if (kk == cpc - 1) above = T(0);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const T um = e > 0 ? uu[e > 0 ? e - 1 : 0] : below;
          const T up = e + 1 < VN ? uu[e + 1 < VN ? e + 1 : e] : above;
          const T ti = timp_linear<T>(acol, um, uu[e], up);
          rr[e] = (T)__dsub_rn(__dadd_rn((double)uu[e], __dmul_rn(a.dtg, (double)ti)), (double)uu[e]);
        }
        if (ok) {
          sU0[q] = vu;
          *reinterpret_cast<V*>(sx + col * pitch + kk * 16) = vec_pack<T>(rr);
          if (a.temp) reinterpret_cast<V*>(a.temp)[gvec0 + q] = vu;
        }
      }
      mbar_arrive(x_full + s);
      mbar_arrive(op_empty + b);
    }
  } else if (warp < 2 + kPipeAsm / 32 + kPipeMaxSolverWarps) {
    // ------------------------------------------------------------------ solvers (phase B) -------------------------
    const int w = warp - 2 - kPipeAsm / 32;
    const int team = w / jobs, j = w - team * jobs;
    if (team < cfg.n_teams) {
      if (w == 0 && lane == 0) role = 3;
      T ac_next = T(0);
      auto fetch_ac = [&](int i) {
        size_t c0;
        int nv;
        tile_of(i, c0, nv);
        const int col = j * 16 + lane_column(lane);
        ac_next = (MATRIX_FREE && i < count && col < nv) ? ((const T*)a.Kcol)[c0 + col] : T(0);
      };
      fetch_ac(team);
      for (int i = team; i < count; i += cfg.n_teams) {
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const T ac = ac_next;
        fetch_ac(i + cfg.n_teams);
        const int s = i % cfg.n_x;
        const int b = MATRIX_FREE ? team : i % cfg.n_w;
        char* sx = x_ring + (size_t)s * x_bytes + kPipeOpBytes;
        char* sw = w_ring + (size_t)b * w_bytes;
        wait(x_full + s, (i / cfg.n_x) & 1, 4);
        if (!MATRIX_FREE) wait(w_full + b, (i / cfg.n_w) & 1, 6);
        const int col = j * 16 + lane_column(lane);
        const bool is_up = lane_is_up(lane);
        const bool valid = col < nvalid;
        const int col_off = col * pitch;
        if (MATRIX_FREE) {
          const double o = __dmul_rn(a.dtg, (double)ac);
          ConstRows<T> rows{(T)o, (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0), sw};
          babe_solve_lanepair<T, ConstRows<T>, RhsFromTile<T>, true>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
        } else {
          StoredRows<T> rows{sw, sw + tile_bytes, sw + 2 * tile_bytes};
          babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, true>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
          fence_proxy_async_smem();  // the factors overwrote dl/du in place; the next writer of these bytes is a bulk copy
          mbar_arrive(w_empty + b);
        }
        mbar_arrive(x_solved + s);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (phase C) ------------------------
    const int et_ = threadIdx.x - 64 - kPipeAsm - 32 * kPipeMaxSolverWarps;
    if (et_ == 0) role = 4;
    const DivConst dc = div_const_prepare(a.dtg);
    const bool emit = MATRIX_FREE && a.emit_w;
    T kc_next[kRounds];
    auto fetch_kc = [&](int i) {
      size_t c0;
      int nv;
      tile_of(i, c0, nv);
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int col = (et_ + r * kPipeStore) >> cshift;
        kc_next[r] = (emit && i < count && col < nv) ? ((const T*)a.Kcol)[c0 + col] : T(0);
      }
    };
    fetch_kc(0);
    for (int i = 0; i < count; ++i) {
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const int total = nvalid * cpc;
      const size_t gvec0 = col0 * (size_t)nlev * sizeof(T) / 16;
      const int s = i % cfg.n_x;
      const char* xs = x_ring + (size_t)s * x_bytes;
      T kc_cur[kRounds];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) kc_cur[r] = kc_next[r];
      fetch_kc(i + 1);
      const V* sU0 = reinterpret_cast<const V*>(xs);
      const char* sx = xs + kPipeOpBytes;
      wait(x_solved + s, (i / cfg.n_x) & 1, 7);
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = et_ + r * kPipeStore;
        if (q >= total) continue;
        const int col = q >> cshift, kk = q & (cpc - 1);
        T e0[VN], ex[VN], eU[VN], eT[VN];
        vec_unpack<T>(sU0[q], e0);
        vec_unpack<T>(*reinterpret_cast<const V*>(sx + col * pitch + kk * 16), ex);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          eU[e] = rsub(e0[e], ex[e]);
          eT[e] = (T)div_const((double)rsub(eU[e], e0[e]), dc);
        }
        reinterpret_cast<V*>(a.U)[gvec0 + q] = vec_pack<T>(eU);
        if (a.Timp) reinterpret_cast<V*>(a.Timp)[gvec0 + q] = vec_pack<T>(eT);
        if (emit) {  // same expressions as column_wfact_kernel
          const double o = __dmul_rn(a.dtg, (double)kc_cur[r]);
          T eo[VN], ed[VN];
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            eo[e] = (T)o;
            ed[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
          }
          const V vo = vec_pack<T>(eo);
          reinterpret_cast<V*>(const_cast<void*>(a.dl))[gvec0 + q] = vo;
          reinterpret_cast<V*>(const_cast<void*>(a.d))[gvec0 + q] = vec_pack<T>(ed);
          reinterpret_cast<V*>(const_cast<void*>(a.du))[gvec0 + q] = vo;
        }
      }
      mbar_arrive(x_empty + s);
    }
  }
  if (a.dbg && role >= 0) {
    acc[0] = clock64() - t_begin;
#pragma unroll
    for (int k = 0; k < 8; ++k) a.dbg[((size_t)blockIdx.x * 5 + role) * 8 + k] = acc[k];
  }
}

constexpr int kSpData = 128;  // data threads per lane: kKeep vectors of a 512-vector tile each
#ifndef CTS_SP_PINGPONG
#define CTS_SP_PINGPONG true
#endif
constexpr int kSpMaxLanes = 3;
constexpr int kSpHeader = 256;  // mbarriers of all lanes

// ---------------------------------------------------------------------------------------------
// Software-pipelined, TMA-only variant of the fused stage: one persistent CTA per SM runs `nlanes` independent pipelines
// ("lanes"), each made of dedicated solver warps and 128 data threads, and EVERY byte of global traffic moves by bulk TMA.
//
// Why (all measured on B200, scratch/solve_latency.cu, solve_contention.cu, solve_noise.cu):
//   * the two-sided product-form Thomas solve of a 16-column tile takes 3400-4000 cycles for a warp that has its SM
//     sub-partition to itself, 4900 with two solver warps on one sub-partition, 7700 with four;
//   * next to warps that stream global memory with LDG/STG the same solve takes 10 000 - 180 000 cycles: its LDS/STS queue up
//     behind the global loads/stores in the sub-partition's memory-instruction path; next to bulk-TMA copies that move the
//     same bytes (6.6 TB/s) it takes 6200 -- the async proxy does not go through that path.  This is what separates the
//     fused stage (0.84-0.90 of the copy bandwidth with LDG/STG, ncu: 39 % long-scoreboard + 34 % barrier stalls) from K4
//     (0.98, TMA only).
// So: operands (base + NT tendencies) arrive as flat 8 KiB bulk copies in a staging area, W as one 512-byte copy per column
// into padded tiles; the data threads combine them in shared memory -> registers (K1), write U0 and the K3 residual to shared
// memory, and after the solve form U and T_imp IN PLACE in shared memory, from where bulk stores take them to HBM.  A lane
// keeps two tiles in flight: while the solver warp works on tile i the operands of tile i+1 land and are combined, and the
// operands of tile i+2 are requested.  Solver warps are the CTA's first warps (hardware warp ids of a CTA's warps 0..3 are a
// rotation of one group of four = one warp per sub-partition).  Waiting is done at hardware named barriers (a warp parked
// at bar.sync issues nothing); mbarriers are only polled for TMA arrivals that have normally long happened.
// Same device functions, operation order and roundings as the kernels above: bit-identical results.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpShape {
  static constexpr int kLaneThreads = TileShape<T>::kThreads + kSpData;
  static constexpr int kMaxThreads = kSpMaxLanes * kLaneThreads;
};

template <typename T, typename C, int NT, bool MATRIX_FREE, int N1SPEC>
__global__ void __launch_bounds__(SpShape<T>::kMaxThreads, 1) column_fused_stage_sp_kernel(const FusedArgs<NT> a, int nlanes) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NT > 0 ? NT : 1;
  constexpr int kCols = TileShape<T>::kCols;
  constexpr int kSolvers = TileShape<T>::kThreads;
  constexpr int kRounds = kKeep;
  constexpr int kFlat = kRounds * kSpData * 16;  // one flat operand / output tile (8 KiB)
  extern __shared__ __align__(128) char smem_all[];
  // role of this thread: solver warps come first (CTA warps 0..nlanes*kSolvers/32-1), then the data threads
  const bool is_solver = (int)threadIdx.x < nlanes * kSolvers;
  const int swarp = threadIdx.x >> 5;
  const int lane_id = is_solver ? swarp % nlanes : ((int)threadIdx.x - nlanes * kSolvers) / kSpData;  // pipeline of this thread
  const int pitch = a.pitch;
  const int tile_bytes = kCols * pitch;
  const int lane_bytes = (MATRIX_FREE ? 2 : 4) * tile_bytes + kFlat + (1 + NT) * kFlat + (MATRIX_FREE ? 3 * kFlat : 0) + 256;
  char* smem = smem_all + kSpHeader + (size_t)lane_id * lane_bytes;
  char* sx = smem;                 // residual -> rp -> dx -> T_imp (out)
  char* sw = smem + tile_bytes;    // stored: dl | d | du tiles; matrix-free: cp scratch
  char* sU0c = smem + (MATRIX_FREE ? 2 : 4) * tile_bytes;  // U0 -> U (out), flat
  V* sU0 = reinterpret_cast<V*>(sU0c);
  char* sOP = sU0c + kFlat;                                 // operand staging: base | terms[0..NT-1], flat
  char* sWo = sOP + (1 + NT) * kFlat;                       // matrix-free + emit_w: dl | d | du to be written, flat
  T* sK = reinterpret_cast<T*>(sWo + (MATRIX_FREE ? 3 * kFlat : 0));  // K/dz^2 of the tile's columns, two tiles: [2][kCols]
  uint64_t* op_full = reinterpret_cast<uint64_t*>(smem_all) + 2 * lane_id;      // operands(i) have landed
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem_all) + 2 * lane_id + 1;   // W(i) has landed
  // hardware named barriers of this lane
  const int kBarR = 1 + 4 * lane_id;   // residual(i) ready: data arrive, solver sync   (kSolvers + 128 threads)
  const int kBarS = 2 + 4 * lane_id;   // dx(i) ready:       solver arrive, data sync  (kSolvers + 128 threads)
  const int kBarD = 3 + 4 * lane_id;   // data threads only (128)
  const int kBarV = 4 + 4 * lane_id;   // solver warps only (kSolvers), Float32 tiles
  constexpr int kBarThreads = kSolvers + kSpData;
  auto bar_arrive = [](int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kBarThreads) : "memory"); };
  auto bar_sync = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kBarThreads) : "memory"); };
  auto bar_data = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSpData) : "memory"); };

  const int nlev = a.nlev;
  const int col_bytes = nlev * (int)sizeof(T);
  const int cpc = col_bytes / 16;
  const int cshift = 31 - __clz(cpc);
  const size_t ntiles = (a.ncols + kCols - 1) / kCols;
  const size_t first = (size_t)blockIdx.x * nlanes + lane_id, stride = (size_t)gridDim.x * nlanes;
  const int count = ntiles > first ? (int)((ntiles - first + stride - 1) / stride) : 0;
  auto tile_of = [&](int i, size_t& col0, int& nvalid) {
    const size_t tile = first + (size_t)i * stride;
    col0 = tile * kCols;
    nvalid = (int)((a.ncols - col0) < (size_t)kCols ? (a.ncols - col0) : (size_t)kCols);
  };
  // optional wait-time profile (CTS_FUSED_TIMELINE): the first solver thread and the first data thread of lane 0
  long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool prof = a.dbg != nullptr && (threadIdx.x == 0 || (int)threadIdx.x == nlanes * kSolvers);
  const long long t_begin = prof ? clock64() : 0;
  auto wait_mbar = [&](uint64_t* bar, uint32_t parity, int tag) {  // (all lanes of a warp take the same path)
    const long long t0 = a.dbg ? clock64() : 0;
    mbar_wait_wd(bar, parity, tag);
    if (prof) acc[tag] += clock64() - t0;
  };
  auto wait_bar = [&](int id, int tag) {
    const long long t0 = a.dbg ? clock64() : 0;
    bar_sync(id);
    if (prof) acc[tag] += clock64() - t0;
  };
  if (threadIdx.x == 0) {
    for (int l = 0; l < 2 * nlanes; ++l) mbar_init(reinterpret_cast<uint64_t*>(smem_all) + l, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (is_solver) {
    // ------------------------------------------------------------------ solver warps -----------------------------
    const int lane = threadIdx.x & 31, warp = swarp / nlanes;  // warp = 16-column job of this lane's tile
    const int col = warp * 16 + lane_column(lane);
    const bool is_up = lane_is_up(lane);
    const int col_off = col * pitch;
    auto issue_w = [&](int i) {  // job 0: W(i) -> sw, one 1-D bulk copy per column and diagonal
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const size_t goff = col0 * (size_t)nlev;
      if (lane == 0) mbar_arrive_expect_tx(w_full, (uint32_t)(3 * nvalid * col_bytes));
      __syncwarp();
      for (int c = lane; c < nvalid; c += 32) {
        const size_t g = goff + (size_t)c * nlev;
        bulk_g2s(sw + c * pitch, (const T*)a.dl + g, col_bytes, w_full);
        bulk_g2s(sw + tile_bytes + c * pitch, (const T*)a.d + g, col_bytes, w_full);
        bulk_g2s(sw + 2 * tile_bytes + c * pitch, (const T*)a.du + g, col_bytes, w_full);
      }
    };
    if (!MATRIX_FREE && warp == 0 && count > 0) issue_w(0);
    for (int i = 0; i < count; ++i) {
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const bool valid = col < nvalid;
      wait_bar(kBarR, 1);
      if (MATRIX_FREE) {
        const T ac = valid ? sK[(i & 1) * kCols + col] : T(0);
        const double o = __dmul_rn(a.dtg, (double)ac);
        ConstRows<T> rows{(T)o, (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0), sw};
        const long long ts = a.dbg ? clock64() : 0;
        babe_solve_lanepair<T, ConstRows<T>, RhsFromTile<T>, CTS_SP_PINGPONG>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
        if (prof) acc[4] += clock64() - ts;
        bar_arrive(kBarS);
      } else {
        wait_mbar(w_full, i & 1, 2);
        StoredRows<T> rows{sw, sw + tile_bytes, sw + 2 * tile_bytes};
        const long long ts = a.dbg ? clock64() : 0;
        babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, CTS_SP_PINGPONG>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
        if (prof) acc[4] += clock64() - ts;
        fence_proxy_async_smem();  // the factors overwrote dl/du in place; the next writer of these bytes is a bulk copy
        bar_arrive(kBarS);
        if (i + 1 < count) {
          if (kSolvers > 32)
            asm volatile("bar.sync %0, %1;" ::"r"(kBarV), "n"(kSolvers) : "memory");  // every solver warp of the lane is done with W(i)
          else
            __syncwarp();
          if (warp == 0) issue_w(i + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ data threads -----------------------------
    const int dt = ((int)threadIdx.x - nlanes * kSolvers) % kSpData;
    const DivConst dc = div_const_prepare(a.dtg);
    const bool emit = MATRIX_FREE && a.emit_w;
    V U0r[kRounds];
    auto issue_ops = [&](int i) {  // one thread: base + terms of tile i -> sOP, one flat bulk copy each
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const size_t goff = col0 * (size_t)nlev;
      const uint32_t flat = (uint32_t)(nvalid * col_bytes);
      mbar_arrive_expect_tx(op_full, (uint32_t)(1 + NT) * flat);
      bulk_g2s(sOP, (const T*)a.ax.base + goff, flat, op_full);
#pragma unroll
      for (int k = 0; k < NT; ++k) bulk_g2s(sOP + (k + 1) * kFlat, (const T*)a.ax.terms[k] + goff, flat, op_full);
    };
    auto fetch_K = [&](int i) {  // K/dz^2 of tile i's columns -> sK[i & 1] (one scalar load per column)
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      if (dt < nvalid) sK[(i & 1) * kCols + dt] = ((const T*)a.Kcol)[col0 + dt];
    };
    auto combine = [&](int i) {  // U0(i) = K1(base, terms): staging -> registers
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const int total = nvalid * cpc;
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = dt + r * kSpData;
        V vb = V{}, vt[NTS];
#pragma unroll
        for (int t = 0; t < NT; ++t) vt[t] = V{};
        if (q < total) {
          vb = reinterpret_cast<const V*>(sOP)[q];
#pragma unroll
          for (int t = 0; t < NT; ++t) vt[t] = reinterpret_cast<const V*>(sOP + (t + 1) * kFlat)[q];
        }
        T eb[VN], uu[VN], et[NTS][VN];
        vec_unpack<T>(vb, eb);
#pragma unroll
        for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t], et[t]);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          T x[NTS];
#pragma unroll
          for (int t = 0; t < NT; ++t) x[t] = et[t][e];
          constexpr int N1 = GroupSplit<NT, N1SPEC>::value;
          if constexpr (N1 >= 0)
            uu[e] = axpy_combine_n1<T, C, NT, (N1 >= 0 ? N1 : 0)>(eb[e], x, a.ax);
          else
            uu[e] = axpy_combine<T, C, NT>(eb[e], x, a.ax);
        }
        U0r[r] = vec_pack<T>(uu);
      }
    };
    if (count > 0) {
      if (dt == 0) issue_ops(0);
      fetch_K(0);
      wait_mbar(op_full, 0, 4);
      combine(0);
      bar_data(kBarD);  // every data thread has read the staging area (and sK[0] is visible)
      if (dt == 0 && count > 1) issue_ops(1);
    }
    for (int i = 0; i < count; ++i) {
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const int total = nvalid * cpc;
      const size_t goff = col0 * (size_t)nlev;
      const size_t gvec0 = goff * sizeof(T) / 16;
      // U0(i) and the K3 residual f = (U0 + dtγ T_imp(U0)) - U0 (level neighbours by warp shuffles) -> shared memory
      const long long tr = a.dbg ? clock64() : 0;
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = dt + r * kSpData;
        const bool ok = q < total;
        const int col = q >> cshift, kk = q & (cpc - 1);
        const T acol = ok ? sK[(i & 1) * kCols + col] : T(0);
        T uu[VN], rr[VN];
        vec_unpack<T>(U0r[r], uu);
        T below = __shfl_up_sync(0xffffffffu, uu[VN - 1], 1, cpc);
        T above = __shfl_down_sync(0xffffffffu, uu[0], 1, cpc);
        if (kk == 0) below = T(0);
        if (kk == cpc - 1) above = T(0);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const T um = e > 0 ? uu[e > 0 ? e - 1 : 0] : below;
          const T up = e + 1 < VN ? uu[e + 1 < VN ? e + 1 : e] : above;
          const T ti = timp_linear<T>(acol, um, uu[e], up);
          rr[e] = (T)__dsub_rn(__dadd_rn((double)uu[e], __dmul_rn(a.dtg, (double)ti)), (double)uu[e]);
        }
        if (ok) {
          sU0[q] = U0r[r];
          *reinterpret_cast<V*>(sx + col * pitch + kk * 16) = vec_pack<T>(rr);
          if (a.temp) reinterpret_cast<V*>(a.temp)[gvec0 + q] = U0r[r];
          if (emit) {  // W = dtγJ - I of the stored copy (same expressions as column_wfact_kernel), staged for a bulk store
            const double o = __dmul_rn(a.dtg, (double)acol);
            T eo[VN], ed[VN];
#pragma unroll
            for (int e = 0; e < VN; ++e) {
              eo[e] = (T)o;
              ed[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
            }
            const V vo = vec_pack<T>(eo);
            reinterpret_cast<V*>(sWo)[q] = vo;
            reinterpret_cast<V*>(sWo + kFlat)[q] = vec_pack<T>(ed);
            reinterpret_cast<V*>(sWo + 2 * kFlat)[q] = vo;
          }
        }
      }
      bar_arrive(kBarR);
      if (prof) acc[1] += clock64() - tr;
      if (i + 1 < count) {  // while the solver works on tile i: the operands of tile i+1 (long landed) -> U0(i+1) in registers
        const long long t0 = a.dbg ? clock64() : 0;
        fetch_K(i + 1);
        wait_mbar(op_full, (i + 1) & 1, 4);
        combine(i + 1);
        bar_data(kBarD);
        if (dt == 0 && i + 2 < count) issue_ops(i + 2);
        if (prof) acc[5] += clock64() - t0;
      }
      wait_bar(kBarS, 3);
      const long long te = a.dbg ? clock64() : 0;
      // epilogue in place: sU0 <- U = U0 - dx ; sx <- T_imp = (U - U0)/dtγ
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = dt + r * kSpData;
        if (q >= total) continue;
        const int col = q >> cshift, kk = q & (cpc - 1);
        V* px = reinterpret_cast<V*>(sx + col * pitch + kk * 16);
        T e0[VN], ex[VN], eU[VN], eT[VN];
        vec_unpack<T>(sU0[q], e0);
        vec_unpack<T>(*px, ex);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          eU[e] = rsub(e0[e], ex[e]);
          eT[e] = (T)div_const((double)rsub(eU[e], e0[e]), dc);
        }
        sU0[q] = vec_pack<T>(eU);
        *px = vec_pack<T>(eT);
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk stores
      bar_data(kBarD);
      if (prof) acc[2] += clock64() - te;
      {
        const long long t0 = a.dbg ? clock64() : 0;
        const uint32_t flat = (uint32_t)(nvalid * col_bytes);
        bool issued = false;
        if (dt == 0) {
          bulk_s2g((T*)a.U + goff, sU0c, flat);
          if (emit) {
            bulk_s2g((T*)const_cast<void*>(a.dl) + goff, sWo, flat);
            bulk_s2g((T*)const_cast<void*>(a.d) + goff, sWo + kFlat, flat);
            bulk_s2g((T*)const_cast<void*>(a.du) + goff, sWo + 2 * kFlat, flat);
          }
          issued = true;
        }
        if (a.Timp && dt < nvalid) {
          bulk_s2g((T*)a.Timp + goff + (size_t)dt * nlev, sx + dt * pitch, col_bytes);
          issued = true;
        }
        if (issued) {
          bulk_commit();
          bulk_wait_read0();  // the stores have read shared memory: sU0 / sx / sWo may be rewritten
        }
        bar_data(kBarD);
        if (prof) acc[6] += clock64() - t0;
      }
    }
  }
  if (prof) {
    acc[0] = clock64() - t_begin;
    unsigned hw;
    asm("mov.u32 %0, %%warpid;" : "=r"(hw));
    acc[7] = (long long)(hw % 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.dbg[((size_t)blockIdx.x * 2 + (threadIdx.x == 0 ? 0 : 1)) * 8 + j] = acc[j];
  }
}

// lanes of the TMA-only kernel that fit one SM for this launch (0: none)
static int sp_lanes(int nt, int tile_bytes, bool mf, size_t* smem_out) {
  const size_t flat = (size_t)kKeep * kSpData * 16;
  const size_t lane = (size_t)(mf ? 2 : 4) * tile_bytes + flat + (size_t)(1 + nt) * flat + (mf ? 3 * flat : 0) + 256;
  int lanes = (int)std::min<size_t>(kSpMaxLanes, ((size_t)227 * 1024 - kSpHeader) / lane);
  static const int env_lanes = getenv("CTS_SP_LANES") ? atoi(getenv("CTS_SP_LANES")) : 0;  // developer aid
  if (env_lanes > 0) lanes = std::min(lanes, env_lanes);
  *smem_out = kSpHeader + (size_t)lanes * lane;
  return lanes;
}

// ring depths of the pipelined kernel for this launch (n_op == 0: does not fit).  CTS_PIPE_OP / CTS_PIPE_X / CTS_PIPE_TEAMS
// override the defaults (developer aid for tuning).
static PipeCfg pipe_config(int nt, int tile_bytes, bool mf, int jobs) {
  const long budget = 227 * 1024 - kPipeBarBytes;
  const long op_b = (long)(1 + nt) * kPipeOpBytes, w_b = (long)(mf ? 1 : 3) * tile_bytes, x_b = kPipeOpBytes + tile_bytes;
  static const int env_op = getenv("CTS_PIPE_OP") ? atoi(getenv("CTS_PIPE_OP")) : 0;
  static const int env_x = getenv("CTS_PIPE_X") ? atoi(getenv("CTS_PIPE_X")) : 0;
  static const int env_teams = getenv("CTS_PIPE_TEAMS") ? atoi(getenv("CTS_PIPE_TEAMS")) : 0;
  PipeCfg c{0, 0, 0, 0};
  const int max_teams = kPipeMaxSolverWarps / jobs;
  int teams = std::min(max_teams, mf ? 4 : 3);
  if (env_teams > 0) teams = std::min(env_teams, max_teams);
  if (teams < 1) return c;
  const int n_w = teams;
  int n_x = env_x > 0 ? std::max(env_x, teams) : teams;
  int n_op = env_op > 0 ? env_op : 2;
  n_x = std::min(n_x, kPipeMaxDepth);
  n_op = std::min(n_op, kPipeMaxDepth);
  if (n_op * op_b + n_w * w_b + n_x * x_b > budget) return c;
  if (env_op <= 0)  // operands first: they are the bulk of the bytes in flight
    while (n_op < 4 && (n_op + 1) * op_b + n_w * w_b + n_x * x_b <= budget) ++n_op;
  if (env_x <= 0)
    while (n_x < kPipeMaxDepth && n_op * op_b + n_w * w_b + (n_x + 1) * x_b <= budget) ++n_x;
  c.n_op = n_op;
  c.n_w = n_w;
  c.n_x = n_x;
  c.n_teams = teams;
  return c;
}
static size_t pipe_smem_bytes(int nt, int tile_bytes, bool mf, const PipeCfg& c) {
  return (size_t)kPipeBarBytes + (size_t)c.n_op * (1 + nt) * kPipeOpBytes + (size_t)c.n_w * (mf ? 1 : 3) * tile_bytes +
         (size_t)c.n_x * (kPipeOpBytes + tile_bytes);
}

template <typename T, typename C, int NT>
int launch_fused(cts_op_column* op, const cts_fused_stage_args* s) {
  cts_ctx* ctx = op->ctx;
  FusedArgs<NT> a;
  a.ax.base = s->base;
  a.ax.out = nullptr;
  a.ax.out2 = nullptr;
  a.ax.c0 = s->c0;
  a.ax.n1 = s->n_grp1;
  a.ax.scale_base = s->c0 != 1.0;
  for (int k = 0; k < NT; ++k) {
    a.ax.terms[k] = s->terms[k];
    a.ax.coef[k] = s->coef[k];
  }
  a.U = s->U;
  a.Timp = s->T_imp;
  a.temp = s->temp;
  a.Kcol = op->a_col;
  a.dtg = s->dtgamma;
  a.dz = op->d.dz;
  a.nlev = op->d.nlev;
  a.ncols = op->ncols;
  a.pitch = pitch_bytes(op->d.nlev, sizeof(T));
  const bool mf = op->matrix_free || s->emit_W;
  a.emit_w = s->emit_W;
  // base + terms in, U (+ T_imp, temp) out, dl/d/du read -- or written by the stage that folds Wfact in
  CTS_COUNT_BYTES(ctx, 1 + NT + 1 + (s->T_imp ? 1 : 0) + (s->temp ? 1 : 0) + ((!op->matrix_free) ? 3 : 0), cts_op_column_ndof(op), sizeof(T));
  if (!op->matrix_free) {
    if (!s->W) return cts_fail(ctx, CTS_EINVAL, "fused stage: stored W required");
    a.dl = s->W->dl;
    a.d = s->W->d;
    a.du = s->W->du;
  } else {
    a.dl = a.d = a.du = nullptr;
    a.emit_w = 0;
  }
  size_t tiles = (op->ncols + kTileCols - 1) / kTileCols;
  static const bool timeline = getenv("CTS_FUSED_TIMELINE") != nullptr;  // developer aid: per-phase clock64 averages
  static const bool no_reg = getenv("CTS_FUSED_NO_REG") != nullptr;      // developer aid: force the 5-tile kernel
  const int cpc = op->d.nlev * (int)sizeof(T) / 16;
  constexpr int kCols = TileShape<T>::kCols;
  // cluster variant (loader CTA + solver CTA on two SMs, fused_cluster.cu): Float64, 64-level columns; opt-in
  static const bool use_cluster = getenv("CTS_FUSED_CLUSTER") != nullptr && getenv("CTS_FUSED_CLUSTER")[0] == '1';
  if constexpr (std::is_same<T, double>::value && std::is_same<C, double>::value) {
    if (use_cluster && !no_reg && cpc == 32) {
      int cspec = 0;
      if constexpr (NT >= 1) {
        if (a.ax.n1 == NT)
          cspec = 1;
        else if (a.ax.n1 == (NT + 1) / 2)
          cspec = 2;
      }
      const int rc = cts_fused_cluster_f64(ctx, NT, cspec, mf, &a, timeline);
      if (rc != CTS_EUNSUPPORTED) return rc;
    }
  }
  // pipelined persistent kernel: a tile is 512 vectors, a column 8..32 of them (a power of two)
  // The pipelined kernel is opt-in (CTS_FUSED_PIPE=1): bit-identical, but measured slower on B200
  static const bool no_pipe = getenv("CTS_FUSED_PIPE") == nullptr || getenv("CTS_FUSED_PIPE")[0] == '0';
  if constexpr (NT <= 5) {
  if (!no_reg && !no_pipe && (cpc == 16 || cpc == 32) && cpc * kCols == 512) {
    const int pcols = kPipeTileVecs / cpc;
    const int ptile = pcols * a.pitch;
    const PipeCfg cfg = pipe_config(NT, ptile, mf, pcols / 16);
    const size_t ptiles = (op->ncols + pcols - 1) / pcols;
    if (cfg.n_op >= 2 && ptiles >= 1) {
      const size_t psmem = pipe_smem_bytes(NT, ptile, mf, cfg);
      const unsigned grid = (unsigned)std::min<size_t>(ptiles, (size_t)ctx->num_sms);
      a.dbg = nullptr;
      if (timeline) {
        CTS_CUDA(ctx, cudaMalloc(&a.dbg, (size_t)grid * 40 * sizeof(long long)));
        CTS_CUDA(ctx, cudaMemsetAsync(a.dbg, 0, (size_t)grid * 40 * sizeof(long long), ctx->stream));
      }
      auto launchp = [&](auto k) -> int {
        CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        k<<<grid, kPipeThreads, psmem, ctx->stream>>>(a, cfg);
        return CTS_OK;
      };
      int spec = 0;
      if constexpr (NT >= 1) {
        if (a.ax.n1 == NT)
          spec = 1;
        else if (a.ax.n1 == (NT + 1) / 2)
          spec = 2;
      }
      if (spec == 1) {
        if (mf) CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, true, 1>));
        else CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, false, 1>));
      } else if (spec == 2) {
        if (mf) CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, true, 2>));
        else CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, false, 2>));
      } else {
        if (mf) CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, true, 0>));
        else CTS_TRY(launchp(column_fused_stage_pipe_kernel<T, C, NT, false, 0>));
      }
      CTS_LAUNCH_CHECK(ctx);
      if (timeline) {
        std::vector<long long> h((size_t)grid * 40);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(a.dbg);
        static const char* roles[5] = {"op-producer", "W-producer", "assemblers", "solver0", "epilogue"};
        static const char* tags[8] = {"total", "op_empty", "op_full", "w_empty", "x_full", "x_empty", "w_full", "x_solved"};
        fprintf(stderr, "[cts pipe timeline] NT=%d mf=%d tiles=%zu grid=%u n_op=%d n_w=%d n_x=%d teams=%d smem=%zu\n", NT, (int)mf, ptiles, grid,
                cfg.n_op, cfg.n_w, cfg.n_x, cfg.n_teams, psmem);
        for (int r = 0; r < 5; ++r) {
          double sum[8] = {0};
          for (unsigned bk = 0; bk < grid; ++bk)
            for (int k = 0; k < 8; ++k) sum[k] += (double)h[((size_t)bk * 5 + r) * 8 + k];
          if (sum[0] == 0) continue;
          fprintf(stderr, "  %-12s total %.0f cycles/CTA; wait share:", roles[r], sum[0] / grid);
          for (int k = 1; k < 8; ++k)
            if (sum[k] > 0) fprintf(stderr, " %s %.1f%%", tags[k], 100.0 * sum[k] / sum[0]);
          fprintf(stderr, "\n");
        }
      }
      return CTS_OK;
    }
  }
  }
  const bool no_pipe_on = !no_pipe && (cpc == 16 || cpc == 32) && cpc * kCols == 512;  // (the opt-in ring-pipelined kernel above took the launch)
  // software-pipelined TMA-only kernel (one persistent CTA per SM, lanes of solver warps + data threads): opt-in
  // (CTS_FUSED_SP=1) -- bit-identical, 5.7-5.9 ms of fused stages per step against 4.9-5.1 ms for the register kernel below
  static const bool no_sp = getenv("CTS_FUSED_SP") == nullptr || getenv("CTS_FUSED_SP")[0] == '0';
  if constexpr (NT <= 5) {
  size_t sp_probe = 0;
  if (!no_reg && !no_sp && !no_pipe_on && cpc >= 1 && cpc * kCols == kKeep * kSpData && (cpc & (cpc - 1)) == 0 &&
      sp_lanes(NT, kCols * a.pitch, mf, &sp_probe) >= 1) {
    const size_t sp_tiles = (op->ncols + kCols - 1) / kCols;
    size_t smem_sp = 0;
    const int sp_nl = sp_lanes(NT, kCols * a.pitch, mf, &smem_sp);
    a.dbg = nullptr;
    unsigned sp_grid = 0;
    auto launch_sp = [&](auto k) -> int {
      CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sp));
      const unsigned grid = (unsigned)std::min<size_t>((sp_tiles + sp_nl - 1) / sp_nl, (size_t)ctx->num_sms);
      sp_grid = grid;
      if (timeline) {
        CTS_CUDA(ctx, cudaMalloc(&a.dbg, (size_t)grid * 16 * sizeof(long long)));
        CTS_CUDA(ctx, cudaMemsetAsync(a.dbg, 0, (size_t)grid * 16 * sizeof(long long), ctx->stream));
      }
      k<<<grid, sp_nl * SpShape<T>::kLaneThreads, smem_sp, ctx->stream>>>(a, sp_nl);
      return CTS_OK;
    };
    int spec = 0;
    if constexpr (NT >= 1) {
      if (a.ax.n1 == NT)
        spec = 1;
      else if (a.ax.n1 == (NT + 1) / 2)
        spec = 2;
    }
    if (spec == 1) {
      if (mf) CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, true, 1>));
      else CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, false, 1>));
    } else if (spec == 2) {
      if (mf) CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, true, 2>));
      else CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, false, 2>));
    } else {
      if (mf) CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, true, 0>));
      else CTS_TRY(launch_sp(column_fused_stage_sp_kernel<T, C, NT, false, 0>));
    }
    CTS_LAUNCH_CHECK(ctx);
    if (timeline) {
      std::vector<long long> h((size_t)sp_grid * 16);
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaFree(a.dbg);
      static const char* roles[2] = {"solver", "data"};
      static const char* tags0[8] = {"total", "wait residual", "wait w_full", "-", "SOLVE", "-", "-", "-"};
      static const char* tags1[8] = {"total", "residual->smem", "epilogue+fence+sync", "wait solved", "wait op_full", "prefetch+combine(i+1)", "store+wait_read", "-"};
      fprintf(stderr, "[cts sp timeline] NT=%d mf=%d tiles=%zu grid=%u lanes=%d smem=%zu\n", NT, (int)mf, sp_tiles, sp_grid, sp_nl, smem_sp);
      for (int r = 0; r < 2; ++r) {
        double sum[8] = {0};
        for (unsigned bk = 0; bk < sp_grid; ++bk)
          for (int j = 0; j < 8; ++j) sum[j] += (double)h[((size_t)bk * 2 + r) * 8 + j];
        fprintf(stderr, "  %-7s total %.0f cycles/CTA (%.0f per tile of a lane);", roles[r], sum[0] / sp_grid, sum[0] / sp_grid / ((double)sp_tiles / sp_grid / sp_nl));
        for (int j = 1; j < 7; ++j)
          if (sum[j] > 0) fprintf(stderr, " %s %.1f%%", (r == 0 ? tags0 : tags1)[j], 100.0 * sum[j] / sum[0]);
        int hist[4] = {0, 0, 0, 0};
        for (unsigned bk = 0; bk < sp_grid; ++bk) hist[h[((size_t)bk * 2 + r) * 8 + 7] & 3]++;
        fprintf(stderr, "  [hw warpid %% 4 of the reporting warp: %d %d %d %d]\n", hist[0], hist[1], hist[2], hist[3]);
      }
    }
    return CTS_OK;
  }
  }
  if (!timeline && !no_reg && cpc >= 1 && cpc * kCols <= KeepOf<T>::value * kFusedThreads && (cpc & (cpc - 1)) == 0) {
    tiles = (op->ncols + kCols - 1) / kCols;
    size_t smem_reg = (size_t)(mf ? 2 : 4) * kCols * a.pitch + 16 + ((kCols * sizeof(T) + 15) / 16) * 16 +
                      (size_t)TileShape<T>::kThreads * (mf ? FusedPark<T, true>::value : FusedPark<T, false>::value) * 16;  // tiles + mbarrier + K/dz^2 + parked U0 of the solver threads
    a.dbg = nullptr;
    auto launch = [&](auto k) -> int {
      CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_reg));
      k<<<(unsigned)tiles, kFusedThreads, smem_reg, ctx->stream>>>(a);
      return CTS_OK;
    };
    // kernels specialised for the two usual operand splits exist up to 5 operands (ARS343's largest stage)
    int spec = 0;
    if constexpr (NT >= 1 && NT <= 5) {
      if (a.ax.n1 == NT)
        spec = 1;
      else if (a.ax.n1 == (NT + 1) / 2)
        spec = 2;
    }
    if constexpr (NT >= 1 && NT <= 5) {
      if (spec == 1) {
        if (mf) CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, true, 1>));
        else CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, false, 1>));
      } else if (spec == 2) {
        if (mf) CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, true, 2>));
        else CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, false, 2>));
      }
    }
    if (spec == 0) {
      if (mf) CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, true, 0>));
      else CTS_TRY(launch(column_fused_stage_reg_kernel<T, C, NT, false, 0>));
    }
    CTS_LAUNCH_CHECK(ctx);
    return CTS_OK;
  }
  size_t smem = (size_t)(mf ? 3 : 5) * kTileCols * a.pitch + 16 + kTileCols * sizeof(T);
  a.dbg = nullptr;
  if (timeline) CTS_CUDA(ctx, cudaMalloc(&a.dbg, tiles * 5 * sizeof(long long)));
  if (mf) {
    auto k = column_fused_stage_kernel<T, C, NT, true>;
    CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)tiles, kFusedThreads, smem, ctx->stream>>>(a);
  } else {
    auto k = column_fused_stage_kernel<T, C, NT, false>;
    CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)tiles, kFusedThreads, smem, ctx->stream>>>(a);
  }
  CTS_LAUNCH_CHECK(ctx);
  if (timeline) {
    std::vector<long long> h(tiles * 5);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(a.dbg);
    double ph[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < tiles; ++i)
      for (int k = 0; k < 4; ++k) ph[k] += (double)(h[i * 5 + k + 1] - h[i * 5 + k]);
    fprintf(stderr, "[cts fused timeline] NT=%d mf=%d tiles=%zu avg cycles: A(load+K1)=%.0f residual=%.0f solve=%.0f C(store)=%.0f\n", NT,
            (int)mf, tiles, ph[0] / tiles, ph[1] / tiles, ph[2] / tiles, ph[3] / tiles);
  }
  return CTS_OK;
}

// ---------------------------------------------------------------------------------------------
// "Right-hand side + column solve" tile kernel: x = W \ rhs, where rhs is formed on the fly from level vectors in
// registers (the operator's level neighbours come from warp shuffles) and never goes to HBM.  Two uses:
//   MODE_ROS  Rosenbrock stage (rosenbrock.jl:164-224):  fU = 0; fU += T_imp(U) [; fU += fU_exp] [; fU += c_j k_j ...];
//             fU *= -dtγ;  k_i = W \ fU            -- T_imp! + the `.+=` chain + ldiv! in one pass
//             ((1 [+1] + nt + 3) w read, 1w written instead of 2w + (nt + 2 [+1]) w + 5w)
//   MODE_JVP  left-preconditioned Jacobian-free matvec of GMRES (newtons_method.jl:173-182,497):
//             q = W \ ((f(x + eps v) - f(x))/eps)  -- K7 + T_imp! + K3 + K8 + ldiv!: 8w instead of 5w + 5w
// Same device functions and roundings as the stand-alone kernels (axpy_chain_vec_kernel, column_residual_kernel,
// tridiag_tile_kernel): bit-identical results.  A CTA owns TileShape<T>::kCols columns (one contiguous range); W arrives
// by bulk TMA while the right-hand side is assembled; the solve is the two-sided product-form Thomas of tile_device.cuh.
// ---------------------------------------------------------------------------------------------
//   MODE_PRE  (matrix-free W only) the first preconditioner solve of a GMRES call, q = W(x) \ b (newtons_method.jl:484):
//             3w instead of 5w
// WGEN ("matrix-free W", cts_op_column_set_matrix_free): the rows of W = dtγ J(x) - I are computed from x inside the kernel
// with the expressions of column_wfact_kernel (frozen coefficients at x for the nonlinear operator) instead of being read:
// 3w less per solve and no Wfact pass -- valid whenever the stored W would have been refreshed at this very x.
constexpr int MODE_ROS = 0, MODE_JVP = 1, MODE_PRE = 2;
constexpr int kRhsMaxTerms = 3;  // chain terms besides T_imp(U) and fU_exp (SSPKnoth needs 2)
struct RhsSolveArgs {
  void* out;
  const void* x;     // ROS: U          JVP: x
  const void* v;     // ROS: fU_exp|0   JVP: Krylov vector
  const void* U0;    // JVP: U0
  const void* f;     // JVP: f(x)
  const void* terms[kRhsMaxTerms];  // ROS: k_j of the C[i, j]/dt terms
  double coef[kRhsMaxTerms];
  int nt;
  double scale;      // ROS: -dtγ
  double eps;        // JVP
  double dtg;
  const void* acol;
  const void* dl;
  const void* d;
  const void* du;
  int nlev;
  size_t ncols;
  int pitch;
};

#ifndef CTS_RHS_SOLVE_MIN_BLOCKS
#define CTS_RHS_SOLVE_MIN_BLOCKS 6  // four tiles + barrier = 33 808 B of shared memory: six CTAs per SM at <= 80 registers
#endif
template <typename T, int MODE, bool NONLIN, int NTC, bool WGEN = false>
__global__ void __launch_bounds__(kFusedThreads, CTS_RHS_SOLVE_MIN_BLOCKS) column_rhs_solve_kernel(const RhsSolveArgs a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int NTS = NTC > 0 ? NTC : 1;
  constexpr int kCols = TileShape<T>::kCols;
  constexpr int kSolvers = TileShape<T>::kThreads;
  constexpr int KEEP = KeepOf<T>::value;
  extern __shared__ __align__(128) char smem[];
  const int pitch = a.pitch;
  const int tile_bytes = kCols * pitch;
  char* sx = smem;
  char* sdl = smem + tile_bytes;
  char* sd = smem + 2 * tile_bytes;
  char* sdu = smem + 3 * tile_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * tile_bytes);
  const int nlev = a.nlev;
  const size_t col0 = (size_t)blockIdx.x * kCols;
  const int nvalid = (int)((a.ncols - col0) < (size_t)kCols ? (a.ncols - col0) : (size_t)kCols);
  const int col_bytes = nlev * (int)sizeof(T);
  const size_t goff = col0 * (size_t)nlev;
  const size_t gvec0 = goff * sizeof(T) / 16;
  if (!WGEN) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(bar, (uint32_t)(3 * nvalid * col_bytes));
    }
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
      const int c = threadIdx.x;
      const size_t g = goff + (size_t)c * nlev;
      bulk_g2s(sdl + c * pitch, (const T*)a.dl + g, col_bytes, bar);
      bulk_g2s(sd + c * pitch, (const T*)a.d + g, col_bytes, bar);
      bulk_g2s(sdu + c * pitch, (const T*)a.du + g, col_bytes, bar);
    }
  }
  const int cpc = col_bytes / 16;  // vectors per column: a power of two <= 32 (a column is a lane group)
  const int cshift = 31 - __clz(cpc);
  const int total = nvalid * cpc;
  const T eps = (T)a.eps;
  constexpr bool round_each = sizeof(T) == 4;
  const bool has_v = MODE == MODE_JVP || MODE == MODE_PRE || a.v != nullptr;
  constexpr int UN = 2;  // rounds whose loads are all issued before the first use (as in the fused stage kernel)
  static_assert(KEEP % UN == 0, "rounds are processed in pairs");
#pragma unroll
  for (int r0 = 0; r0 < KEEP; r0 += UN) {
    V vx[UN], vv[UN], vu[UN], vf[UN], vt[NTS][UN];
    T acol[UN];
    bool ok[UN];
    int so[UN], kk[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int q = (int)threadIdx.x + (r0 + u) * kFusedThreads;
      ok[u] = q < total;
      const int col = q >> cshift;
      kk[u] = q & (cpc - 1);
      so[u] = col * pitch + kk[u] * 16;
      const size_t gi = gvec0 + (size_t)q;
      vx[u] = V{};
      vv[u] = V{};
      vu[u] = V{};
      vf[u] = V{};
      acol[u] = T(0);
      if (ok[u]) {
        vx[u] = reinterpret_cast<const V*>(a.x)[gi];
        if (has_v) vv[u] = reinterpret_cast<const V*>(a.v)[gi];
        if (MODE == MODE_JVP) {
          vu[u] = reinterpret_cast<const V*>(a.U0)[gi];
          vf[u] = reinterpret_cast<const V*>(a.f)[gi];
        }
#pragma unroll
        for (int k = 0; k < NTC; ++k) vt[k][u] = reinterpret_cast<const V*>(a.terms[k])[gi];
        acol[u] = ((const T*)a.acol)[col0 + col];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      T ex[VN], ev[VN], eu[VN], ef[VN], x2[VN], eo[VN];
      vec_unpack<T>(vx[u], ex);
      vec_unpack<T>(vv[u], ev);
      vec_unpack<T>(vu[u], eu);
      vec_unpack<T>(vf[u], ef);
      if (WGEN) {  // rows of W = dtγ J(x) - I at this thread's levels (column_wfact_kernel's expressions) -> shared tiles
        T xb = __shfl_up_sync(0xffffffffu, ex[VN - 1], 1, cpc);
        T xa = __shfl_down_sync(0xffffffffu, ex[0], 1, cpc);
        if (kk[u] == 0) xb = T(0);
        if (kk[u] == cpc - 1) xa = T(0);
        T rl[VN], rd[VN], ru[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          if (!NONLIN) {
            const double o = __dmul_rn(a.dtg, (double)acol[u]);
            rl[e] = (T)o;
            ru[e] = (T)o;
            rd[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
          } else {
            const T um = e > 0 ? ex[e > 0 ? e - 1 : 0] : xb;
            const T up = e + 1 < VN ? ex[e + 1 < VN ? e + 1 : e] : xa;
            const double lo = __dmul_rn(a.dtg, (double)kface(acol[u], um, ex[e]));
            const double hi = __dmul_rn(a.dtg, (double)kface(acol[u], ex[e], up));
            rl[e] = (T)lo;
            ru[e] = (T)hi;
            rd[e] = (T)__dsub_rn(-__dadd_rn(lo, hi), 1.0);
          }
        }
        if (ok[u]) {
          *reinterpret_cast<V*>(sdl + so[u]) = vec_pack<T>(rl);
          *reinterpret_cast<V*>(sd + so[u]) = vec_pack<T>(rd);
          *reinterpret_cast<V*>(sdu + so[u]) = vec_pack<T>(ru);
        }
      }
      if (MODE == MODE_PRE) {  // the right-hand side is the array itself
        if (ok[u]) *reinterpret_cast<V*>(sx + so[u]) = vv[u];
        continue;
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) x2[e] = MODE == MODE_JVP ? radd(ex[e], rmul(eps, ev[e])) : ex[e];  // K7
      T below = __shfl_up_sync(0xffffffffu, x2[VN - 1], 1, cpc);
      T above = __shfl_down_sync(0xffffffffu, x2[0], 1, cpc);
      if (kk[u] == 0) below = T(0);
      if (kk[u] == cpc - 1) above = T(0);
      T ti[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const T um = e > 0 ? x2[e > 0 ? e - 1 : 0] : below;
        const T up = e + 1 < VN ? x2[e + 1 < VN ? e + 1 : e] : above;
        ti[e] = NONLIN ? timp_nonlinear<T>(acol[u], um, x2[e], up) : timp_linear<T>(acol[u], um, x2[e], up);
      }
      if (MODE == MODE_JVP) {
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const T f2 = (T)__dsub_rn(__dadd_rn((double)eu[e], __dmul_rn(a.dtg, (double)ti[e])), (double)x2[e]);  // K3
          eo[e] = rdiv(rsub(f2, ef[e]), eps);                                                                   // K8
        }
      } else {
        T et[NTS][VN];
#pragma unroll
        for (int k = 0; k < NTC; ++k) vec_unpack<T>(vt[k][u], et[k]);
#pragma unroll
        for (int e = 0; e < VN; ++e) {  // axpy_chain_vec_kernel's evaluation: Float64 accumulator, Float32 states round per term
          double acc = op_add(0.0, op_mul(1.0, (double)ti[e]));
          if (round_each) acc = (double)(T)acc;
          if (has_v) {
            acc = op_add(acc, op_mul(1.0, (double)ev[e]));
            if (round_each) acc = (double)(T)acc;
          }
#pragma unroll
          for (int k = 0; k < NTC; ++k) {
            acc = op_add(acc, op_mul(a.coef[k], (double)et[k][e]));
            if (round_each) acc = (double)(T)acc;
          }
          eo[e] = (T)op_mul(acc, a.scale);
        }
      }
      if (ok[u]) *reinterpret_cast<V*>(sx + so[u]) = vec_pack<T>(eo);
    }
  }
  __syncthreads();
  if (threadIdx.x < kSolvers) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = warp * 16 + lane_column(lane);
    const bool is_up = lane_is_up(lane);
    const bool valid = col < nvalid;
    if (!WGEN) mbar_wait(bar, 0);
    StoredRows<T> rows{sdl, sd, sdu};
    babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, false>(rows, RhsFromTile<T>{}, sx, col * pitch, nlev, is_up, valid);
  }
  // x leaves as one bulk copy per column (padded tile -> contiguous global range), like K4
  fence_proxy_async_smem();  // generic-proxy writes of x -> visible to the async proxy
  __syncthreads();
  if ((int)threadIdx.x < nvalid) {
    const int c = threadIdx.x;
    bulk_s2g((T*)a.out + goff + (size_t)c * nlev, sx + c * pitch, col_bytes);
    bulk_commit();
    bulk_wait_read0();
  }
}

template <typename T>
int launch_rhs_solve(cts_op_column* op, RhsSolveArgs& a, int mode) {
  cts_ctx* ctx = op->ctx;
  constexpr int kCols = TileShape<T>::kCols;
  a.nlev = op->d.nlev;
  a.ncols = op->ncols;
  a.pitch = pitch_bytes(op->d.nlev, sizeof(T));
  a.acol = op->a_col;
  const size_t tiles = (op->ncols + kCols - 1) / kCols;
  const size_t smem = (size_t)4 * kCols * a.pitch + 16;
  auto launch = [&](auto k) -> int {
    CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)tiles, kFusedThreads, smem, ctx->stream>>>(a);
    return CTS_OK;
  };
  const bool nl = op->d.nonlinear;
  if (mode == MODE_JVP && !a.dl) {
    if (nl) CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_JVP, true, 0, true>));
    else CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_JVP, false, 0, true>));
  } else if (mode == MODE_PRE) {
    if (nl) CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_PRE, true, 0, true>));
    else CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_PRE, false, 0, true>));
  } else if (mode == MODE_JVP) {
    if (nl) CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_JVP, true, 0>));
    else CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_JVP, false, 0>));
  } else {
#define ROS_CASE(N)                                                              \
  case N:                                                                        \
    if (nl) CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_ROS, true, N>));      \
    else CTS_TRY(launch(column_rhs_solve_kernel<T, MODE_ROS, false, N>));        \
    break;
    switch (a.nt) {
      ROS_CASE(0) ROS_CASE(1) ROS_CASE(2) ROS_CASE(3)
      default:
        return CTS_EUNSUPPORTED;
    }
#undef ROS_CASE
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

}  // namespace

// a column must be a power-of-two number (<= 32) of 16-byte vectors and a tile must fit the threads' kKeep vectors
static bool rhs_solve_shape_ok(const cts_op_column* op);
bool cts_column_rhs_solve_fusable(const cts_op_column* op) { return !op->matrix_free && rhs_solve_shape_ok(op); }
// the matrix-free form of the Krylov kernels (W rows computed from x in the kernel): needs the flag, not a stored W
bool cts_column_krylov_matrix_free(const cts_op_column* op) { return op->matrix_free && rhs_solve_shape_ok(op); }
static bool rhs_solve_shape_ok(const cts_op_column* op) {
  const int es = op->ft == CTS_F64 ? 8 : 4;
  const int cb = op->d.nlev * es;
  if (op->d.nlev < 2 || (op->d.nlev % 2) || (cb % 16)) return false;
  const int cpc = cb / 16;
  const int cols = op->ft == CTS_F64 ? TileShape<double>::kCols : TileShape<float>::kCols;
  const int keep = op->ft == CTS_F64 ? KeepOf<double>::value : KeepOf<float>::value;
  return cpc <= 32 && (cpc & (cpc - 1)) == 0 && cpc * cols <= keep * kFusedThreads;
}

static bool rhs_solve_w_ok(const cts_op_column* op, const cts_tridiag* W) {
  return W && W->nlev == op->d.nlev && W->ncols == op->ncols && cts_aligned16(W->dl) && cts_aligned16(W->d) && cts_aligned16(W->du);
}

// q = W(x) \ b with the rows of W generated in the kernel (matrix-free operator)
int cts_column_precond_matrix_free(cts_op_column* op, void* q, const void* x, const void* b, double dtg) {
  if (!op || !q || !x || !b) return CTS_EINVAL;
  if (!cts_column_krylov_matrix_free(op) || !cts_aligned16(q) || !cts_aligned16(x) || !cts_aligned16(b)) return CTS_EUNSUPPORTED;
  RhsSolveArgs a{};
  a.out = q;
  a.x = x;
  a.v = b;
  a.dtg = dtg;
  CTS_COUNT_BYTES(op->ctx, 3, cts_op_column_ndof(op), cts_sizeof(op->ft));  // x, b in; q out
  return op->ft == CTS_F64 ? launch_rhs_solve<double>(op, a, MODE_PRE) : launch_rhs_solve<float>(op, a, MODE_PRE);
}

// W == nullptr: matrix-free operator, the rows of W are generated in the kernel from x
int cts_column_jvp_precond(cts_op_column* op, void* q, const void* x, const void* v, double eps, const void* U0, const void* f,
                           double dtg, const cts_tridiag* W) {
  if (!op || !q || !x || !v || !U0 || !f) return CTS_EINVAL;
  if (!cts_aligned16(q) || !cts_aligned16(x) || !cts_aligned16(v) || !cts_aligned16(U0) || !cts_aligned16(f)) return CTS_EUNSUPPORTED;
  if (!W) {
    if (!cts_column_krylov_matrix_free(op)) return CTS_EUNSUPPORTED;
    RhsSolveArgs a{};
    a.out = q;
    a.x = x;
    a.v = v;
    a.U0 = U0;
    a.f = f;
    a.eps = eps;
    a.dtg = dtg;
    CTS_COUNT_BYTES(op->ctx, 5, cts_op_column_ndof(op), cts_sizeof(op->ft));  // x, v, U0, f in; q out
    return op->ft == CTS_F64 ? launch_rhs_solve<double>(op, a, MODE_JVP) : launch_rhs_solve<float>(op, a, MODE_JVP);
  }
  if (!cts_column_rhs_solve_fusable(op) || !rhs_solve_w_ok(op, W))
    return CTS_EUNSUPPORTED;
  RhsSolveArgs a{};
  a.out = q;
  a.x = x;
  a.v = v;
  a.U0 = U0;
  a.f = f;
  a.eps = eps;
  a.dtg = dtg;
  a.dl = W->dl;
  a.d = W->d;
  a.du = W->du;
  CTS_COUNT_BYTES(op->ctx, 8, cts_op_column_ndof(op), cts_sizeof(op->ft));  // x, v, U0, f, dl, d, du in; q out
  return op->ft == CTS_F64 ? launch_rhs_solve<double>(op, a, MODE_JVP) : launch_rhs_solve<float>(op, a, MODE_JVP);
}

int cts_column_rosenbrock_stage(cts_op_column* op, void* k_out, const void* U, const void* fU_exp, int nt, const double* coef,
                                const void* const* terms, double dtg, const cts_tridiag* W) {
  if (!op || !k_out || !U || nt < 0) return CTS_EINVAL;
  if (nt > kRhsMaxTerms || !cts_column_rhs_solve_fusable(op) || !rhs_solve_w_ok(op, W) || !cts_aligned16(k_out) || !cts_aligned16(U) ||
      (fU_exp && !cts_aligned16(fU_exp)))
    return CTS_EUNSUPPORTED;
  RhsSolveArgs a{};
  a.out = k_out;
  a.x = U;
  a.v = fU_exp;
  a.nt = nt;
  for (int k = 0; k < nt; ++k) {
    if (!cts_aligned16(terms[k])) return CTS_EUNSUPPORTED;
    a.terms[k] = terms[k];
    a.coef[k] = coef[k];
  }
  a.scale = -dtg;
  a.dtg = dtg;
  a.dl = W->dl;
  a.d = W->d;
  a.du = W->du;
  CTS_COUNT_BYTES(op->ctx, 1 + (fU_exp ? 1 : 0) + nt + 3 + 1, cts_op_column_ndof(op), cts_sizeof(op->ft));
  return op->ft == CTS_F64 ? launch_rhs_solve<double>(op, a, MODE_ROS) : launch_rhs_solve<float>(op, a, MODE_ROS);
}

namespace {

template <typename T, typename C>
int dispatch_fused(cts_op_column* op, const cts_fused_stage_args* s) {
  switch (s->n_grp1 + s->n_grp2) {
#define CASE(N) \
  case N:       \
    return launch_fused<T, C, N>(op, s);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
        CASE(13) CASE(14)
#undef CASE
    default:
      return CTS_EUNSUPPORTED;
  }
}

}  // namespace

bool cts_column_fusable(const cts_op_column* op) {
  const int es = op->ft == CTS_F64 ? 8 : 4;
  size_t smem = (size_t)5 * kTileCols * pitch_bytes(op->d.nlev, es) + 16;
  return !op->d.nonlinear && op->d.nlev >= 2 && (op->d.nlev % 2) == 0 && ((op->d.nlev * es) % 16) == 0 &&
         smem <= 200 * 1024;
}

int cts_column_fused_stage(cts_op_column* op, const cts_fused_stage_args* s) {
  if (!op || !s) return CTS_EINVAL;
  if (!cts_column_fusable(op)) return CTS_EUNSUPPORTED;
  bool aligned = cts_aligned16(s->base) && cts_aligned16(s->U) && (!s->T_imp || cts_aligned16(s->T_imp)) &&
                 (!s->temp || cts_aligned16(s->temp));
  for (int k = 0; k < s->n_grp1 + s->n_grp2; ++k) aligned = aligned && cts_aligned16(s->terms[k]);
  if (!op->matrix_free && s->W)
    aligned = aligned && cts_aligned16(s->W->dl) && cts_aligned16(s->W->d) && cts_aligned16(s->W->du);
  if (!aligned) return CTS_EUNSUPPORTED;
  if (op->ft == CTS_F64) return dispatch_fused<double, double>(op, s);
  if (s->compute_native) return dispatch_fused<float, float>(op, s);
  return dispatch_fused<float, double>(op, s);
}

template <typename T, typename C, int NT>
int launch_final(cts_op_column* op, const cts_final_update_args* s, size_t jr_begin, size_t jr_end) {
  cts_ctx* ctx = op->ctx;
  if (jr_begin >= jr_end) return CTS_OK;
  const cts_column_desc& d = op->d;
  FinalArgs<NT> a{};
  a.ax.base = s->u;
  a.ax.out = s->u;
  a.ax.n1 = s->n_grp1;
  for (int k = 0; k < NT; ++k) {
    a.ax.terms[k] = s->terms[k];
    a.ax.coef[k] = s->coef[k];
  }
  a.U = s->U_last;
  a.S = d.S_lev;
  a.coef_last = s->coef_last;
  a.g = 1.0 + 0.5 * sin(2.0 * M_PI * s->t_exp_last);
  a.nu = d.nu;
  a.nlev = (unsigned)d.nlev;
  a.nx = (unsigned)d.nx;
  a.row_elems = (unsigned)(d.nx * (size_t)d.nlev);
  a.ny_local = d.ny_local;
  a.halo = d.halo;
  a.has_src = d.S_lev != nullptr;
  a.jr_begin = jr_begin;
  a.jr_end = jr_end;
  CTS_COUNT_BYTES(ctx, 3 + NT, (jr_end - jr_begin) * a.row_elems, sizeof(T));  // u, U_s, terms in; u out
  constexpr int VN = Vec<T>::N;
  const unsigned xblocks = (unsigned)((a.row_elems / VN + 255) / 256);
  a.strip_rows = march_strip_rows(column_final_update_kernel<T, C, NT>, ctx, jr_end - jr_begin, xblocks);
  dim3 grid(xblocks, (unsigned)((jr_end - jr_begin + a.strip_rows - 1) / a.strip_rows));
  column_final_update_kernel<T, C, NT><<<grid, 256, 0, ctx->stream>>>(a);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

template <typename T, typename C>
int dispatch_final(cts_op_column* op, const cts_final_update_args* s, size_t jr_begin, size_t jr_end) {
  switch (s->n_grp1 + s->n_grp2) {
#define CASE(N) \
  case N:       \
    return launch_final<T, C, N>(op, s, jr_begin, jr_end);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    default:
      return CTS_EUNSUPPORTED;
  }
}

bool cts_column_final_fusable(const cts_op_column* op) {
  const cts_column_desc& d = op->d;
  const int vn = op->ft == CTS_F64 ? 2 : 4;
  const size_t row_elems = d.nx * (size_t)d.nlev;
  return d.nu != 0.0 && (d.nlev % vn) == 0 && row_elems < (1ull << 31) && (d.ny_local + 7) / 8 <= 65535 &&
         (!d.S_lev || cts_aligned16(d.S_lev));
}

static int final_update_rows(cts_op_column* op, const cts_final_update_args* s, size_t jr_begin, size_t jr_end) {
  if (op->ft == CTS_F64) return dispatch_final<double, double>(op, s, jr_begin, jr_end);
  if (s->compute_native) return dispatch_final<float, float>(op, s, jr_begin, jr_end);
  return dispatch_final<float, double>(op, s, jr_begin, jr_end);
}

bool cts_column_final_update_supported(const cts_op_column* op, const cts_final_update_args* s) {
  if (!op || !s || !s->u || !s->U_last || s->u == s->U_last) return false;
  if (!cts_column_final_fusable(op) || s->n_grp1 + s->n_grp2 > 10) return false;
  bool aligned = cts_aligned16(s->u) && cts_aligned16(s->U_last);
  for (int k = 0; k < s->n_grp1 + s->n_grp2; ++k) aligned = aligned && cts_aligned16(s->terms[k]);
  return aligned;
}

int cts_column_fused_final_update(cts_op_column* op, const cts_final_update_args* s) {
  if (!op || !s || !s->u || !s->U_last || s->u == s->U_last) return CTS_EINVAL;
  if (!cts_column_final_update_supported(op, s)) return CTS_EUNSUPPORTED;
  return final_update_rows(op, s, 0, op->d.ny_local);
}

bool cts_column_residual_fusable(const cts_op_column* op) {
  const int vn = op->ft == CTS_F64 ? 2 : 4;
  return op->d.nlev >= vn && (op->d.nlev % vn) == 0;
}

int cts_column_fused_residual(cts_op_column* op, void* out, const void* x, const void* dx, double eps, const void* U0,
                              const void* f, double dtg) {
  if (!op || !out || !x || !U0 || (dx && !f)) return CTS_EINVAL;
  if (!cts_column_residual_fusable(op) || !cts_aligned16(out) || !cts_aligned16(x) || !cts_aligned16(U0) ||
      (dx && (!cts_aligned16(dx) || !cts_aligned16(f))))
    return CTS_EUNSUPPORTED;
  cts_ctx* ctx = op->ctx;
  const size_t n = cts_op_column_ndof(op);
  const int nl = op->d.nlev;
#define LAUNCH(T, NONLIN, JVP)                                                                                        \
  column_residual_kernel<T, NONLIN, JVP><<<(unsigned)((n / Vec<T>::N + 255) / 256), 256, 0, ctx->stream>>>(            \
      (T*)out, (const T*)x, (const T*)dx, (const T*)U0, (const T*)f, (const T*)op->a_col, (T)eps, dtg, nl, n / Vec<T>::N)
  const bool nonlin = op->d.nonlinear, jvp = dx != nullptr;
  CTS_COUNT_BYTES(ctx, jvp ? 5 : 3, n, cts_sizeof(op->ft));
  if (op->ft == CTS_F64) {
    if (nonlin) { if (jvp) LAUNCH(double, true, true); else LAUNCH(double, true, false); }
    else        { if (jvp) LAUNCH(double, false, true); else LAUNCH(double, false, false); }
  } else {
    if (nonlin) { if (jvp) LAUNCH(float, true, true); else LAUNCH(float, true, false); }
    else        { if (jvp) LAUNCH(float, false, true); else LAUNCH(float, false, false); }
  }
#undef LAUNCH
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

bool cts_op_column_is_linear(const cts_op_column* op) { return !op->d.nonlinear; }
bool cts_op_column_has_halo(const cts_op_column* op) { return op->d.halo != 0; }
bool cts_op_column_matrix_free(const cts_op_column* op) { return op->matrix_free; }
cts_ctx* cts_op_column_ctx(const cts_op_column* op) { return op->ctx; }
cts_dtype cts_op_column_dtype(const cts_op_column* op) { return op->ft; }

extern "C" {

int cts_op_column_create(cts_ctx* ctx, cts_dtype ft, const cts_column_desc* desc, cts_op_column** out) {
  CTS_CHECK_CTX(ctx);
  if (!out || !desc || desc->nlev < 1 || desc->nx < 1 || desc->ny_local < 1 || !desc->K_col || desc->dz == 0.0 ||
      (desc->halo != 0 && desc->halo != 1))
    return cts_fail(ctx, CTS_EINVAL, "cts_op_column_create: bad descriptor");
  if (desc->halo == 0 && ctx->nranks > 1 && desc->nu != 0.0)
    return cts_fail(ctx, CTS_EINVAL, "horizontal coupling across ranks needs halo = 1");
  auto* op = new cts_op_column();
  op->ctx = ctx;
  op->ft = ft;
  op->d = *desc;
  op->rows = desc->ny_local + 2 * (size_t)desc->halo;
  op->ncols = op->rows * desc->nx;
  op->matrix_free = false;
  op->a_col = nullptr;
  int r = cts_alloc(ctx, op->ncols * cts_sizeof(ft), &op->a_col);
  if (r != CTS_OK) {
    delete op;
    return r;
  }
  unsigned blocks = (unsigned)((op->ncols + 255) / 256);
  if (ft == CTS_F64)
    column_acol_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double*)op->a_col, (const double*)desc->K_col, desc->dz, op->ncols);
  else
    column_acol_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float*)op->a_col, (const float*)desc->K_col, (float)desc->dz, op->ncols);
  CTS_LAUNCH_CHECK(ctx);
  *out = op;
  return CTS_OK;
}

int cts_op_column_destroy(cts_op_column* op) {
  if (op && op->a_col) cts_free(op->ctx, op->a_col);
  delete op;
  return CTS_OK;
}

size_t cts_op_column_ndof(const cts_op_column* op) { return op ? op->ncols * (size_t)op->d.nlev : 0; }

int cts_op_column_set_matrix_free(cts_op_column* op, int enabled) {
  if (!op) return CTS_EINVAL;
  // linear operator: the fused stage kernels regenerate W from K_col; nonlinear operator: the Newton-Krylov tile kernels
  // generate the frozen-coefficient W(x) from the Newton iterate (column_rhs_solve_kernel<.., WGEN>); every other path of a
  // matrix-free operator still goes through Wfact / ldiv! on the stored W
  op->matrix_free = enabled != 0;
  return CTS_OK;
}

int cts_op_column_make_W(cts_op_column* op, cts_tridiag** W) {
  if (!op || !W) return CTS_EINVAL;
  size_t bytes = cts_op_column_ndof(op) * cts_sizeof(op->ft);
  auto* w = new cts_tridiag();
  w->nlev = op->d.nlev;
  w->ncols = op->ncols;
  int r = cts_alloc(op->ctx, bytes, &w->dl);
  if (!r) r = cts_alloc(op->ctx, bytes, &w->d);
  if (!r) r = cts_alloc(op->ctx, bytes, &w->du);
  if (r) {
    cts_op_column_free_W(op, w);
    return r;
  }
  *W = w;
  return CTS_OK;
}

int cts_op_column_free_W(cts_op_column* op, cts_tridiag* W) {
  if (!op || !W) return CTS_EINVAL;
  cts_free(op->ctx, W->dl);
  cts_free(op->ctx, W->d);
  cts_free(op->ctx, W->du);
  delete W;
  return CTS_OK;
}

int cts_op_column_T_imp(void* vop, void* du, const void* u, double t) {
  auto* op = static_cast<cts_op_column*>(vop);
  if (!op || !du || !u || du == u) return CTS_EINVAL;
  cts_ctx* ctx = op->ctx;
  size_t n = cts_op_column_ndof(op);
  size_t blocks = (n + 255) / 256;
  const int nl = op->d.nlev;
  CTS_COUNT_BYTES(ctx, 2, n, cts_sizeof(op->ft));
  if (op->ft == CTS_F64) {
    if (op->d.nonlinear)
      column_timp_kernel<double, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)du, (const double*)u, (const double*)op->a_col, nl, n);
    else
      column_timp_kernel<double, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)du, (const double*)u, (const double*)op->a_col, nl, n);
  } else {
    if (op->d.nonlinear)
      column_timp_kernel<float, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)du, (const float*)u, (const float*)op->a_col, nl, n);
    else
      column_timp_kernel<float, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)du, (const float*)u, (const float*)op->a_col, nl, n);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

int cts_op_column_Wfact(void* vop, void* vW, const void* u, double dtgamma, double t) {
  auto* op = static_cast<cts_op_column*>(vop);
  auto* W = static_cast<cts_tridiag*>(vW);
  if (!op || !W || !u) return CTS_EINVAL;
  cts_ctx* ctx = op->ctx;
  size_t n = cts_op_column_ndof(op);
  size_t blocks = (n + 255) / 256;
  const int nl = op->d.nlev;
  CTS_COUNT_BYTES(ctx, op->d.nonlinear ? 4 : 3, n, cts_sizeof(op->ft));
  if (op->ft == CTS_F64) {
    if (op->d.nonlinear)
      column_wfact_kernel<double, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)W->dl, (double*)W->d, (double*)W->du, (const double*)u, (const double*)op->a_col, dtgamma, nl, n);
    else
      column_wfact_kernel<double, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)W->dl, (double*)W->d, (double*)W->du, (const double*)u, (const double*)op->a_col, dtgamma, nl, n);
  } else {
    if (op->d.nonlinear)
      column_wfact_kernel<float, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)W->dl, (float*)W->d, (float*)W->du, (const float*)u, (const float*)op->a_col, dtgamma, nl, n);
    else
      column_wfact_kernel<float, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)W->dl, (float*)W->d, (float*)W->du, (const float*)u, (const float*)op->a_col, dtgamma, nl, n);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

int cts_op_column_ldiv(void* vop, void* x, const void* vW, const void* b) {
  auto* op = static_cast<cts_op_column*>(vop);
  if (!op || !vW) return CTS_EINVAL;
  return cts_tridiag_solve_batched(op->ctx, op->ft, static_cast<const cts_tridiag*>(vW), b, x);
}

int cts_op_column_jac_mul(void* vop, void* y, const void* vW, const void* x) {
  auto* op = static_cast<cts_op_column*>(vop);
  if (!op || !vW) return CTS_EINVAL;
  return cts_tridiag_matvec(op->ctx, op->ft, static_cast<const cts_tridiag*>(vW), x, y);
}

// T_exp! on the owned rows [jr_begin, jr_end)
static int column_T_exp_rows(cts_op_column* op, void* du_exp, const void* u, double t, size_t jr_begin, size_t jr_end) {
  if (!op || !du_exp || !u || du_exp == u) return CTS_EINVAL;
  if (jr_begin >= jr_end) return CTS_OK;
  cts_ctx* ctx = op->ctx;
  const cts_column_desc& d = op->d;
  // time factor of the source, evaluated once on the host in Float64 (docs/src/tutorials/imex_diffusion.md:65-69)
  double g = 1.0 + 0.5 * sin(2.0 * M_PI * t);
  int has_src = d.S_lev != nullptr;
  const int vn = op->ft == CTS_F64 ? 2 : 4;
  const bool vec = (d.nlev % vn) == 0 && cts_aligned16(du_exp) && cts_aligned16(u) && (!has_src || cts_aligned16(d.S_lev));
  const size_t row_elems = d.nx * (size_t)d.nlev;
  const size_t nrows = jr_end - jr_begin;
  CTS_COUNT_BYTES(ctx, 2, nrows * row_elems, cts_sizeof(op->ft));
  if (vec && row_elems < (1ull << 31) && d.nu != 0.0 && (nrows + 3) / 4 <= 65535) {
    // strips of 8 rows (25 % of the rows are fetched twice, from L2): measured best of 8/16/32/64/128 on B200 for a full slab;
    // small launches pick the strip height that fills the waves of CTAs (march_strip_rows)
    const unsigned xblocks = (unsigned)((row_elems / vn + 255) / 256);
    const int ROWS = op->ft == CTS_F64 ? march_strip_rows(column_texp_march_kernel<double, 2>, ctx, nrows, xblocks)
                                       : march_strip_rows(column_texp_march_kernel<float, 4>, ctx, nrows, xblocks);
    dim3 grid(xblocks, (unsigned)((nrows + ROWS - 1) / ROWS));
    if (op->ft == CTS_F64)
      column_texp_march_kernel<double, 2><<<grid, 256, 0, ctx->stream>>>((double*)du_exp, (const double*)u, (const double*)d.S_lev, g, d.nu, (unsigned)d.nlev, (unsigned)d.nx, (unsigned)row_elems, d.ny_local, d.halo, has_src, jr_begin, jr_end, ROWS);
    else
      column_texp_march_kernel<float, 4><<<grid, 256, 0, ctx->stream>>>((float*)du_exp, (const float*)u, (const float*)d.S_lev, (float)g, (float)d.nu, (unsigned)d.nlev, (unsigned)d.nx, (unsigned)row_elems, d.ny_local, d.halo, has_src, jr_begin, jr_end, ROWS);
    CTS_LAUNCH_CHECK(ctx);
    return CTS_OK;
  }
  if (jr_begin != 0 || jr_end != d.ny_local) return CTS_EUNSUPPORTED;  // the flat kernels take all owned rows
  size_t n_owned = d.ny_local * d.nx * (size_t)d.nlev;
  size_t blocks = (n_owned + 255) / 256;
  if (vec) {
    blocks = (n_owned / vn + 255) / 256;
    if (op->ft == CTS_F64)
      column_texp_kernel<double, 2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)du_exp, (const double*)u, (const double*)d.S_lev, g, d.nu, d.nlev, d.nx, d.ny_local, d.halo, has_src);
    else
      column_texp_kernel<float, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)du_exp, (const float*)u, (const float*)d.S_lev, (float)g, (float)d.nu, d.nlev, d.nx, d.ny_local, d.halo, has_src);
  } else {
    if (op->ft == CTS_F64)
      column_texp_kernel<double, 1><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)du_exp, (const double*)u, (const double*)d.S_lev, g, d.nu, d.nlev, d.nx, d.ny_local, d.halo, has_src);
    else
      column_texp_kernel<float, 1><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)du_exp, (const float*)u, (const float*)d.S_lev, (float)g, (float)d.nu, d.nlev, d.nx, d.ny_local, d.halo, has_src);
  }
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

int cts_op_column_T_exp(void* vop, void* du_exp, void* du_lim, const void* u, double t) {
  auto* op = static_cast<cts_op_column*>(vop);
  if (!op) return CTS_EINVAL;
  return column_T_exp_rows(op, du_exp, u, t, 0, op->d.ny_local);
}

// dss!(u): refresh the ghost rows from their owners, enqueued on `stream` (periodic local copies on one rank, the
// NVLink peer-memory halo kernel or NCCL send/recv otherwise)
static int column_halo_on(cts_op_column* op, cudaStream_t stream, void* u) {
  cts_ctx* ctx = op->ctx;
  const size_t row = op->d.nx * (size_t)op->d.nlev;
  const size_t es = cts_sizeof(op->ft);
  char* b = static_cast<char*>(u);
  char* ghost_lo = b;
  char* first = b + row * es;
  char* last = b + op->d.ny_local * row * es;
  char* ghost_hi = b + (op->d.ny_local + 1) * row * es;
  if (ctx->nranks <= 1) {  // periodic in the slow horizontal axis: the neighbour is this rank
    CTS_CUDA(ctx, cudaMemcpyAsync(ghost_lo, last, row * es, cudaMemcpyDeviceToDevice, stream));
    CTS_CUDA(ctx, cudaMemcpyAsync(ghost_hi, first, row * es, cudaMemcpyDeviceToDevice, stream));
    return CTS_OK;
  }
  if (ctx->peer_mode == 0) CTS_TRY(cts_comm_enable_peer_halo(ctx, row * es));  // first dss!: collective set-up
  return cts_halo_exchange_on(ctx, stream, op->ft, row, first, last, ghost_lo, ghost_hi);
}

int cts_op_column_dss(void* vop, void* u, double t) {
  auto* op = static_cast<cts_op_column*>(vop);
  if (!op || !u) return CTS_EINVAL;
  if (!op->d.halo) return CTS_OK;  // nothing is duplicated
  return column_halo_on(op, op->ctx->stream, u);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Halo / compute overlap.  The boundary work of a stencil pass runs on a second (high-priority) stream together with
// the halo exchange it depends on, the interior rows on the compute stream; the two meet again before the call returns,
// so callers keep plain stream-ordered semantics.  Captured into a CUDA graph this is a fork/join.
// ---------------------------------------------------------------------------------------------
constexpr size_t kEdgeRows = 8;  // one strip of the marching kernels

bool cts_column_overlap_ok(const cts_op_column* op) {
  const cts_column_desc& d = op->d;
  const int vn = op->ft == CTS_F64 ? 2 : 4;
  static const bool off = getenv("CTS_NO_OVERLAP") != nullptr;  // developer switch
  return !off && d.halo == 1 && d.nu != 0.0 && (d.nlev % vn) == 0 && d.ny_local >= 4 * kEdgeRows &&
         d.nx * (size_t)d.nlev < (1ull << 31);
}

int cts_ctx_ensure_comm_stream(cts_ctx* ctx) {
  if (ctx->comm_stream) return CTS_OK;
  int lo = 0, hi = 0;
  CTS_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = greatest priority
  CTS_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
  CTS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
  CTS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
  return CTS_OK;
}

namespace {
struct StreamSwap {  // launches issued through the context go to `s` for the lifetime of the object
  cts_ctx* ctx;
  cudaStream_t saved;
  StreamSwap(cts_ctx* c, cudaStream_t s) : ctx(c), saved(c->stream) { c->stream = s; }
  ~StreamSwap() { ctx->stream = saved; }
};
}  // namespace

int cts_column_dss_T_exp_overlapped(cts_op_column* op, void* du_exp, void* U, double t_dss, double t_exp) {
  if (!op || !du_exp || !U) return CTS_EINVAL;
  const cts_column_desc& d = op->d;
  const int vn = op->ft == CTS_F64 ? 2 : 4;
  const bool march = cts_aligned16(du_exp) && cts_aligned16(U) && (!d.S_lev || cts_aligned16(d.S_lev)) && (d.nlev % vn) == 0;
  if (!cts_column_overlap_ok(op) || !march) {
    int r = cts_op_column_dss(op, U, t_dss);
    return r != CTS_OK ? r : cts_op_column_T_exp(op, du_exp, nullptr, U, t_exp);
  }
  cts_ctx* ctx = op->ctx;
  CTS_TRY(cts_ctx_ensure_comm_stream(ctx));
  if (ctx->nranks > 1 && ctx->peer_mode == 0)  // collective set-up must not happen between the fork and the join
    CTS_TRY(cts_comm_enable_peer_halo(ctx, d.nx * (size_t)d.nlev * cts_sizeof(op->ft)));
  // fork: the stage value is complete on the compute stream
  CTS_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->stream));
  CTS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  {
    StreamSwap sw(ctx, ctx->comm_stream);
    CTS_TRY(column_halo_on(op, ctx->comm_stream, U));
    // the two strips whose stencil reads a ghost row
    CTS_TRY(column_T_exp_rows(op, du_exp, U, t_exp, 0, kEdgeRows));
    CTS_TRY(column_T_exp_rows(op, du_exp, U, t_exp, d.ny_local - kEdgeRows, d.ny_local));
  }
  CTS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  // rows whose 5-point stencil stays inside the owned rows: no ghost row is read, nothing the exchange writes is touched
  CTS_TRY(column_T_exp_rows(op, du_exp, U, t_exp, kEdgeRows, d.ny_local - kEdgeRows));
  CTS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));  // join
  return CTS_OK;
}

// dss!(U_s); final update (with the last T_exp! folded in); dss!(u) -- the tail of an IMEX-ARK step (imex_ark.jl:234,79-101,
// 98).  Second stream: halo(U_s) -> boundary strips of the final update -> halo(u); compute stream: interior rows.
int cts_column_final_update_overlapped(cts_op_column* op, const cts_final_update_args* s) {
  if (!op || !s) return CTS_EINVAL;
  if (!cts_column_final_update_supported(op, s)) return CTS_EUNSUPPORTED;
  const cts_column_desc& d = op->d;
  void* U_last = const_cast<void*>(s->U_last);
  if (!cts_column_overlap_ok(op)) {
    CTS_TRY(cts_op_column_dss(op, U_last, s->t_exp_last));
    CTS_TRY(final_update_rows(op, s, 0, d.ny_local));
    return cts_op_column_dss(op, s->u, s->t_exp_last);
  }
  cts_ctx* ctx = op->ctx;
  CTS_TRY(cts_ctx_ensure_comm_stream(ctx));
  if (ctx->nranks > 1 && ctx->peer_mode == 0)
    CTS_TRY(cts_comm_enable_peer_halo(ctx, d.nx * (size_t)d.nlev * cts_sizeof(op->ft)));
  CTS_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->stream));
  CTS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  {
    StreamSwap sw(ctx, ctx->comm_stream);
    CTS_TRY(column_halo_on(op, ctx->comm_stream, U_last));
    CTS_TRY(final_update_rows(op, s, 0, kEdgeRows));
    CTS_TRY(final_update_rows(op, s, d.ny_local - kEdgeRows, d.ny_local));
    CTS_TRY(column_halo_on(op, ctx->comm_stream, s->u));
  }
  CTS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  CTS_TRY(final_update_rows(op, s, kEdgeRows, d.ny_local - kEdgeRows));
  CTS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  return CTS_OK;
}
