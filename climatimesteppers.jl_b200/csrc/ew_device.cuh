// Device helpers shared by elementwise.cu and ops.cu: 16-byte vectors, rounding-exact scalar ops and the
// K1 combine  out = (c0*base + G1) + G2  (fused_increment.jl:220-236 evaluation order).
#pragma once
#include <cuda_runtime.h>

namespace cts_ew {

template <typename T>
struct Vec;  // 16-byte vector of T
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

template <typename T>
__device__ __forceinline__ void vec_unpack(const typename Vec<T>::type& v, T (&e)[Vec<T>::N]);
template <>
__device__ __forceinline__ void vec_unpack<double>(const double2& v, double (&e)[2]) {
  e[0] = v.x;
  e[1] = v.y;
}
template <>
__device__ __forceinline__ void vec_unpack<float>(const float4& v, float (&e)[4]) {
  e[0] = v.x;
  e[1] = v.y;
  e[2] = v.z;
  e[3] = v.w;
}
template <typename T>
__device__ __forceinline__ typename Vec<T>::type vec_pack(const T (&e)[Vec<T>::N]);
template <>
__device__ __forceinline__ double2 vec_pack<double>(const double (&e)[2]) {
  return make_double2(e[0], e[1]);
}
template <>
__device__ __forceinline__ float4 vec_pack<float>(const float (&e)[4]) {
  return make_float4(e[0], e[1], e[2], e[3]);
}

// rounding-exact scalar ops (no FMA contraction) in the compute type C
__device__ __forceinline__ double op_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double op_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double op_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double op_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float op_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float op_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float op_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float op_div(float a, float b) { return __fdiv_rn(a, b); }

// ---------------------------------------------------------------------------------------------
// K1: out = (c0*base + G1) + G2
// ---------------------------------------------------------------------------------------------
template <int NT>
struct AxpyArgs {
  const void* base;
  void* out;
  void* out2;
  const void* terms[NT > 0 ? NT : 1];
  double coef[NT > 0 ? NT : 1];
  double c0;
  int n1;       // terms in group 1 (the rest are group 2)
  int scale_base;  // c0 != 1
};

template <typename T, typename C, int NT>
__device__ __forceinline__ T axpy_combine(T b, const T (&x)[NT > 0 ? NT : 1], const AxpyArgs<NT>& a) {
  C r = (C)b;
  if (a.scale_base) r = op_mul((C)a.c0, r);
  if (NT > 0) {
    C g1 = 0, g2 = 0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      C p = op_mul((C)a.coef[k], (C)x[k]);
      if (k < a.n1)
        g1 = (k == 0) ? p : op_add(g1, p);
      else
        g2 = (k == a.n1) ? p : op_add(g2, p);
    }
    if (a.n1 > 0) r = op_add(r, g1);
    if (a.n1 < NT) r = op_add(r, g2);
  }
  return (T)r;
}

// Same combine with the group split known at compile time: one multiply and one add per term, no selects.  (With a
// run-time n1 the compiler evaluates both "add to g1" and "add to g2" for every term and selects: 4 FSEL + 1 extra DADD per
// term and element, which matters in the issue-bound fused column stage.)
template <typename T, typename C, int NT, int N1>
__device__ __forceinline__ T axpy_combine_n1(T b, const T (&x)[NT > 0 ? NT : 1], const AxpyArgs<NT>& a) {
  static_assert(N1 >= 0 && N1 <= NT, "group-1 count");
  C r = (C)b;
  if (a.scale_base) r = op_mul((C)a.c0, r);
  if (NT > 0) {
    C g1 = 0, g2 = 0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      C p = op_mul((C)a.coef[k], (C)x[k]);
      if (k < N1)
        g1 = (k == 0) ? p : op_add(g1, p);
      else
        g2 = (k == N1) ? p : op_add(g2, p);
    }
    if (N1 > 0) r = op_add(r, g1);
    if (N1 < NT) r = op_add(r, g2);
  }
  return (T)r;
}

}  // namespace cts_ew
