// Definitions shared by the fused column stage kernels of ops.cu and fused_cluster.cu.
#pragma once
#include "ew_device.cuh"
#include "tile_device.cuh"

namespace cts_fused {

template <int NT>
struct FusedArgs {
  cts_ew::AxpyArgs<NT> ax;   // base/terms/coef (out/out2 unused)
  void* U;
  void* Timp;
  void* temp;
  const void* Kcol;
  const void* dl;
  const void* d;
  const void* du;
  double dtg;
  double dz;
  int nlev;
  size_t ncols;
  int pitch;
  int emit_w;      // matrix-free launch that also writes W = dtγJ - I into dl/d/du (replaces a separate Wfact pass)
  long long* dbg;  // optional phase timeline / wait profile (CTS_FUSED_TIMELINE=1)
};

// N1SPEC: how a kernel learns the split of the NT operands into the explicit group (first n1) and the implicit group:
// 0 = from the arguments at run time, 1 = all in group 1 (n1 == NT), 2 = n1 == ceil(NT/2) (stage i of an IMEX-ARK tableau
// with dense lower triangles: i-1 explicit and i-2 implicit tendencies).  A compile-time split needs no selects.
template <int NT, int N1SPEC>
struct GroupSplit {
  static constexpr int value = N1SPEC == 1 ? NT : (N1SPEC == 2 ? (NT + 1) / 2 : -1);
};

// linear operator row: a*((um - 2 uc) + up)
template <typename T>
__device__ __forceinline__ T timp_linear(T a, T um, T uc, T up) {
  return cts_tile::rmul(a, cts_tile::radd(cts_tile::rsub(um, cts_tile::rmul(T(2), uc)), up));
}

}  // namespace cts_fused

// Cluster variant of the fused stage (fused_cluster.cu): loader CTA + solver CTA on two SMs.  `args` is a FusedArgs<nt>;
// returns CTS_EUNSUPPORTED when no instantiation exists for (nt, n1spec, matrix_free) or clusters cannot be scheduled.
struct cts_ctx;
int cts_fused_cluster_f64(cts_ctx* ctx, int nt, int n1spec, bool matrix_free, const void* args, bool timeline);
