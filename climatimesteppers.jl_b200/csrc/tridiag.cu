// K4: batched per-column tridiagonal solve  W dx = f  (`ldiv!(Δx, W, f)`, newtons_method.jl:633) and the
// matching `mul!`.
//
// Memory plan (the kernel is HBM bound: 3 diagonals + rhs in, x out = 5w bytes per DOF):
//   * a CTA owns a tile of 32 consecutive columns = one contiguous 32*nlev*w byte range per array;
//   * the four input arrays are staged into shared memory by 1-D bulk TMA (cp.async.bulk, one 512-byte
//     copy per column so the columns land on a conflict-free pitch) completing on one mbarrier;
//   * a lane pair per column runs the two-sided Thomas elimination out of shared memory with
//     LDS.128/STS.128 (tile_device.cuh), halving the dependent-division chain of plain Thomas;
//   * x overwrites the rhs tile and leaves through bulk TMA stores.
//   * 66 KB of shared memory per CTA -> 3 CTAs (192 threads) resident per SM; the block scheduler
//     overlaps one tile's loads with another tile's elimination.
#include <cstdlib>

#include "cts_internal.h"
#include "tile_device.cuh"

using namespace cts_tile;

namespace {

// LOADER: 0 = cooperative LDG/STS, 1 = 1-D bulk TMA (one 512-byte copy per column), 2 = cp.async (LDGSTS)
template <typename T, int LOADER>
__global__ void __launch_bounds__(kTileThreads) tridiag_tile_kernel(const T* __restrict__ dl, const T* __restrict__ d,
                                                                    const T* __restrict__ du, const T* rhs, T* x,
                                                                    int nlev, size_t ncols, int pitch) {
  // (16-column tiles for both element types: with Float32, 32-column tiles measured 0.71 ms vs 0.55 ms per solve)
  constexpr int kCols = kTileCols;
  extern __shared__ __align__(128) char smem[];
  const int tile_bytes = kCols * pitch;
  char* sdl = smem;
  char* sd = smem + tile_bytes;
  char* sdu = smem + 2 * tile_bytes;
  char* sr = smem + 3 * tile_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * tile_bytes);

  const size_t col0 = (size_t)blockIdx.x * kCols;
  const int nvalid = (int)((ncols - col0) < (size_t)kCols ? (ncols - col0) : (size_t)kCols);
  const int col_bytes = nlev * (int)sizeof(T);
  const size_t goff = col0 * (size_t)nlev;

  if (LOADER == 1) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(bar, (uint32_t)(4 * nvalid * col_bytes));
    }
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
      const int c = threadIdx.x;
      const size_t g = goff + (size_t)c * nlev;
      bulk_g2s(sdl + c * pitch, dl + g, col_bytes, bar);
      bulk_g2s(sd + c * pitch, d + g, col_bytes, bar);
      bulk_g2s(sdu + c * pitch, du + g, col_bytes, bar);
      bulk_g2s(sr + c * pitch, rhs + g, col_bytes, bar);
    }
    mbar_wait(bar, 0);
  } else if (LOADER == 2) {
    cpasync_load_tile(sdl, reinterpret_cast<const char*>(dl + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sd, reinterpret_cast<const char*>(d + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sdu, reinterpret_cast<const char*>(du + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sr, reinterpret_cast<const char*>(rhs + goff), nvalid, col_bytes, pitch);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
  } else {
    coop_load_tile(sdl, reinterpret_cast<const char*>(dl + goff), nvalid, col_bytes, pitch);
    coop_load_tile(sd, reinterpret_cast<const char*>(d + goff), nvalid, col_bytes, pitch);
    coop_load_tile(sdu, reinterpret_cast<const char*>(du + goff), nvalid, col_bytes, pitch);
    coop_load_tile(sr, reinterpret_cast<const char*>(rhs + goff), nvalid, col_bytes, pitch);
    __syncthreads();
  }

  // lane pair (l, l+16) of warp w owns column 16*w + (l & 15)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = warp * 16 + lane_column(lane);
  const bool is_up = lane_is_up(lane);
  StoredRows<T> rows{sdl, sd, sdu};
  babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, true>(rows, RhsFromTile<T>{}, sr, col * pitch, nlev, is_up, col < nvalid);

  if (LOADER == 1) {
    fence_proxy_async_smem();  // generic-proxy writes of x -> visible to the async proxy
    __syncthreads();
    if ((int)threadIdx.x < nvalid) {
      const int c = threadIdx.x;
      bulk_s2g(x + goff + (size_t)c * nlev, sr + c * pitch, col_bytes);
      bulk_commit();
      bulk_wait_read0();
    }
  } else {
    __syncthreads();
    coop_store_tile(reinterpret_cast<char*>(x + goff), sr, nvalid, col_bytes, pitch);
  }
}

// Generic fallback (any nlev, any alignment): one thread per column straight from global memory,
// same operation sequence as the tile kernel.  `fac` is an n-element scratch for cp/lp.
template <typename T>
__global__ void tridiag_generic_kernel(const T* dl, const T* d, const T* du, const T* rhs, T* x, T* fac, int n,
                                       size_t ncols) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const size_t o = c * (size_t)n;
  dl += o; d += o; du += o; rhs += o; x += o; fac += o;
  if (n == 1) {
    x[0] = rmul(rhs[0], rdiv(T(1), d[0]));
    return;
  }
  const int m = n / 2;
  T fp = T(0), rp = T(0);
  {
    ProductSweep<T> sw;
    for (int i = 0; i < m; ++i) {
      sw.row(i == 0 ? T(0) : dl[i], d[i], du[i], rhs[i], i, fp, rp);
      fac[i] = fp;
      x[i] = rp;
    }
  }
  T cpm = fp, rpm = rp;
  fp = T(0); rp = T(0);
  {
    ProductSweep<T> sw;
    for (int j = n - 1, s = 0; j >= m; --j, ++s) {
      sw.row(j == n - 1 ? T(0) : du[j], d[j], dl[j], rhs[j], s, fp, rp);
      fac[j] = fp;
      x[j] = rp;
    }
  }
  T lpm = fp, spm = rp;
  T det = rsub(T(1), rmul(cpm, lpm));
  T xm1 = rdiv(rsub(rpm, rmul(cpm, spm)), det);
  T xm = rsub(spm, rmul(lpm, xm1));
  x[m - 1] = xm1;
  x[m] = xm;
  for (int i = m - 2; i >= 0; --i) x[i] = rsub(x[i], rmul(fac[i], x[i + 1]));
  for (int j = m + 1; j < n; ++j) x[j] = rsub(x[j], rmul(fac[j], x[j - 1]));
}

template <typename T>
__global__ void __launch_bounds__(256) tridiag_matvec_kernel(const T* __restrict__ dl, const T* __restrict__ d,
                                                             const T* __restrict__ du, const T* __restrict__ x,
                                                             T* __restrict__ y, int nlev, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lev = (int)(i % (size_t)nlev);
  T acc = rmul(d[i], x[i]);
  if (lev > 0) acc = radd(rmul(dl[i], x[i - 1]), acc);
  if (lev < nlev - 1) acc = radd(acc, rmul(du[i], x[i + 1]));
  y[i] = acc;
}

int tile_loader() {  // CTS_TILE_LOADER = 0 | 1 | 2 (default 1: bulk TMA, the fastest measured); CTS_DISABLE_TMA=1 forces 0
  static int v = -1;
  if (v < 0) {
    v = 1;
    const char* e = getenv("CTS_TILE_LOADER");
    if (e && e[0] >= '0' && e[0] <= '2') v = e[0] - '0';
    const char* d = getenv("CTS_DISABLE_TMA");
    if (d && d[0] == '1' && v == 1) v = 0;
  }
  return v;
}

template <typename T>
int solve_t(cts_ctx* ctx, const cts_tridiag* W, const void* rhs, void* x) {
  const int nlev = W->nlev;
  const size_t ncols = W->ncols;
  const int col_bytes = nlev * (int)sizeof(T);
  const int pitch = pitch_bytes(nlev, sizeof(T));
  constexpr int kCols = kTileCols;
  const size_t smem = (size_t)4 * kCols * pitch + 16;
  bool tile_ok = nlev >= 2 && (nlev % 2) == 0 && (col_bytes % 16) == 0 && smem <= 200 * 1024 &&
                 cts_aligned16(W->dl) && cts_aligned16(W->d) && cts_aligned16(W->du) && cts_aligned16(rhs) &&
                 cts_aligned16(x);
  if (tile_ok) {
    size_t tiles = (ncols + kCols - 1) / kCols;
    const int loader = tile_loader();
    auto launch = [&](auto k) -> int {
      CTS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<(unsigned)tiles, kTileThreads, smem, ctx->stream>>>((const T*)W->dl, (const T*)W->d, (const T*)W->du,
                                                              (const T*)rhs, (T*)x, nlev, ncols, pitch);
      return CTS_OK;
    };
    if (loader == 1)
      CTS_TRY(launch(tridiag_tile_kernel<T, 1>));
    else if (loader == 2)
      CTS_TRY(launch(tridiag_tile_kernel<T, 2>));
    else
      CTS_TRY(launch(tridiag_tile_kernel<T, 0>));
    CTS_LAUNCH_CHECK(ctx);
    return CTS_OK;
  }
  // generic path
  void* fac = nullptr;
  CTS_CUDA(ctx, cudaMallocAsync(&fac, sizeof(T) * ncols * (size_t)nlev, ctx->stream));
  size_t blocks = (ncols + 127) / 128;
  tridiag_generic_kernel<T><<<(unsigned)blocks, 128, 0, ctx->stream>>>((const T*)W->dl, (const T*)W->d, (const T*)W->du,
                                                                       (const T*)rhs, (T*)x, (T*)fac, nlev, ncols);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(fac, ctx->stream);
  if (e != cudaSuccess) return cts_fail(ctx, CTS_ECUDA, std::string("tridiag_generic: ") + cudaGetErrorString(e));
  return CTS_OK;
}

}  // namespace

extern "C" {

int cts_tridiag_solve_batched(cts_ctx* ctx, cts_dtype ft, const cts_tridiag* W, const void* rhs, void* x) {
  CTS_CHECK_CTX(ctx);
  if (!W || !rhs || !x || W->nlev <= 0 || !W->dl || !W->d || !W->du)
    return cts_fail(ctx, CTS_EINVAL, "cts_tridiag_solve_batched: bad argument");
  if (W->ncols == 0) return CTS_OK;
  CTS_COUNT_BYTES(ctx, 5, W->ncols * (size_t)W->nlev, cts_sizeof(ft));  // K4: dl, d, du, rhs in; x out
  return ft == CTS_F64 ? solve_t<double>(ctx, W, rhs, x) : solve_t<float>(ctx, W, rhs, x);
}

int cts_tridiag_matvec(cts_ctx* ctx, cts_dtype ft, const cts_tridiag* W, const void* x, void* y) {
  CTS_CHECK_CTX(ctx);
  if (!W || !x || !y || x == y) return cts_fail(ctx, CTS_EINVAL, "cts_tridiag_matvec: bad argument");
  size_t n = W->ncols * (size_t)W->nlev;
  if (n == 0) return CTS_OK;
  CTS_COUNT_BYTES(ctx, 5, n, cts_sizeof(ft));
  size_t blocks = (n + 255) / 256;
  if (ft == CTS_F64)
    tridiag_matvec_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)W->dl, (const double*)W->d,
                                                                             (const double*)W->du, (const double*)x,
                                                                             (double*)y, W->nlev, n);
  else
    tridiag_matvec_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)W->dl, (const float*)W->d,
                                                                            (const float*)W->du, (const float*)x,
                                                                            (float*)y, W->nlev, n);
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

}  // extern "C"
