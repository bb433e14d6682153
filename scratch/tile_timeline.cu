// Phase timeline of the tridiagonal tile kernel: where do the ~20k cycles per tile go?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
using namespace cts_tile;
template <int LOADER>
__global__ void __launch_bounds__(kTileThreads) k(const double* dl, const double* d, const double* du, const double* rhs,
                                                  double* x, int nlev, size_t ncols, int pitch, long long* ts) {
  extern __shared__ __align__(128) char smem[];
  long long t0 = clock64();
  const int tile_bytes = kTileCols * pitch;
  char *sdl = smem, *sd = smem + tile_bytes, *sdu = smem + 2 * tile_bytes, *sr = smem + 3 * tile_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * tile_bytes);
  const size_t col0 = (size_t)blockIdx.x * kTileCols;
  const int nvalid = 32, col_bytes = nlev * 8;
  const size_t goff = col0 * (size_t)nlev;
  if (LOADER == 1) {
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); mbar_arrive_expect_tx(bar, 4 * nvalid * col_bytes); }
    __syncthreads();
    if (threadIdx.x < nvalid) {
      int c = threadIdx.x; size_t g = goff + (size_t)c * nlev;
      bulk_g2s(sdl + c * pitch, dl + g, col_bytes, bar); bulk_g2s(sd + c * pitch, d + g, col_bytes, bar);
      bulk_g2s(sdu + c * pitch, du + g, col_bytes, bar); bulk_g2s(sr + c * pitch, rhs + g, col_bytes, bar);
    }
    mbar_wait(bar, 0);
  } else {
    cpasync_load_tile(sdl, (const char*)(dl + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sd, (const char*)(d + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sdu, (const char*)(du + goff), nvalid, col_bytes, pitch);
    cpasync_load_tile(sr, (const char*)(rhs + goff), nvalid, col_bytes, pitch);
    cp_async_commit(); cp_async_wait_all(); __syncthreads();
  }
  long long t1 = clock64();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = warp * 16 + lane_column(lane);
  StoredRows<double> rows{sdl, sd, sdu};
  babe_solve_lanepair<double>(rows, RhsFromTile<double>{}, sr, col * pitch, nlev, lane_is_up(lane), true);
  __syncthreads();
  long long t2 = clock64();
  coop_store_tile((char*)(x + goff), sr, nvalid, col_bytes, pitch);
  long long t3 = clock64();
  if (threadIdx.x == 0) { ts[blockIdx.x * 4 + 0] = t0; ts[blockIdx.x * 4 + 1] = t1; ts[blockIdx.x * 4 + 2] = t2; ts[blockIdx.x * 4 + 3] = t3; }
}
int main(int argc, char** argv) {
  int loader = argc > 1 ? atoi(argv[1]) : 1;
  const int nlev = 64; const size_t ncols = 1 << 21, n = ncols * nlev;
  double *dl, *d, *du, *r, *x; long long* ts;
  cudaMalloc(&dl, n * 8); cudaMalloc(&d, n * 8); cudaMalloc(&du, n * 8); cudaMalloc(&r, n * 8); cudaMalloc(&x, n * 8);
  size_t tiles = ncols / 32; cudaMalloc(&ts, tiles * 32);
  std::vector<double> h(n, 0.5); cudaMemcpy(dl, h.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(du, h.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(r, h.data(), n * 8, cudaMemcpyHostToDevice); std::fill(h.begin(), h.end(), -3.0); cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
  int pitch = pitch_bytes(nlev, 8); size_t smem = 4 * 32 * pitch + 16;
  auto kk = loader == 1 ? k<1> : k<2>;
  cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a); kk<<<tiles, 64, smem>>>(dl, d, du, r, x, nlev, ncols, pitch, ts); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b); printf("loader %d: %.3f ms  err=%s\n", loader, ms, cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<long long> t(tiles * 4); cudaMemcpy(t.data(), ts, tiles * 32, cudaMemcpyDeviceToHost);
  double s1 = 0, s2 = 0, s3 = 0; long long tmin = t[0], tmax = 0;
  for (size_t i = 0; i < tiles; ++i) { s1 += t[4*i+1]-t[4*i]; s2 += t[4*i+2]-t[4*i+1]; s3 += t[4*i+3]-t[4*i+2]; tmin = std::min(tmin, t[4*i]); tmax = std::max(tmax, t[4*i+3]); }
  printf("avg cycles per tile: load %.0f  compute %.0f  store-issue %.0f ; kernel span %lld cycles; tiles/SM-slot %.1f\n", s1 / tiles, s2 / tiles, s3 / tiles, tmax - tmin, tiles / (148.0 * 3));
  return 0;
}
