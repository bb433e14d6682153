// Does streaming global traffic issued by OTHER warps of the SM slow down the (shared-memory resident, latency-bound)
// tridiagonal solve?  4 solver warps per SM (one per sub-partition) + 16 "noise" warps that stream a large array
//   mode 0: no noise        mode 1: LDG.128 + STG.128 copy loop (through the LSU)       mode 2: bulk-TMA copy g2s -> s2g
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
using namespace cts_tile;

__global__ void __launch_bounds__(640, 1) k(int nlev, int pitch, int reps, int mode, const double2* src, double2* dst, size_t nvec,
                                            long long* out, volatile int* stop) {
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = 16 * pitch;
  __shared__ int done_solvers;
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) { done_solvers = 0; for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1); fence_mbar_init(); }
  __syncthreads();
  if (warp < 4) {
    char* base = smem + (size_t)warp * 4 * tile_bytes;
    char *sx = base, *sdl = base + tile_bytes, *sd = base + 2 * tile_bytes, *sdu = base + 3 * tile_bytes;
    const int col = lane_column(lane);
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
      for (int i = lane; i < 16 * nlev; i += 32) {
        const int c = i / nlev, l = i % nlev;
        reinterpret_cast<double*>(sdl + c * pitch)[l] = 0.5;
        reinterpret_cast<double*>(sdu + c * pitch)[l] = 0.25;
        reinterpret_cast<double*>(sd + c * pitch)[l] = -3.0 - 0.01 * l;
        reinterpret_cast<double*>(sx + c * pitch)[l] = 1.0 + 0.001 * (c + l);
      }
      __syncwarp();
      const long long t0 = clock64();
      StoredRows<double> rows{sdl, sd, sdu};
      babe_solve_lanepair<double, StoredRows<double>, RhsFromTile<double>, true>(rows, RhsFromTile<double>{}, sx, col * pitch, nlev, lane_is_up(lane), true);
      __syncwarp();
      total += clock64() - t0;
    }
    if (lane == 0) { out[blockIdx.x * 4 + warp] = total / reps; atomicAdd(&done_solvers, 1); }
    if (lane == 0 && blockIdx.x == 0) out[4096 + warp] = (long long)(reinterpret_cast<double*>(sx)[3] * 1e6);
  } else {
    // noise warps: stream until the four solver warps of this CTA are done
    const int nw = warp - 4;  // 0..15
    char* stage = smem + 4 * 4 * tile_bytes + (size_t)nw * 4096;  // 4 KiB per noise warp (mode 2)
    size_t pos = ((size_t)blockIdx.x * 16 + nw) * 512;            // in 16-byte vectors; each chunk = 512 vectors = 8 KiB
    const size_t stride = (size_t)gridDim.x * 16 * 512;
    long long moved = 0;
    int phase = 0;
    while (*((volatile int*)&done_solvers) < 4) {
      if (pos + 512 > nvec) pos = ((size_t)blockIdx.x * 16 + nw) * 512;
      if (mode == 1) {
        double2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = src[pos + lane + 32 * j];
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[pos + lane + 32 * j] = v[j];
      } else if (mode == 2) {
        if (lane == 0) {
          mbar_arrive_expect_tx(bars + nw, 4096);
          bulk_g2s(stage, src + pos, 4096, bars + nw);
          mbar_wait(bars + nw, phase);
          fence_proxy_async_smem();
          bulk_s2g(dst + pos, stage, 4096);
          bulk_commit();
          bulk_wait_read0();
        }
        phase ^= 1;
        __syncwarp();
      } else {
        __nanosleep(200);
      }
      pos += stride;
      moved += mode == 2 ? 4096 : 8192;
    }
    if (lane == 0) atomicAdd((unsigned long long*)&out[8000], (unsigned long long)moved);
  }
}

int main() {
  const int nlev = 64, pitch = pitch_bytes(nlev, 8), reps = 200, ctas = 148;
  const size_t nvec = (size_t)1 << 27;  // 2 GiB per array
  double2 *src, *dst;
  cudaMalloc(&src, nvec * 16); cudaMalloc(&dst, nvec * 16); cudaMemset(src, 0, nvec * 16);
  long long* d; cudaMalloc(&d, 8192 * sizeof(long long));
  const size_t smem = (size_t)4 * 4 * 16 * pitch + 16 * 4096;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(d, 0, 8192 * sizeof(long long));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<<<ctas, 640, smem>>>(nlev, pitch, reps, mode, src, dst, nvec, d, nullptr);
    cudaEventRecord(b);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = e2;
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(8192);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < ctas * 4; ++i) s += (double)h[i];
    printf("mode %d (%s): %8.0f cycles per solve; noise traffic %.2f TB/s (read+write) over %.2f ms  %s\n", mode,
           mode == 0 ? "no noise" : mode == 1 ? "LDG/STG copy by 16 warps" : "bulk-TMA copy by 16 warps", s / (ctas * 4),
           2.0 * (double)h[8000] / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(e));
  }
  return 0;
}
