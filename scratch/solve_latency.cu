// Pure latency of the two-sided product-form Thomas solve (tile_device.cuh) for ONE warp = 16 columns x 64 levels, data
// resident in shared memory, no global traffic: how many cycles does a solve take alone on an SM, and with 2/4/8 solver
// warps sharing the SM?  (The fused stage kernels are bounded by this number: DESIGN.md §5.1.)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
using namespace cts_tile;

template <typename T, bool PP>
__global__ void k(int nlev, int pitch, int reps, long long* out) {
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = 16 * pitch;
  char* base = smem + (size_t)warp * 4 * tile_bytes;  // per warp: x | dl | d | du
  char *sx = base, *sdl = base + tile_bytes, *sd = base + 2 * tile_bytes, *sdu = base + 3 * tile_bytes;
  const int col = lane_column(lane);
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    for (int i = lane; i < 16 * nlev; i += 32) {  // refill (the solve overwrites its inputs)
      const int c = i / nlev, l = i % nlev;
      reinterpret_cast<T*>(sdl + c * pitch)[l] = T(0.5);
      reinterpret_cast<T*>(sdu + c * pitch)[l] = T(0.25);
      reinterpret_cast<T*>(sd + c * pitch)[l] = T(-3.0) - T(0.01) * T(l);
      reinterpret_cast<T*>(sx + c * pitch)[l] = T(1.0) + T(0.001) * T(c + l);
    }
    __syncwarp();
    const long long t0 = clock64();
    StoredRows<T> rows{sdl, sd, sdu};
    babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, PP>(rows, RhsFromTile<T>{}, sx, col * pitch, nlev, lane_is_up(lane), true);
    __syncwarp();
    total += clock64() - t0;
  }
  if (lane == 0) out[blockIdx.x * (blockDim.x / 32) + warp] = total / reps;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[4096] = (long long)(reinterpret_cast<T*>(sx)[3] * 1e6);  // keep the result alive
}

template <typename T, bool PP>
void run(const char* name, int warps, int ctas) {
  const int nlev = 64, pitch = pitch_bytes(nlev, sizeof(T)), reps = 50;
  const size_t smem = (size_t)warps * 4 * 16 * pitch;
  long long* d;
  cudaMalloc(&d, 8192 * sizeof(long long));
  auto kk = k<T, PP>;
  cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kk<<<ctas, 32 * warps, smem>>>(nlev, pitch, reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(ctas * warps);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (auto v : h) s += (double)v;
  printf("%-28s warps/SM %d: %8.0f cycles per solve (16 columns x %d levels)  %s\n", name, warps, s / h.size(), nlev, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int w : {1, 2, 4, 6}) {
    run<double, false>("f64 copy-pipelined", w, 148);
    run<double, true>("f64 ping-pong", w, 148);
  }
  run<float, false>("f32 copy-pipelined", 1, 148);
  run<float, true>("f32 ping-pong", 1, 148);
  run<float, true>("f32 ping-pong", 4, 148);
  return 0;
}
