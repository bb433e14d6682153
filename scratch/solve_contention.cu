// How much does the product-form Thomas solve (one warp = 16 columns x 64 levels, data in shared memory) slow down when
// several solver warps share ONE SM sub-partition (same %warpid % 4), compared with one solver per sub-partition?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
using namespace cts_tile;

// mode 0: the first `nsolve` warps with (%warpid % 4 == 0) solve, all others exit.   mode 1: warps with %warpid % 4 == (slot index) ...
template <bool PP>
__global__ void k(int nlev, int pitch, int reps, int nsolve, int same_smsp, long long* out, int* wid) {
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned hw;
  asm("mov.u32 %0, %%warpid;" : "=r"(hw));
  // pick the solver warps: same_smsp -> hardware warp ids 0,4,8,... (one sub-partition); else 0,1,2,3 (one per sub-partition)
  int slot = -1;
  if (same_smsp) { if (hw % 4 == 0 && (int)(hw / 4) < nsolve) slot = hw / 4; }
  else           { if ((int)hw < nsolve) slot = hw; }
  if (lane == 0 && blockIdx.x == 0) wid[warp] = (int)hw;
  if (slot < 0) return;
  const int tile_bytes = 16 * pitch;
  char* base = smem + (size_t)slot * 4 * tile_bytes;
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
    babe_solve_lanepair<double, StoredRows<double>, RhsFromTile<double>, PP>(rows, RhsFromTile<double>{}, sx, col * pitch, nlev, lane_is_up(lane), true);
    __syncwarp();
    total += clock64() - t0;
  }
  if (lane == 0) out[blockIdx.x * 8 + slot] = total / reps;
  if (lane == 0 && blockIdx.x == 0) out[4096 + slot] = (long long)(reinterpret_cast<double*>(sx)[3] * 1e6);
}

template <bool PP>
void run(const char* name, int nsolve, int same) {
  const int nlev = 64, pitch = pitch_bytes(nlev, 8), reps = 50, ctas = 148, warps = 16;
  const size_t smem = (size_t)nsolve * 4 * 16 * pitch;
  long long* d; int* w;
  cudaMalloc(&d, 8192 * sizeof(long long)); cudaMemset(d, 0, 8192 * sizeof(long long)); cudaMalloc(&w, 64 * sizeof(int));
  auto kk = k<PP>;
  cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kk<<<ctas, 32 * warps, smem>>>(nlev, pitch, reps, nsolve, same, d, w);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(ctas * 8);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0; int n = 0;
  for (auto v : h) if (v > 0) { s += (double)v; ++n; }
  std::vector<int> hw(16); cudaMemcpy(hw.data(), w, 16 * sizeof(int), cudaMemcpyDeviceToHost);
  printf("%-10s %d solver warps per SM, %s: %8.0f cycles per solve (%d samples) %s   [hw warp ids of CTA 0:", name, nsolve,
         same ? "ALL on one sub-partition " : "one per sub-partition    ", n ? s / n : 0.0, n, cudaGetErrorString(e));
  for (int i = 0; i < 16; ++i) printf(" %d", hw[i]);
  printf("]\n");
  cudaFree(d); cudaFree(w);
}

int main() {
  for (int n : {1, 2, 4}) {
    run<false>("copy-pipe", n, 0);
    run<false>("copy-pipe", n, 1);
    run<true>("ping-pong", n, 0);
    run<true>("ping-pong", n, 1);
  }
  return 0;
}
