#include <cstdio>
#include <cuda_runtime.h>
// dependent-chain latency and multi-warp throughput of FP64 ops on this GPU
template <int OP>
__global__ void lat(double* out, long long* cyc, double a, double b, int iters) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = __fma_rn(x, a, b);
    if (OP == 1) x = __dmul_rn(x, a);
    if (OP == 2) x = __dadd_rn(x, b);
    if (OP == 3) x = __drcp_rn(x);
    if (OP == 4) { x = __dsub_rn(b, __dmul_rn(a, x)); x = __drcp_rn(x); }   // Thomas pivot chain
    if (OP == 5) { float y = (float)x; y = __fmaf_rn(y, 1.0001f, 0.5f); x = y; }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&c, 8); cudaMemset(d, 0, 1 << 22);
  const char* names[] = {"DFMA", "DMUL", "DADD", "DRCP", "pivot(DMUL,DSUB,DRCP)", "F2F+FFMA+F2F"};
  int iters = 4096;
  for (int op = 0; op < 6; ++op) {
    for (int warps : {1, 4, 8, 16, 32}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: lat<0><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters); break;
          case 1: lat<1><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters); break;
          case 2: lat<2><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters); break;
          case 3: lat<3><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters); break;
          case 4: lat<4><<<1, 32 * warps>>>(d, c, 0.3, 2.5, iters); break;
          case 5: lat<5><<<1, 32 * warps>>>(d, c, 0.3, 2.5, iters); break;
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      }
      printf("%-24s warps/SM=%2d  cycles/iter=%.1f\n", names[op], warps, (double)h / iters);
    }
  }
  return 0;
}
