// div_const(x, prepare(d)) == __ddiv_rn(x, d) bit for bit, for random x (all exponents, incl. zero/denormal/inf) and random d
#include <cstdio>
#include <cstdint>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
__device__ uint64_t sm64(uint64_t x){x+=0x9E3779B97F4A7C15ull;x=(x^(x>>30))*0xBF58476D1CE4E5B9ull;x=(x^(x>>27))*0x94D049BB133111EBull;return x^(x>>31);}
__global__ void chk(unsigned long long* bad, unsigned long long* fast, uint64_t seed){
  uint64_t i = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;
  uint64_t hd = sm64(seed*7919 + blockIdx.x);               // one divisor per block, exponent near 1 (dtγ-like) or anywhere
  uint64_t bed = (blockIdx.x & 1) ? 1000 + (hd % 46) : 1 + (hd % 2046);
  double d = __longlong_as_double((long long)((hd & 0x800fffffffffffffull) | (bed << 52)));
  cts_tile::DivConst dc = cts_tile::div_const_prepare(d);
  for (int k=0;k<64;++k){
    uint64_t h = sm64(seed ^ (i*64+k));
    uint64_t be = (k & 3) ? 900 + (sm64(h) % 250) : (sm64(h) % 2048);
    double x = __longlong_as_double((long long)((h & 0x800fffffffffffffull) | (be << 52)));
    double a = cts_tile::div_const(x, dc), b = __ddiv_rn(x, d);
    bool same = __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
    if (!same) atomicAdd(bad, 1ull);
  }
}
int main(){ unsigned long long *b,*f,h; cudaMalloc(&b,8); cudaMalloc(&f,8); cudaMemset(b,0,8);
  for(int r=0;r<8;++r) chk<<<148*64,256>>>(b,f,424242+r); cudaDeviceSynchronize();
  cudaMemcpy(&h,b,8,cudaMemcpyDeviceToHost);
  printf("checked %llu quotients: mismatches vs __ddiv_rn = %llu (%s)\n", 8ull*148*64*256*64, h, cudaGetErrorString(cudaGetLastError())); return 0; }
