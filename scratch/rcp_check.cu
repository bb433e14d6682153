// check: rcp_checked(x) == __drcp_rn(x) (IEEE) bit for bit on the safe range
#include <cstdio>
#include <cstdint>
#include "../climatimesteppers.jl_b200/csrc/tile_device.cuh"
__device__ uint64_t sm64(uint64_t x){x+=0x9E3779B97F4A7C15ull;x=(x^(x>>30))*0xBF58476D1CE4E5B9ull;x=(x^(x>>27))*0x94D049BB133111EBull;return x^(x>>31);}
__global__ void chk(unsigned long long* bad, unsigned long long* badf, uint64_t seed){
  uint64_t i = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;
  for (int k=0;k<64;++k){
    uint64_t h = sm64(seed ^ (i*64+k));
    uint64_t be = 25 + (sm64(h) % (2021-25+1));
    uint64_t bits = (h & 0x800fffffffffffffull) | (be << 52);
    double x = __longlong_as_double((long long)bits);
    double a = cts_tile::rcp_checked(x), b = __drcp_rn(x);
    if (__double_as_longlong(a) != __double_as_longlong(b)) atomicAdd(bad, 1ull);
    uint32_t fb = (uint32_t)(h>>32); uint32_t fe = 1 + ((fb>>23)&0xff)%254; fb = (fb & 0x807fffffu) | (fe<<23);
    float xf = __int_as_float((int)fb);
    float af = cts_tile::rcp_checked(xf), bf = __frcp_rn(xf);
    if (__float_as_int(af) != __float_as_int(bf)) atomicAdd(badf, 1ull);
  }
}
int main(){ unsigned long long *b,*bf,h[2]; cudaMalloc(&b,8); cudaMalloc(&bf,8); cudaMemset(b,0,8); cudaMemset(bf,0,8);
  for(int r=0;r<8;++r) chk<<<148*64,256>>>(b,bf,1234567+r); cudaDeviceSynchronize();
  cudaMemcpy(&h[0],b,8,cudaMemcpyDeviceToHost); cudaMemcpy(&h[1],bf,8,cudaMemcpyDeviceToHost);
  printf("checked %llu values: f64 mismatches vs __drcp_rn = %llu, f32 mismatches vs __frcp_rn = %llu (%s)\n", 8ull*148*64*256*64, h[0], h[1], cudaGetErrorString(cudaGetLastError())); return 0; }
