// What can a 4B-read + 4B-write stream reach on this B200, per access pattern? (scratch)
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
constexpr int N = 8192;
// A: linear float4 copy, one CTA per 4096 floats chunk, 256 threads x 4 float4
__global__ void __launch_bounds__(256) copy_linear(const float4* __restrict__ in, float4* __restrict__ out, size_t n4){
  size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  float4 v[4];
  #pragma unroll
  for (int k=0;k<4;k++) v[k] = __ldg(in + i + k*256);
  #pragma unroll
  for (int k=0;k<4;k++) out[i + k*256] = v[k];
}
// A2: same, 16 float4 per thread (like a block per thread: 256 B in flight per thread), coalesced
__global__ void __launch_bounds__(128) copy_linear16(const float4* __restrict__ in, float4* __restrict__ out){
  size_t i = (size_t)blockIdx.x * 2048 + threadIdx.x;
  float4 v[16];
  #pragma unroll
  for (int k=0;k<16;k++) v[k] = __ldg(in + i + k*128);
  #pragma unroll
  for (int k=0;k<16;k++) out[i + k*128] = v[k];
}
// C: block pattern (the direct kernel's access pattern, no math): thread = 8x8 block
__global__ void __launch_bounds__(128) copy_blocks(const float* __restrict__ in, float* __restrict__ out){
  const int bx = blockIdx.y*32 + threadIdx.x; const size_t by = (size_t)blockIdx.x*4 + threadIdx.y;
  const float4* src = reinterpret_cast<const float4*>(in + by*8*N + bx*8);
  float4* dst = reinterpret_cast<float4*>(out + by*8*N + bx*8);
  float4 v[16];
  #pragma unroll
  for (int r=0;r<8;r++){ v[2*r] = __ldg(src + r*(N/4)); v[2*r+1] = __ldg(src + r*(N/4) + 1); }
  #pragma unroll
  for (int r=0;r<8;r++){ dst[r*(N/4)] = v[2*r]; dst[r*(N/4)+1] = v[2*r+1]; }
}
// D: tile pattern, coalesced: warp = 8 rows x 1 KiB, lane reads chunk l and l+32 of each row
__global__ void __launch_bounds__(128) copy_tiles(const float* __restrict__ in, float* __restrict__ out){
  const int tx = blockIdx.y; const size_t by = (size_t)blockIdx.x*4 + threadIdx.y;
  const float4* src = reinterpret_cast<const float4*>(in + by*8*N + tx*256) + threadIdx.x;
  float4* dst = reinterpret_cast<float4*>(out + by*8*N + tx*256) + threadIdx.x;
  float4 v[16];
  #pragma unroll
  for (int r=0;r<8;r++){ v[2*r] = __ldg(src + r*(N/4)); v[2*r+1] = __ldg(src + r*(N/4) + 32); }
  #pragma unroll
  for (int r=0;r<8;r++){ dst[r*(N/4)] = v[2*r]; dst[r*(N/4)+32] = v[2*r+1]; }
}
// E: persistent grid-stride linear copy
__global__ void __launch_bounds__(512) copy_persist(const float4* __restrict__ in, float4* __restrict__ out, size_t n4){
  for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x*blockDim.x*4){
    float4 v[4]; size_t s = (size_t)gridDim.x*blockDim.x;
    #pragma unroll
    for (int k=0;k<4;k++) if (i+k*s<n4) v[k] = __ldg(in+i+k*s);
    #pragma unroll
    for (int k=0;k<4;k++) if (i+k*s<n4) out[i+k*s] = v[k];
  }
}
template<class F> void timeit(const char* name, F launch){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i=0;i<5;i++) launch(i);
  cudaDeviceSynchronize();
  float best=1e9;
  for (int rep=0;rep<3;rep++){ cudaEventRecord(e0); for (int i=0;i<200;i++) launch(i); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if (ms/200<best) best=ms/200; }
  printf("%-16s %8.1f us  %8.1f GB/s  (%s)\n", name, best*1e3, 2.0*N*N*4/best/1e6, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  float *in[4], *out[4]; size_t bytes=(size_t)N*N*4;
  for (int i=0;i<4;i++){ cudaMalloc(&in[i],bytes); cudaMalloc(&out[i],bytes); cudaMemset(in[i],i+1,bytes); }
  size_t n4=(size_t)N*N/4;
  timeit("cudaMemcpyD2D", [&](int i){ cudaMemcpyAsync(out[i%4],in[i%4],bytes,cudaMemcpyDeviceToDevice); });
  timeit("linear f4x4", [&](int i){ copy_linear<<<n4/1024,256>>>((float4*)in[i%4],(float4*)out[i%4],n4); });
  timeit("linear f4x16", [&](int i){ copy_linear16<<<n4/2048,128>>>((float4*)in[i%4],(float4*)out[i%4]); });
  timeit("blocks 8x8/thr", [&](int i){ copy_blocks<<<dim3(N/8/4, N/8/32),dim3(32,4)>>>(in[i%4],out[i%4]); });
  timeit("tiles coalesced", [&](int i){ copy_tiles<<<dim3(N/8/4, N/256),dim3(32,4)>>>(in[i%4],out[i%4]); });
  for (int g : {148, 296, 592, 1184}) { char nm[32]; snprintf(nm,32,"persist g=%d",g); timeit(nm, [&](int i){ copy_persist<<<g,512>>>((float4*)in[i%4],(float4*)out[i%4],n4); }); }
  return 0;
}
