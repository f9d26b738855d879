#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c){
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2rz(float2 a, float2 b){
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__global__ void k(const float4* in, float4* out){
  float4 v = in[threadIdx.x];
  float4 w = in[threadIdx.x+32];
  float2 a = make_float2(v.x, v.y), b = make_float2(v.z, v.w);
  float2 t = make_float2(0.35355339f, 0.35355339f);
  float2 nt = make_float2(-0.35355339f, -0.35355339f);
  float2 acc = ffma2(a, t, make_float2(0.f,0.f));
  acc = ffma2(b, nt, acc);
  // broadcast operand
  float2 bb = make_float2(w.x, w.x);
  acc = ffma2(bb, make_float2(0.5f, -0.5f), acc);
  // mixed pair from different regs
  float2 mx = make_float2(v.x, w.y);
  acc = ffma2(mx, t, acc);
  acc = fadd2rz(acc, make_float2(0.5f,0.5f));
  out[threadIdx.x] = make_float4(acc.x, acc.y, truncf(acc.x), roundf(acc.y)/16.0f);
}
