// pipe-throughput microbenchmarks (scratch; informs the kernel design)
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 4096
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c){
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template<int MODE> __global__ void k(float* out, float s, unsigned long long* cyc){
  float a[8]; unsigned long long p[8];
  for (int i=0;i<8;i++){ a[i] = threadIdx.x*0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i])<<32) | __float_as_uint(a[i]+1.f);}
  float t = s; unsigned long long tp = ((unsigned long long)__float_as_uint(s)<<32)|__float_as_uint(s);
  long long t0 = clock64();
  #pragma unroll 1
  for (int it=0; it<N_ITER; it++){
    #pragma unroll
    for (int i=0;i<8;i++){
      if (MODE==0) a[i] = __fmaf_rn(a[i], t, a[(i+1)&7]);               // FFMA 3 regs
      if (MODE==1) a[i] = __fmaf_rn(a[i], 0.35355339f, a[(i+1)&7]);     // FFMA imm
      if (MODE==2) p[i] = ffma2(p[i], tp, p[(i+1)&7]);                  // FFMA2 regs
      if (MODE==3) { unsigned long long c2 = 0x3eb504f33eb504f3ull; p[i] = ffma2(p[i], c2, p[(i+1)&7]); } // FFMA2 imm
      if (MODE==4) a[i] = truncf(a[i]);                                 // FRND
      if (MODE==5) a[i] = __fadd_rz(a[i], t);                           // FADD.RZ
      if (MODE==6) a[i] = __uint_as_float(__float_as_uint(a[i]) & 0x80000000u | 0x3f000000u) + a[(i+1)&7]; // LOP3+FADD
      if (MODE==7) a[i] = __fdiv_rn(a[i], t);                           // true division
      if (MODE==8) { unsigned r; asm volatile("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r | 0x3f800000u); } // F2IP + LOP3
      if (MODE==9) a[i] = fminf(fmaxf(a[i], 0.f), t);                   // 2x FMNMX
    }
  }
  long long t1 = clock64();
  float r=0; for (int i=0;i<8;i++){ r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i]>>32)); }
  out[blockIdx.x*blockDim.x+threadIdx.x] = r;
  if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
template<int MODE> void run(const char* name, int warps){
  float* out; unsigned long long* cyc; cudaMalloc(&out, 1<<24); cudaMalloc(&cyc, 8);
  k<MODE><<<148, warps*32>>>(out, 1.0001f, cyc); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148, warps*32>>>(out, 1.0001f, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); unsigned long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
  double inst = (double)N_ITER*8*warps;   // warp-instructions per SM
  printf("%-14s warps/SM=%2d  cycles=%llu  warp-inst/clk/SM=%.3f  (ms=%.3f)\n", name, warps, c, inst/c, ms);
  cudaFree(out); cudaFree(cyc);
}
int main(){
  for (int w : {4, 8, 16, 32}) {
    run<0>("FFMA reg", w); run<1>("FFMA imm", w); run<2>("FFMA2 reg", w); run<3>("FFMA2 imm", w);
    run<4>("FRND", w); run<5>("FADD.RZ", w); run<6>("LOP3+FADD", w); run<7>("fdiv_rn", w); run<8>("F2IP.U8+LOP3", w); run<9>("FMNMX x2", w);
  }
  return 0;
}
