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
      // mixes: does scalar FP32 work co-issue beside FFMA2 (heavy/lite halves of the FMA pipe)?
      if (MODE==10) { p[i] = ffma2(p[i], tp, p[(i+1)&7]); a[i] = __fmaf_rn(a[i], t, a[(i+1)&7]); }            // FFMA2 + FFMA
      if (MODE==11) { p[i] = ffma2(p[i], tp, p[(i+1)&7]); a[i] = a[i] * t; }                                   // FFMA2 + FMUL
      if (MODE==12) { p[i] = ffma2(p[i], tp, p[(i+1)&7]); a[i] = __uint_as_float(__float_as_uint(a[i]) ^ (it * 0x9e3779b9u)); } // FFMA2 + LOP3/IMAD
      if (MODE==13) { p[i] = ffma2(p[i], tp, p[(i+1)&7]); p[(i+3)&7] = ffma2(p[(i+3)&7], tp, p[(i+5)&7]); a[i] = __fmaf_rn(a[i], t, a[(i+1)&7]); } // 2 FFMA2 + 1 FFMA
      if (MODE==14) { a[i] = __fmaf_rn(a[i], t, a[(i+1)&7]); a[(i+4)&7] = a[(i+4)&7] * t; }                    // FFMA + FMUL
    }
  }
  long long t1 = clock64();
  float r=0; for (int i=0;i<8;i++){ r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i]>>32)); }
  out[blockIdx.x*blockDim.x+threadIdx.x] = r;
  if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
// legacy tensor path (mma.sync, SASS HMMA): issue rate of the shapes a batched 8x8 contraction could use
template<int MODE> __global__ void kmma(float* out, unsigned long long* cyc){
  float c[8][4]; unsigned a[4], b[2];
  for (int i=0;i<8;i++) for (int j=0;j<4;j++) c[i][j] = threadIdx.x*0.001f + i + j;
  for (int j=0;j<4;j++) a[j] = __float_as_uint(1.0f + threadIdx.x*0.01f + j);
  for (int j=0;j<2;j++) b[j] = __float_as_uint(0.5f + threadIdx.x*0.02f + j);
  long long t0 = clock64();
  #pragma unroll 1
  for (int it=0; it<N_ITER; it++){
    #pragma unroll
    for (int i=0;i<8;i++){
      if (MODE==0) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      if (MODE==1) asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
      if (MODE==2) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      if (MODE==3) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
    }
  }
  long long t1 = clock64();
  float r=0; for (int i=0;i<8;i++) for (int j=0;j<4;j++) r += c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x] = r;
  if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
template<int MODE> void runmma(const char* name, int warps, double macs){
  float* out; unsigned long long* cyc; cudaMalloc(&out, 1<<24); cudaMalloc(&cyc, 8);
  kmma<MODE><<<148, warps*32>>>(out, cyc); cudaDeviceSynchronize();
  kmma<MODE><<<148, warps*32>>>(out, cyc); cudaDeviceSynchronize();
  unsigned long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
  double inst = (double)N_ITER*8*warps;
  printf("%-22s warps/SM=%2d  cycles=%llu  warp-inst/clk/SM=%.3f  MAC/clk/SM=%.0f (FFMA2: 126)\n", name, warps, c, inst/c, inst/c*macs);
  cudaFree(out); cudaFree(cyc);
}
template<int MODE> void run(const char* name, int warps, int per_iter = 8){
  float* out; unsigned long long* cyc; cudaMalloc(&out, 1<<24); cudaMalloc(&cyc, 8);
  k<MODE><<<148, warps*32>>>(out, 1.0001f, cyc); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148, warps*32>>>(out, 1.0001f, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); unsigned long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
  double inst = (double)N_ITER*per_iter*warps;   // warp-instructions per SM
  printf("%-14s warps/SM=%2d  cycles=%llu  warp-inst/clk/SM=%.3f  (ms=%.3f)\n", name, warps, c, inst/c, ms);
  cudaFree(out); cudaFree(cyc);
}
int main(){
  for (int w : {4, 8, 16, 32}) {
    run<0>("FFMA reg", w); run<1>("FFMA imm", w); run<2>("FFMA2 reg", w); run<3>("FFMA2 imm", w);
    run<4>("FRND", w); run<5>("FADD.RZ", w); run<6>("LOP3+FADD", w); run<7>("fdiv_rn", w); run<8>("F2IP.U8+LOP3", w); run<9>("FMNMX x2", w);
    run<10>("FFMA2+FFMA", w, 16); run<11>("FFMA2+FMUL", w, 16); run<12>("FFMA2+LOP3", w, 16); run<13>("2FFMA2+FFMA", w, 24); run<14>("FFMA+FMUL", w, 16);
  }
  for (int w : {4, 8, 16, 32}) {
    runmma<0>("HMMA m16n8k8 tf32", w, 1024); runmma<1>("HMMA m16n8k4 tf32", w, 512);
    runmma<2>("HMMA m16n8k16 bf16", w, 2048); runmma<3>("HMMA m16n8k8 bf16", w, 1024);
  }
  return 0;
}
