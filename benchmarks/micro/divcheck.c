#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
int main(int argc,char**argv){
  for (int di=(argc>1?atoi(argv[1]):1); di<=255; di++){
    float d=(float)di, nd=-d, r=1.0f/d; unsigned long long badq=0,badc=0; int shown=0;
    #pragma omp parallel for reduction(+:badq,badc)
    for (long long i=0;i<(1ll<<32);i++){
      uint32_t u=(uint32_t)i; float x; memcpy(&x,&u,4);
      if (!isfinite(x)) continue;
      float q0=x*r; float e=fmaf(q0,nd,x); float q=fmaf(e,r,q0); float ref=x/d;
      uint32_t a,b; memcpy(&a,&q,4); memcpy(&b,&ref,4);
      if (a!=b) badq++;
      float ca=roundf(q), cb=roundf(ref); memcpy(&a,&ca,4); memcpy(&b,&cb,4);
      if (a!=b){ badc++; 
        #pragma omp critical
        if (shown<4){ shown++; printf("  d=%d x=%a (%.9g, bits %08x) q=%a ref=%a\n",di,x,x,u,q,ref);} }
    }
    if (badq||badc) printf("d=%d badq=%llu badc=%llu\n",di,badq,badc);
    if (argc>1) break;
  }
  return 0;
}
