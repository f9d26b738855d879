#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
int main(int argc,char**argv){
  int di=atoi(argv[1]); float d=(float)di, nd=-d, r=1.0f/d; float mx=0, mn=1e30;
  for (long long i=0;i<(1ll<<31);i++){
      uint32_t u=(uint32_t)i; float x; memcpy(&x,&u,4);
      if (!isfinite(x)) continue;
      float q0=x*r; float e=fmaf(q0,nd,x); float q=fmaf(e,r,q0); float ref=x/d;
      if (q!=ref){ if (x>mx) mx=x; if (x<mn) mn=x; }
  }
  printf("d=%d quotient mismatches for x in [%a, %a]\n",di,mn,mx); return 0; }
