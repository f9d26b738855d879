#!/bin/bash
# usage: mix.sh  -> compiles scratch/one.cu and prints instruction mix of k_tma RT kernel
cd /root/repo/cuda-dct-idct_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I. -I../../include -Xptxas -v $EXTRA -cubin -o /root/repo/scratch/one.cubin /root/repo/scratch/one.cu 2>&1 | grep -E "registers|spill|error" 
cuobjdump -sass /root/repo/scratch/one.cubin > /root/repo/scratch/one.sass
for k in k_tma k_direct; do echo "== $k"; awk -v k=$k '/Function :/{f=($0 ~ k)} f' /root/repo/scratch/one.sass | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${1:-12} | tr '\n' ' '; echo; awk -v k=$k '/Function :/{f=($0 ~ k)} f' /root/repo/scratch/one.sass | grep -cE "^\s+/\*[0-9a-f]{4}\*/"; done
