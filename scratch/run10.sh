#!/bin/bash
L=scratch/v_w12
for pad in 0 32 256 1024 4096; do PAD=$pad python scratch/exp3.py w8; done
for pad in 0 32 256 1024; do PAD=$pad B200DCT_LIB_DIR=$L B200DCT_TMA_WARPS=12 MODES=rt python scratch/exp3.py w12; done
MODES=fwd LAUNCHES=6 python scratch/exp3.py x > gpurun_out/plain_fwd.log 2>&1 && ncu --set full --clock-control none -k regex:k_tma -s 2 -c 2 -o gpurun_out/prof_tma_fwd_w8 env MODES=fwd LAUNCHES=6 python scratch/exp3.py x > gpurun_out/ncu_fwd.log 2>&1
B200DCT_TMA_WARPS=6 MODES=fwd LAUNCHES=6 python scratch/exp3.py x > gpurun_out/plain_fwd6.log 2>&1 && B200DCT_TMA_WARPS=6 ncu --set full --clock-control none -k regex:k_tma -s 2 -c 2 -o gpurun_out/prof_tma_fwd_w6 env MODES=fwd LAUNCHES=6 python scratch/exp3.py x > gpurun_out/ncu_fwd6.log 2>&1
