#!/bin/bash
L=scratch/v_w12
run(){ B200DCT_LIB_DIR=$L ITERS=200 "$@" python benchmarks/experiments/exp_time.py "$*" 2>&1 | grep "f32 tma"; }
run env B200DCT_TMA_WARPS=8
run env B200DCT_TMA_WARPS=8 B200DCT_TMA_L2PROMO=0
run env B200DCT_TMA_WARPS=8 B200DCT_TMA_L2PROMO=2
run env B200DCT_TMA_WARPS=10
run env B200DCT_TMA_WARPS=10 B200DCT_TMA_L2PROMO=0
run env B200DCT_TMA_WARPS=10 B200DCT_TMA_GRID=118
run env B200DCT_TMA_WARPS=12 B200DCT_TMA_GRID=98
run env B200DCT_TMA_WARPS=5 B200DCT_TMA_GRID=296
run env B200DCT_TMA_WARPS=4 B200DCT_TMA_GRID=296
run env B200DCT_TMA_WARPS=8 B200DCT_TMA_GRID=140
for n in 8448 12288; do N=$n B200DCT_LIB_DIR=$L B200DCT_TMA_WARPS=8 python benchmarks/experiments/exp_time.py "N$n w8" 2>&1 | grep "f32"; N=$n B200DCT_LIB_DIR=$L B200DCT_TMA_WARPS=12 python benchmarks/experiments/exp_time.py "N$n w12" 2>&1 | grep "f32 tma"; done
