#!/bin/bash
# warps sweep of the TMA kernel on two builds
for lib in "" scratch/v_w12; do
  for w in 4 5 6 7 8 9; do
    B200DCT_LIB_DIR=${lib:-cuda-dct-idct_b200} B200DCT_TMA_WARPS=$w ITERS=200 python benchmarks/experiments/exp_time.py "${lib:-main}_w$w" 2>&1 | grep tma
  done
done
for n in 1024 2048 4096 16384; do
  for w in 6 8; do N=$n ITERS=100 B200DCT_TMA_WARPS=$w python benchmarks/experiments/exp_time.py "main_w${w}" 2>&1 | grep -E "f32"; done
done
