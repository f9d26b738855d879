#!/bin/bash
export B200DCT_LIB_DIR=scratch/v_w12
for run in 1 4 8 32; do
  for w in 8 12; do
    B200DCT_TMA_RUN=$run B200DCT_TMA_WARPS=$w MODES=rt NPAIRS=8 python scratch/exp6.py "run$run w$w" 2>&1 | grep -E "rotate (1|4|8) pairs"
  done
  B200DCT_TMA_RUN=$run B200DCT_TMA_WARPS=8 MODES=fwd NPAIRS=8 python scratch/exp6.py "run$run w8" 2>&1 | grep -E "rotate (1|4|8) pairs"
done
