#!/bin/bash
export B200DCT_LIB_DIR=scratch/v_w12
for w in 8 9 10 12; do
  for run in 2 3; do
    B200DCT_TMA_WARPS=$w B200DCT_TMA_RUN=$run MODES=rt NPAIRS=8 python scratch/exp6.py "w$w run$run" 2>&1 | grep -E "rotate (1|8) pairs"
  done
done
B200DCT_TMA_WARPS=10 B200DCT_TMA_RUN=2 MODES=fwd NPAIRS=8 python scratch/exp6.py "w10 run2" 2>&1 | grep -E "rotate (1|8) pairs"
unset B200DCT_LIB_DIR
python scratch/exp2.py main_run2
