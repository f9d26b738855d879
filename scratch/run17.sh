#!/bin/bash
for run in 1 2 4 8 32; do
  B200DCT_TMA_RUN=$run MODES=rt NPAIRS=8 python scratch/exp6.py "run$run" 2>&1 | grep -E "rotate (1|4|8) pairs"
  B200DCT_TMA_RUN=$run MODES=fwd NPAIRS=8 python scratch/exp6.py "run$run" 2>&1 | grep -E "rotate (1|4|8) pairs"
done
