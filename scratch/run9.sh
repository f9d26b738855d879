#!/bin/bash
python scratch/exp2.py w8
B200DCT_TMA_WARPS=6 python scratch/exp2.py w6
B200DCT_TMA_WARPS=4 B200DCT_TMA_GRID=296 python scratch/exp2.py 2x4
B200DCT_TMA_WARPS=3 B200DCT_TMA_GRID=444 python scratch/exp2.py 3x3
B200DCT_TMA_WARPS=2 B200DCT_TMA_GRID=592 python scratch/exp2.py 4x2
B200DCT_TMA_WARPS=4 B200DCT_TMA_GRID=296 B200DCT_TMA_L2PROMO=2 python scratch/exp2.py 2x4_promo128
B200DCT_TMA_STATIC=1 MODES=fwd,inv python scratch/exp2.py w8_static
PATHSEL=direct python scratch/exp2.py direct
