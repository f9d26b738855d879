"""Tiny driver for ncu: python scratch/prof_rt.py <tma|direct> <f32|u8> [N] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
path = {"tma": 2, "direct": 1}[sys.argv[1]]; dt = sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8192; L = int(sys.argv[4]) if len(sys.argv) > 4 else 6
x = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32)
x = x.float() if dt == "f32" else x.to(torch.uint8)
ins = [x.clone() for _ in range(2)]; outs = [torch.empty_like(x) for _ in range(2)]
plan = m.Plan(path=path)
for i in range(L): m.roundtrip(ins[i % 2], out=outs[i % 2], plan=plan)
torch.cuda.synchronize(); print("ok", m.api.last_path())
