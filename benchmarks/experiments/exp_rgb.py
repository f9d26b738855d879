"""b200dct_roundtrip_rgb at 8192^2 (us per call); B200DCT_LIB_DIR selects a variant build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
N = int(os.environ.get("N", 8192))
a = [torch.randint(0, 256, (N, N, 3), device="cuda", dtype=torch.uint8) for _ in range(2)]
b = [torch.empty_like(a[0]) for _ in range(2)]
for name, plan in (("default (factored inverse)", m.Plan()), ("exact inverse", m.Plan(inverse=m.api.INVERSE_EXACT)), ("k=10 runtime mask", m.Plan(keep=m.zigzag_mask(10)))):
    best = 1e9
    for rep in range(3):
        for i in range(3): m.roundtrip_rgb(a[i % 2], out=b[i % 2], plan=plan)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30): m.roundtrip_rgb(a[i % 2], out=b[i % 2], plan=plan)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(f"rgb {N}^2 {name:28s} {best * 1e3:8.1f} us   {N * N / best / 1e6:7.1f} Gpixel/s (x3 planes)  lib={os.environ.get('B200DCT_LIB_DIR', 'default')}", flush=True)
