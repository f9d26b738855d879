"""Tiles per ticket (B200DCT_TMA_RUN, read at library load): 8192^2 f32 round trip rotating NP buffer pairs,
16384^2, forward-only, dense (12-warp flavour).  Round 1 saw single-tile tickets collapse to 113-118 us once
the working set exceeded ~1 GiB or more than 9 warps/SM were used (profiles/r01_tma_run_scheduler.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import cuda_dct_idct_b200 as m

def t(fn, iters=60):
    best = 1e9
    for rep in range(3):
        for i in range(5): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3

run = os.environ.get("B200DCT_TMA_RUN", "default")
N = 8192
k, n = np.mgrid[0:8, 0:8]
T = (np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8)) * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)
plan, dplan = m.Plan(), m.Plan(T=T, path=m.api.PATH_TMA)
for NP in (1, 2, 4, 8, 16):
    a = [torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float() for _ in range(NP)]
    b = [torch.empty_like(a[0]) for _ in range(NP)]
    r = t(lambda i: m.roundtrip(a[i % NP], out=b[i % NP], plan=plan))
    f = t(lambda i: m.forward(a[i % NP], coef=b[i % NP], plan=plan))
    d = t(lambda i: m.roundtrip(a[i % NP], out=b[i % NP], plan=dplan))
    print(f"RUN={run} 8192^2 x {NP:2d} pairs ({NP * 0.5:4.1f} GiB): round trip {r:7.2f} us  forward {f:7.2f} us  dense(12 warps) {d:7.2f} us", flush=True)
    del a, b
    torch.cuda.empty_cache()
a = torch.randint(0, 256, (16384, 16384), device="cuda", dtype=torch.int32).float(); b = torch.empty_like(a)
print(f"RUN={run} 16384^2: round trip {t(lambda i: m.roundtrip(a, out=b, plan=plan), 20):7.2f} us  dense {t(lambda i: m.roundtrip(a, out=b, plan=dplan), 20):7.2f} us", flush=True)
a8 = a.to(torch.uint8); b8 = torch.empty_like(a8)
print(f"RUN={run} 16384^2 u8 on the TMA family (16 warps): {t(lambda i: m.roundtrip(a8, out=b8, plan=m.Plan(path=m.api.PATH_TMA)), 20):7.2f} us", flush=True)
