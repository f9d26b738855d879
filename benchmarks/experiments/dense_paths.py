"""Dense (exact DCT-II) round trip: direct vs TMA family, symmetric vs chain kernels, 8192^2 and 16384^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import cuda_dct_idct_b200 as m
k, n = np.mgrid[0:8, 0:8]
T = (np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8)) * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)
for N in (8192, 16384):
    nb = 4 if N == 8192 else (2 if N < 16384 else 1)
    a = [torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float() for _ in range(nb)]
    b = [torch.empty_like(a[0]) for _ in range(nb)]
    for dense, dn in ((m.api.DENSE_AUTO, "sym"), (m.api.DENSE_CHAIN, "chain"), (m.api.DENSE_MMA, "mma")):
        for path, pn in ((m.api.PATH_DIRECT, "direct"), (m.api.PATH_TMA, "tma")):
            if dn == "mma" and pn == "tma":
                continue
            plan = m.Plan(T=T, dense=dense, path=path)
            best = 1e9
            for rep in range(3):
                for i in range(3): m.roundtrip(a[i % nb], out=b[i % nb], plan=plan)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(40): m.roundtrip(a[i % nb], out=b[i % nb], plan=plan)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 40)
            print(f"dense {dn:5s} {pn:6s} {N}^2: {best * 1e3:8.1f} us  {8 * N * N / best / 1e6:8.1f} GB/s", flush=True)
