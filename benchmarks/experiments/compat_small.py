"""The drop-in two-call path (dct_all_blocks_cuda then idct_all_blocks_cuda through libb200dct_compat.so)
at the README's small sizes: wall time per pair of calls (they are synchronous, like the reference's),
the device time the wrappers report, with and without the opt-in T cache; the unmodified reference
(oracle/_ref) beside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import cuda_dct_idct_b200 as m
import refgpu
from oracle import oracle as o

L = m.api.compat_lib()
L.b200dct_compat_set_options(1, 0)
T = torch.from_numpy(o.haweel_T()).cuda()
have_ref = refgpu.available("newappr")
if have_ref:
    refgpu.set_quant("newappr", o.jpeg_Q())
for N in (256, 512, 1024, 2048, 4096):
    img = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float()
    work, coef, rec = img.clone(), torch.empty_like(img), torch.empty_like(img)
    row = f"{N:5d}^2:"
    for cache in (0, 1):
        L.b200dct_compat_cache_transform(cache)
        reps = 200
        dev_ms = 0.0
        for i in range(reps + 20):
            if i == 20:
                torch.cuda.synchronize(); t0 = time.perf_counter(); dev_ms = 0.0
            m.dct_all_blocks_cuda(work, N, N, T, coef); dev_ms += L.b200dct_compat_last_ms()
            m.idct_all_blocks_cuda(coef, N, N, T, rec); dev_ms += L.b200dct_compat_last_ms()
        wall = (time.perf_counter() - t0) / reps
        row += f"  cache_T={cache}: wall {wall * 1e6:7.1f} us/pair, device {dev_ms / reps * 1e3:7.1f} us"
    if have_ref:
        ts = []
        for i in range(30):
            _, t1 = refgpu.dct("newappr", work, T, coef); _, t2 = refgpu.idct("newappr", coef, T, rec); ts.append(t1 + t2)
        t0 = time.perf_counter()
        for i in range(50):
            refgpu.dct("newappr", work, T, coef); refgpu.idct("newappr", coef, T, rec)
        wall = (time.perf_counter() - t0) / 50
        row += f"  | reference: wall {wall * 1e6:7.1f} us/pair, device {min(ts[5:]) * 1e3:7.1f} us"
    print(row, flush=True)
