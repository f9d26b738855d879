"""Early path of the direct family (loads + transform before griddepcontrol.wait, stores after): A/B with
B200DCT_EARLY_LOADS=0/1 (read at library load), rotating independent buffer pairs, back-to-back launches."""
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

tag = os.environ.get("B200DCT_EARLY_LOADS", "1")
k, n = np.mgrid[0:8, 0:8]
T = (np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8)) * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)
cases = [("u8 8192^2 default", 8192, torch.uint8, m.Plan()),
         ("u8 8192^2 exact inverse", 8192, torch.uint8, m.Plan(inverse=m.api.INVERSE_EXACT)),
         ("u8 8192^2 k=10", 8192, torch.uint8, m.Plan(keep=m.zigzag_mask(10))),
         ("u8 4096^2 default", 4096, torch.uint8, m.Plan()),
         ("f32 8192^2 direct", 8192, torch.float32, m.Plan(path=m.api.PATH_DIRECT)),
         ("f32 4096^2 (auto=direct)", 4096, torch.float32, m.Plan()),
         ("f32 2048^2 (auto=direct)", 2048, torch.float32, m.Plan()),
         ("f32 8192^2 dense chain direct", 8192, torch.float32, m.Plan(T=T, dense=m.api.DENSE_CHAIN)),
         ("f32 8192^2 tma (headline)", 8192, torch.float32, m.Plan())]
rgb_in = [torch.randint(0, 256, (8192, 8192, 3), device="cuda", dtype=torch.uint8) for _ in range(3)]
rgb_out = [torch.empty_like(x) for x in rgb_in]
rplan = m.Plan()
print(f"EARLY={tag} {'rgb 8192^2 (3 planes)':32s} {t(lambda i: m.roundtrip_rgb(rgb_in[i % 3], out=rgb_out[i % 3], plan=rplan), 20):8.2f} us  path={m.api.last_path()}", flush=True)
del rgb_in, rgb_out
torch.cuda.empty_cache()
for name, N, dt, plan in cases:
    NP = 6 if N >= 4096 else 40
    a = [torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).to(dt) for _ in range(NP)]
    b = [torch.empty_like(x) for x in a]
    us = t(lambda i: m.roundtrip(a[i % NP], out=b[i % NP], plan=plan))
    print(f"EARLY={tag} {name:32s} {us:8.2f} us  path={m.api.last_path()}", flush=True)
    del a, b
    torch.cuda.empty_cache()
