"""Direct family: same pixel count (64 Mpixel), different aspect ratios -- does the order in which
CTAs walk the image (grid.x = groups of 4 block-rows, fastest; grid.y = groups of 32 block-columns)
cost DRAM locality on wide images?  Also the TMA family for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m

def t(fn, iters=40):
    best = 1e9
    for rep in range(3):
        for i in range(5): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3

for dt in (torch.float32, torch.uint8):
    for H, W in ((65536, 1024), (32768, 2048), (16384, 4096), (8192, 8192), (4096, 16384), (2048, 32768), (1024, 65536)):
        a = [torch.randint(0, 256, (H, W), device="cuda", dtype=torch.int32).to(dt) for _ in range(4)]
        b = [torch.empty_like(x) for x in a]
        res = []
        for path in (m.api.PATH_DIRECT, m.api.PATH_TMA):
            plan = m.Plan(path=path)
            res.append(t(lambda i: m.roundtrip(a[i % 4], out=b[i % 4], plan=plan)))
        print(f"{str(dt):14s} {H:6d} x {W:6d}: direct {res[0]:7.2f} us   tma {res[1]:7.2f} us", flush=True)
        del a, b
        torch.cuda.empty_cache()
