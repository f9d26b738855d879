"""A/B timing of one library build: python benchmarks/experiments/exp_time.py [tag]   (B200DCT_LIB_DIR selects the build)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("B200DCT_LIB_DIR", "default")
N = int(os.environ.get("N", 8192)); iters = int(os.environ.get("ITERS", 300))
dev = torch.device("cuda")
def bench(dtype, path, coef=False):
    plan = m.Plan(path=path)
    if dtype == torch.float32:
        ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(4)]
    else:
        ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.uint8) for _ in range(4)]
    outs = [torch.empty_like(x) for x in ins]
    cf = torch.empty(N, N, device=dev) if coef else None
    try:
        for i in range(5): m.roundtrip(ins[i % 4], out=outs[i % 4], coef=cf, plan=plan)
    except m.B200DCTError as e:
        return None
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): m.roundtrip(ins[i % 4], out=outs[i % 4], coef=cf, plan=plan)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best
for name, dt, bpp in (("f32", torch.float32, 8), ("u8", torch.uint8, 2)):
    for pname, p in (("tma", 2), ("direct", 1)):
        ms = bench(dt, p)
        if ms is None: print(f"[{tag}] {name} {pname}: unavailable"); continue
        print(f"[{tag}] {name:3s} {pname:6s} N={N}: {ms*1e3:8.1f} us  {N*N/ms/1e6:8.1f} Gpx/s  {bpp*N*N/ms/1e6:8.1f} GB/s", flush=True)
