"""python benchmarks/experiments/exp2.py <tag> : times fwd / inv / rt / dense-rt on f32 8192^2 for the env-selected config"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import cuda_dct_idct_b200 as m
tag = sys.argv[1]; N = int(os.environ.get("N", 8192)); iters = int(os.environ.get("ITERS", 200))
modes = os.environ.get("MODES", "rt,fwd,inv,dense").split(",")
path = {"tma": 2, "direct": 1, "auto": 0}[os.environ.get("PATHSEL", "tma")]
dev = torch.device("cuda")
k = np.arange(8)[:, None]; n = np.arange(8)[None, :]
T2 = (np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8)) * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)
ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(4)]
outs = [torch.empty_like(x) for x in ins]
plan = m.Plan(path=path); dplan = m.Plan(T=T2, path=path)
coefs = [m.forward(x, plan=plan) for x in ins]
def t(fn):
    for i in range(5): fn(i)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / iters)
    return best
fns = {"rt": lambda i: m.roundtrip(ins[i % 4], out=outs[i % 4], plan=plan),
       "fwd": lambda i: m.forward(ins[i % 4], coef=outs[i % 4], plan=plan),
       "inv": lambda i: m.inverse(coefs[i % 4], img=outs[i % 4], plan=plan),
       "dense": lambda i: m.roundtrip(ins[i % 4], out=outs[i % 4], plan=dplan)}
for md in modes:
    ms = t(fns[md]); print(f"[{tag}] {md:5s} {m.api.last_path():6s} {ms*1e3:8.1f} us {8*N*N/ms/1e6:8.1f} GB/s", flush=True)
