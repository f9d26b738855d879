"""footprint diagnostics: env MODES=fwd|rt PATHSEL=tma|direct N NPAIRS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
tag = sys.argv[1]; N = int(os.environ.get("N", 8192)); md = os.environ.get("MODES", "fwd"); NB = int(os.environ.get("NPAIRS", 8))
path = {"tma": 2, "direct": 1}[os.environ.get("PATHSEL", "tma")]
dev = torch.device("cuda"); plan = m.Plan(path=path)
ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(NB)]
outs = [torch.empty(N, N, device=dev) for _ in range(NB)]
fn = (lambda i, j: m.forward(ins[i], coef=outs[j], plan=plan)) if md == "fwd" else (lambda i, j: m.roundtrip(ins[i], out=outs[j], plan=plan))
K = 128
def run(name, sel):
    for i in range(2 * NB): fn(*sel(i))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): fn(*sel(i))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"[{tag}] N={N} {md} {m.api.last_path()} {name:18s} {ms * 1e3:7.1f} us {8*N*N/ms/1e6:7.0f} GB/s", flush=True)
for k in (1, 2, 4, 8, 16, 32):
    if k <= NB: run(f"rotate {k} pairs", lambda i, k=k: (i % k, i % k))
