"""rotation diagnostics.  env MODES=fwd|rt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
tag = sys.argv[1]; N = 8192; md = os.environ.get("MODES", "fwd"); NB = 8
dev = torch.device("cuda"); plan = m.Plan(path=2)
ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(NB)]
outs = [torch.empty(N, N, device=dev) for _ in range(NB)]
fn = (lambda i, j: m.forward(ins[i], coef=outs[j], plan=plan)) if md == "fwd" else (lambda i, j: m.roundtrip(ins[i], out=outs[j], plan=plan))
K = 96
def run(name, sel):
    for i in range(8): fn(*sel(i))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): fn(*sel(i))
    e1.record(); torch.cuda.synchronize()
    print(f"[{tag}] {md} {name:28s} {e0.elapsed_time(e1) / K * 1e3:7.1f} us", flush=True)
run("same pair", lambda i: (0, 0))
run("rotate 2 pairs", lambda i: (i % 2, i % 2))
run("rotate 4 pairs", lambda i: (i % 4, i % 4))
run("rotate 8 pairs", lambda i: (i % 8, i % 8))
run("rotate in only (4)", lambda i: (i % 4, 0))
run("rotate out only (4)", lambda i: (0, i % 4))
run("rotate in 2 / out 2", lambda i: (i % 2, (i // 2) % 2))
