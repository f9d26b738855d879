"""pitch experiment: python benchmarks/experiments/exp3.py <tag>   env PAD (floats of row padding), MODES, N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
tag = sys.argv[1]; N = int(os.environ.get("N", 8192)); iters = int(os.environ.get("ITERS", 200)); PAD = int(os.environ.get("PAD", 0))
modes = os.environ.get("MODES", "rt,fwd").split(",")
L = int(os.environ.get("LAUNCHES", 0))
dev = torch.device("cuda"); plan = m.Plan(path=2)
def mk(): 
    t = torch.randint(0, 256, (N, N + PAD), device=dev, dtype=torch.int32).float()
    return t[:, :N]
ins = [mk() for _ in range(4)]; outs = [mk() for _ in range(4)]
fns = {"rt": lambda i: m.roundtrip(ins[i % 4], out=outs[i % 4], plan=plan),
       "fwd": lambda i: m.forward(ins[i % 4], coef=outs[i % 4], plan=plan)}
if L:
    for i in range(L): fns[modes[0]](i)
    torch.cuda.synchronize(); print("ok"); sys.exit(0)
for md in modes:
    fn = fns[md]
    for i in range(5): fn(i)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / iters)
    print(f"[{tag}] pad={PAD:5d} {md:4s} {best*1e3:8.1f} us {8*N*N/best/1e6:8.1f} GB/s", flush=True)
