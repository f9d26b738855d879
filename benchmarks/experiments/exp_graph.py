"""Captured launches: a CUDA graph of 8 fused round trips (4 rotating 8192^2 f32 buffer pairs), replayed; us per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
N = int(os.environ.get("N", 8192)); dev = torch.device("cuda")
ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(4)]
outs = [torch.empty_like(x) for x in ins]
for name, path in (("auto", 0), ("direct", 1), ("tma", 2)):
    plan = m.Plan(path=path)
    for i in range(4): m.roundtrip(ins[i], out=outs[i], plan=plan)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        for i in range(8): m.roundtrip(ins[i % 4], out=outs[i % 4], plan=plan, stream=torch.cuda.current_stream())
        used = m.api.last_path()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(25): g.replay()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 200)
    print(f"[graph replay N={N}] plan={name:6s} captured path={used:6s} {best*1e3:8.1f} us per launch", flush=True)
