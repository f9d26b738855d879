"""Times forward / fused round trip with the compact coefficient outputs (f32 plane, i16 plane, i16 zig-zag stream)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
N = int(os.environ.get("N", 8192)); iters = int(os.environ.get("ITERS", 200))
dev = torch.device("cuda")
def t(fn):
    for i in range(5): fn(i)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / iters)
    return best
for name, dt in (("f32", torch.float32), ("u8", torch.uint8)):
    ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).to(dt) for _ in range(4)]
    outs = [torch.empty_like(x) for x in ins]
    cf32 = [torch.empty(N, N, device=dev) for _ in range(2)]
    ci16 = [torch.empty(N, N, device=dev, dtype=torch.int16) for _ in range(2)]
    czz = [m.api.empty_zigzag(N, N, dev) for _ in range(2)]
    direct = m.Plan(path=1)
    rows = [("forward -> f32 plane (auto)", lambda i: m.forward(ins[i % 4], coef=cf32[i % 2])),
            ("forward -> i16 plane (auto)", lambda i: m.forward(ins[i % 4], coef=ci16[i % 2])),
            ("forward -> i16 plane (direct)", lambda i: m.forward(ins[i % 4], coef=ci16[i % 2], plan=direct)),
            ("forward -> i16 zig-zag stream", lambda i: m.forward(ins[i % 4], coef=czz[i % 2], zigzag=True)),
            ("inverse <- i16 plane (direct)", lambda i: m.inverse(ci16[i % 2], img=outs[i % 4], plan=direct)),
            ("inverse <- i16 zig-zag stream", lambda i: m.inverse(czz[i % 2], img=outs[i % 4], zigzag=True)),
            ("round trip + i16 plane (auto)", lambda i: m.roundtrip(ins[i % 4], out=outs[i % 4], coef=ci16[i % 2])),
            ("round trip + i16 zig-zag stream", lambda i: m.roundtrip(ins[i % 4], out=outs[i % 4], coef=czz[i % 2], zigzag=True))]
    for label, fn in rows:
        ms = t(fn); print(f"[{name} N={N}] {label:34s} {m.api.last_path():6s} {ms*1e3:8.1f} us  {N*N/ms/1e6:8.1f} Gpx/s", flush=True)
