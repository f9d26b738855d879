"""b200dct_roundtrip_any on ragged / unaligned images vs the aligned fast path (us per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
dev = torch.device("cuda"); iters = int(os.environ.get("ITERS", 100))
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / iters)
    return best
for dt in (torch.float32, torch.uint8):
    for H, W in ((8192, 8192), (8191, 8191), (8185, 8187), (4096, 4099), (1081, 1923)):
        x = torch.randint(0, 256, (H, W), device=dev, dtype=torch.int32).to(dt); o = torch.empty_like(x)
        ms = t(lambda: m.roundtrip_any(x, out=o))
        print(f"[any] {str(dt):14s} {H}x{W}: {m.api.last_path():6s} {ms*1e3:8.1f} us  {H*W/ms/1e6:8.1f} Gpx/s", flush=True)
