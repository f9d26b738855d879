"""C-loop timing (b200dct_time_calls) of both kernel families at small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
import os as _os
SIZES = [int(x) for x in _os.environ.get("SIZES", "64,128,256,512,1024,1536,2048,3072,4096").split(",")]
for N in SIZES:
    x = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float(); y = torch.empty_like(x)
    r = []
    for path in (2, 1):
        r.append(m.api.time_calls("roundtrip", x, y, plan=m.Plan(path=path), iters=300) * 1e3)
    x8 = x.to(torch.uint8); y8 = torch.empty_like(x8)
    r8 = [m.api.time_calls("roundtrip", x8, y8, plan=m.Plan(path=p), iters=300) * 1e3 for p in (2, 1)]
    print(f"N={N:5d} f32 tma {r[0]:7.2f} us direct {r[1]:7.2f} us | u8 tma {r8[0]:7.2f} us direct {r8[1]:7.2f} us", flush=True)
