"""AUTO's family threshold for f32 round trips (direct below, TMA above): C-loop timings of both families,
back-to-back dependent launches on one buffer pair (b200dct_time_calls)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m

for N in (3072, 4096, 5120, 6144, 7168, 8192, 10240, 12288):
    a = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float()
    b = torch.empty_like(a)
    r = {}
    for name, path in (("direct", m.api.PATH_DIRECT), ("tma", m.api.PATH_TMA)):
        plan = m.Plan(path=path)
        m.api.time_calls("roundtrip", a, b, plan=plan, iters=20)
        r[name] = min(m.api.time_calls("roundtrip", a, b, plan=plan, iters=200) for _ in range(3)) * 1e3
    print(f"{N:6d}^2 ({N * N / 2**20:6.1f} Mpixel): direct {r['direct']:7.2f} us   tma {r['tma']:7.2f} us   {'direct' if r['direct'] < r['tma'] else 'tma'}", flush=True)
