import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
from cuda_dct_idct_b200 import api
N = 8192; dev = torch.device("cuda"); L = m.lib(); plan = m.Plan()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for dt, code, es in ((torch.float32, 0, 4), (torch.uint8, 1, 1)):
    x = torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32)
    x = x.float() if dt == torch.float32 else x.to(torch.uint8)
    out = torch.empty_like(x)
    nb = int(L.b200dct_metrics_workspace_bytes(N, N)); ws = torch.empty(nb // 8, dtype=torch.float64, device=dev)
    acc = torch.zeros(3, dtype=torch.float64, device=dev); acc2 = torch.zeros(2, dtype=torch.float64, device=dev)
    def fused(): api._check(L.b200dct_roundtrip_metrics(plan._h, x.data_ptr(), code, N * es, out.data_ptr(), code, N * es, None, 0, 0, N, N, acc.data_ptr(), ws.data_ptr(), nb, s))
    def sep():
        api._check(L.b200dct_roundtrip(plan._h, x.data_ptr(), code, N * es, out.data_ptr(), code, N * es, None, 0, 0, N, N, s))
        api._check(L.b200dct_metrics_accumulate(x.data_ptr(), out.data_ptr(), code, N * es, N, N, acc2.data_ptr(), s))
    for name, fn in (("fused", fused), ("separate", sep)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{str(dt):14s} {name:9s} {e0.elapsed_time(e1) / 100 * 1e3:7.1f} us", flush=True)
