import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
N = 8192
for dt in (torch.float32, torch.uint8):
    h_in = (torch.randint(0, 256, (N, N), dtype=torch.int32).float() if dt == torch.float32 else torch.randint(0, 256, (N, N), dtype=torch.uint8)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    for _ in range(3): m.roundtrip_host(h_in, h_out)
    t0 = time.perf_counter()
    for _ in range(10): m.roundtrip_host(h_in, h_out)
    dt_s = (time.perf_counter() - t0) / 10
    b = h_in.numel() * h_in.element_size()
    print(f"chunk={os.environ.get('B200DCT_HOST_CHUNK_MB','def')} {str(dt):14s} {dt_s*1e3:7.3f} ms  {N*N/dt_s/1e9:6.2f} Gpx/s  {b/dt_s/1e9:5.1f} GB/s each way", flush=True)
