"""PCIe ceilings for the e2e leg: H2D alone, D2H alone, both at once (256 MiB pinned buffers)."""
import torch, time
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H concurrently", both)):
    dt = t(fn); print(f"{name:26s} {dt*1e3:7.3f} ms  {n/dt/1e9:6.1f} GB/s per direction", flush=True)
for chunk_mb in (4, 16, 64):
    c = chunk_mb << 20
    def chunked():
        for o in range(0, n, c):
            with torch.cuda.stream(s1): d_in[o:o+c].copy_(h_in[o:o+c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o+c].copy_(d_out[o:o+c], non_blocking=True)
    dt = t(chunked, 5); print(f"both, {chunk_mb:3d} MiB chunks        {dt*1e3:7.3f} ms  {n/dt/1e9:6.1f} GB/s per direction", flush=True)
