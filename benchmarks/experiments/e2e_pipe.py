"""Where does the host pipeline lose 11 % against the raw PCIe bidirectional rate?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m
N = 8192; rows = 512; NS = 4
h_in = torch.randint(0, 256, (N, N), dtype=torch.int32).float().pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
din = [torch.empty(rows, N, device="cuda") for _ in range(NS)]; dout = [torch.empty(rows, N, device="cuda") for _ in range(NS)]
st = [torch.cuda.Stream() for _ in range(NS)]
plan = m.Plan()
def run(kernel, same_buf):
    for i, r0 in enumerate(range(0, N, rows)):
        s = st[i % NS]
        with torch.cuda.stream(s):
            din[i % NS].copy_(h_in[r0:r0 + rows], non_blocking=True)
            if kernel: m.roundtrip(din[i % NS], out=dout[i % NS], plan=plan, stream=s)
            src = din[i % NS] if same_buf else dout[i % NS]
            h_out[r0:r0 + rows].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
for name, k, sb in (("copies only (H2D -> D2H of the same buffer)", False, True), ("copies only (D2H of another buffer)", False, False), ("with the transform kernel", True, False)):
    run(k, sb); t0 = time.perf_counter()
    for _ in range(10): run(k, sb)
    dt = (time.perf_counter() - t0) / 10
    print(f"{name:48s} {dt*1e3:7.3f} ms  {N*N*4/dt/1e9:5.1f} GB/s each way", flush=True)
# three dedicated streams + events
s_h2d, s_k, s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
NS2 = 6
din2 = [torch.empty(rows, N, device="cuda") for _ in range(NS2)]; dout2 = [torch.empty(rows, N, device="cuda") for _ in range(NS2)]
def run3():
    ev_h = [None] * NS2; ev_k = [None] * NS2; ev_d = [None] * NS2
    for i, r0 in enumerate(range(0, N, rows)):
        b = i % NS2
        with torch.cuda.stream(s_h2d):
            if ev_k[b] is not None: s_h2d.wait_event(ev_k[b])      # din[b] consumed by its kernel
            din2[b].copy_(h_in[r0:r0 + rows], non_blocking=True); ev_h[b] = torch.cuda.Event(); ev_h[b].record(s_h2d)
        with torch.cuda.stream(s_k):
            s_k.wait_event(ev_h[b])
            if ev_d[b] is not None: s_k.wait_event(ev_d[b])        # dout[b] drained
            m.roundtrip(din2[b], out=dout2[b], plan=plan, stream=s_k); ev_k[b] = torch.cuda.Event(); ev_k[b].record(s_k)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_k[b])
            h_out[r0:r0 + rows].copy_(dout2[b], non_blocking=True); ev_d[b] = torch.cuda.Event(); ev_d[b].record(s_d2h)
    torch.cuda.synchronize()
run3(); t0 = time.perf_counter()
for _ in range(10): run3()
dt = (time.perf_counter() - t0) / 10
print(f"{'three engine streams + events, kernel':48s} {dt*1e3:7.3f} ms  {N*N*4/dt/1e9:5.1f} GB/s each way", flush=True)
