"""per-launch timing: is the collapse tied to particular buffer pairs?  env MODES=fwd|rt, NBUF"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, statistics
import cuda_dct_idct_b200 as m
tag = sys.argv[1]; N = 8192; NBUF = int(os.environ.get("NBUF", 4)); md = os.environ.get("MODES", "fwd")
dev = torch.device("cuda"); plan = m.Plan(path=2)
ins = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(NBUF)]
outs = [torch.empty(N, N, device=dev) for _ in range(NBUF)]
print("in ptrs ", [hex(x.data_ptr()) for x in ins]); print("out ptrs", [hex(x.data_ptr()) for x in outs])
fn = (lambda i, j: m.forward(ins[i], coef=outs[j], plan=plan)) if md == "fwd" else (lambda i, j: m.roundtrip(ins[i], out=outs[j], plan=plan))
for i in range(8): fn(i % NBUF, i % NBUF)
torch.cuda.synchronize()
K = 64
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
evs[0].record()
for i in range(K):
    fn(i % NBUF, i % NBUF); evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(K)]
for b in range(NBUF):
    sel = ts[b::NBUF]; print(f"[{tag}] {md} pair {b}: median {statistics.median(sel):7.1f} us  min {min(sel):7.1f} max {max(sel):7.1f}")
# same pair repeatedly
for b in range(min(NBUF, 4)):
    evs[0].record()
    for i in range(K): fn(b, b)
    evs[1].record(); torch.cuda.synchronize()
    print(f"[{tag}] {md} pair {b} repeated: {evs[0].elapsed_time(evs[1]) / K * 1e3:7.1f} us")
# cross pairs: in b, out b+1
for b in range(min(NBUF, 4)):
    evs[0].record()
    for i in range(K): fn(b, (b + 1) % NBUF)
    evs[1].record(); torch.cuda.synchronize()
    print(f"[{tag}] {md} in {b} -> out {(b+1)%NBUF} repeated: {evs[0].elapsed_time(evs[1]) / K * 1e3:7.1f} us")
