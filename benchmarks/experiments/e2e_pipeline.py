"""e2e A/B: synchronous b200dct_roundtrip_host vs the image-overlapping HostPipeline, per chunk size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m

N = 8192
steps = int(os.environ.get("STEPS", "20"))
for name, dtype in (("f32", torch.float32), ("u8", torch.uint8)):
    h_in = torch.randint(0, 256, (N, N), dtype=torch.int32).to(dtype).pin_memory()
    h_outs = [torch.empty(N, N, dtype=dtype).pin_memory() for _ in range(2)]
    plan = m.Plan()
    for i in range(2):
        m.roundtrip_host(h_in, h_outs[i % 2], plan=plan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        m.roundtrip_host(h_in, h_outs[i % 2], plan=plan)
    torch.cuda.synchronize()
    print(f"{name} sync        {(time.perf_counter() - t0) / steps * 1e3:8.3f} ms/step", flush=True)
    for chunk_mb in (4, 8, 16, 32, 64):
        for slots in (3, 4, 6):
            t0 = time.perf_counter()
            pipe = m.HostPipeline(plan=plan, chunk_bytes=chunk_mb << 20, slots=slots)
            t_create = time.perf_counter() - t0
            for i in range(2):
                pipe.submit(h_in, h_outs[i % 2])
            pipe.drain()
            t0 = time.perf_counter()
            for i in range(steps):
                pipe.submit(h_in, h_outs[i % 2])
            t_submit = time.perf_counter() - t0
            pipe.drain()
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            pipe.close()
            t_close = time.perf_counter() - t0
            print(f"{name} pipe chunk {chunk_mb:3d} MiB slots {slots}: {dt / steps * 1e3:8.3f} ms/step  (submit returned after {t_submit * 1e3:7.2f} ms; "
                  f"create {t_create * 1e3:6.2f} ms, close {t_close * 1e3:6.2f} ms)", flush=True)
