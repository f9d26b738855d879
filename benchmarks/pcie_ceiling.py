#!/usr/bin/env python
"""Copies-only host<->device ceiling of the box for the e2e leg, at 1..N ranks.

    python -m torch.distributed.run --nproc-per-node N benchmarks/pcie_ceiling.py [--mib 256]

Every rank (one per GPU, like bench.py) moves `mib` MiB host->device and, concurrently on a second
stream, `mib` MiB device->host between pinned buffers and its own GPU, `reps` times back to back;
no kernel runs.  Time = max over ranks, so the figure is what the BOX sustains when all N GPUs pull
at once (PCIe switches, root complexes and host memory are shared).  Rank 0 prints one JSON line;
the lines for N = 1, 2, 4, 8 are committed as profiles/r02_pcie_ceiling.json and bench.py reports
e2e.frac_of_pcie_ceiling against them."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_dct_idct_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, local_rank, world = m.dist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def timed(fns):
        for f in fns:
            f()
        torch.cuda.synchronize()
        m.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        m.dist.barrier()
        return m.dist.max_over_ranks(dt, dev)

    res = {"ranks": world, "mib_each_way_per_rank": args.mib, "reps": args.reps}
    for name, fns in (("h2d_alone", [h2d]), ("d2h_alone", [d2h]), ("duplex", [h2d, d2h])):
        dt = timed(fns)
        res[name + "_ms"] = dt * 1e3
        res[name + "_gb_s_each_way_per_rank"] = n / dt / 1e9
        res[name + "_gb_s_each_way_aggregate"] = n * world / dt / 1e9
    if rank == 0:
        print(json.dumps(res))
    m.dist.barrier()
    m.dist.shutdown()


if __name__ == "__main__":
    main()
