#!/usr/bin/env python
"""BASELINE.json configs[4]: a 32768 x 32768 uint8 image (or a batch of 64 8192^2 images --
the same bytes, a batch is one tall image) striped by block-rows over the N GPUs of one
box.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29541 benchmarks/stripes.py [--dtype u8|f32] [--rows 32768] [--gather]

Every rank materialises its own stripe on its GPU with a counter-based generator (no H2D),
transforms it with the fused kernel (no halo, no collective), and the slowest rank's CUDA
event time defines the step.  --gather additionally times the OPTIONAL final NCCL
all-gather of the finished stripes, reported separately (it is NVLink-bound, not part of
the transform).  Parity: each rank checks two bands of its stripe against the CPU oracle
on the same generated pixels (the oracle is only the checker).  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cuda_dct_idct_b200 as m  # noqa: E402


def splitmix_u8_device(n_rows, W, row0, seed, dev):
    """v = splitmix64(seed + linear pixel index) & 255, identical to tests/inputs.splitmix_u8."""
    out = torch.empty((n_rows, W), dtype=torch.uint8, device=dev)
    chunk = max(8, (1 << 26) // W)
    for r in range(0, n_rows, chunk):
        h = min(chunk, n_rows - r)
        idx = torch.arange((row0 + r) * W, (row0 + r + h) * W, dtype=torch.int64, device=dev) + seed
        z = idx * -7046029254386353131                      # 0x9E3779B97F4A7C15 as int64 (wraps mod 2^64)
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * -4658895280553007687   # 0xBF58476D1CE4E5B9
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * -7723592293110705685   # 0x94D049BB133111EB
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        out[r:r + h] = (z & 255).to(torch.uint8).view(h, W)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--cols", type=int, default=32768)
    ap.add_argument("--dtype", default="u8", choices=["u8", "f32"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gather", action="store_true", help="time the optional NCCL all-gather of the stripes")
    ap.add_argument("--inverse", default="auto", choices=["auto", "exact"],
                    help="auto: the library default (factored +-1 LSB inverse for u8 output); exact: the reference's chains")
    ap.add_argument("--fused-gather", action="store_true",
                    help="transform straight into rank 0's image over NVLink peer stores (dist.PeerImage)")
    args = ap.parse_args()

    rank, local_rank, world = m.dist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    H, W = args.rows, args.cols
    r0, r1 = m.stripe_rows(H, world, rank)
    img = splitmix_u8_device(r1 - r0, W, r0, 42, dev)
    if args.dtype == "f32":
        img = img.float()
    out = torch.empty_like(img)
    plan = m.Plan(inverse=m.api.INVERSE_EXACT if args.inverse == "exact" else m.api.INVERSE_AUTO)
    exact = args.dtype == "f32" or args.inverse == "exact"     # else: u8 pixels within 1 LSB of the oracle

    def same(got, ref):
        if args.dtype == "f32":
            return np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        return np.array_equal(got, ref) if exact else int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1

    for _ in range(args.warmup):
        m.roundtrip(img, out=out, plan=plan)
    torch.cuda.synchronize()
    m.dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m.roundtrip(img, out=out, plan=plan)
    e1.record()
    torch.cuda.synchronize()
    m.dist.barrier()
    ms = m.dist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)

    # parity of what was just computed: first and last 16 rows of this stripe vs the oracle
    from oracle import oracle as o
    import inputs

    ok = True
    for a in (0, (r1 - r0) - 16):
        band = inputs.splitmix_u8(16 * W, 42, (r0 + a) * W).reshape(16, W)
        ref = o.roundtrip(band if args.dtype == "u8" else band.astype(np.float32))
        got = out[a:a + 16].cpu().numpy()
        ok = ok and same(got, ref)
    ok_all = m.dist.sum_over_ranks(0.0 if ok else 1.0, dev) == 0.0

    gather_ms = None
    if args.gather and world > 1:
        for _ in range(2):
            m.dist.gather_stripes(out, H)
        torch.cuda.synchronize()
        m.dist.barrier()
        e0.record()
        full = m.dist.gather_stripes(out, H)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = m.dist.max_over_ranks(e0.elapsed_time(e1), dev)
        del full
    fused_ms, fused_ok = None, None
    if args.fused_gather and world > 1:
        peer = m.dist.PeerImage(H, W, img.dtype, dev)
        dst = peer.stripe_on(0, r0, r1)               # rows [r0, r1) of rank 0's full image
        for _ in range(3):
            m.roundtrip(img, out=dst, plan=plan)
            peer.barrier()
        torch.cuda.synchronize()
        m.dist.barrier()
        e0.record()
        for _ in range(args.steps):
            m.roundtrip(img, out=dst, plan=plan)
            peer.barrier()
        e1.record()
        torch.cuda.synchronize()
        m.dist.barrier()
        fused_ms = m.dist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
        if rank == 0:      # the assembled image: bands across every stripe boundary vs the oracle
            full = peer.local()
            fused_ok = True
            for r in range(world):
                a0, a1 = m.stripe_rows(H, world, r)
                for b0 in (a0, a1 - 16):
                    band = inputs.splitmix_u8(16 * W, 42, b0 * W).reshape(16, W)
                    ref = o.roundtrip(band if args.dtype == "u8" else band.astype(np.float32))
                    got = full[b0:b0 + 16].cpu().numpy()
                    fused_ok = fused_ok and same(got, ref)
        m.dist.barrier()
    es = 1 if args.dtype == "u8" else 4
    if rank == 0:
        print(json.dumps({
            "workload": f"{H}x{W} {args.dtype} image striped by block-rows over {world} GPU(s) (= {H * W // (8192 * 8192)} images of 8192^2)",
            "n_gpus": world, "ms_per_step": ms, "gpixel_s": H * W / ms / 1e6, "gb_s_per_gpu": 2 * es * H * W / world / ms / 1e6,
            "rows_per_gpu": r1 - r0, "kernel_path": m.api.last_path(), "parity_vs_oracle": bool(ok_all),
            "parity_criterion": "bit-exact" if exact else "u8 within 1 LSB (factored inverse, the library default for 8-bit output)",
            "optional_gather_ms": gather_ms, "fused_transform_plus_gather_ms": fused_ms,
            "fused_gather_parity": fused_ok, "steps": args.steps, "collective_on_data_path": "none"}))
    m.dist.barrier()
    m.dist.shutdown()


if __name__ == "__main__":
    main()
