#!/usr/bin/env python
"""Size / variant / mask sweep (BASELINE.json configs[2..4]) on ONE B200.

    python benchmarks/sweep.py [--out gpurun_out/sweep] [--quick]

For every README size 256^2..8192^2 (and 16384^2 for the dense "exact DCT" variant) times,
with CUDA events over back-to-back launches rotating 4 buffer pairs:
  new/fused      b200dct_roundtrip, f32 and u8 (one launch)
  new/split      b200dct_forward + b200dct_inverse (the drop-in two-call API, two launches)
  new/mask k     fused, retained-coefficient masks k = 6..10 (Q_PARAM kernels)
  new/dense      fused with a true DCT-II matrix as T (the cublasDCT* replacement)
  ref/*          the UNMODIFIED reference kernels from oracle/_ref (HpApprDCT, fastApprDCT,
                 cublasDCTv2, cublasDCT) -- their own printed cudaEvent times (DCT + IDCT)
Sizes whose in+out planes fit the 126 MB L2 (<= 2048^2 f32) are flagged cache-resident.
Writes <out>.json and <out>.md.  The reference libraries are baselines beside the number,
never on the product path.
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


def dct2_matrix():
    k = np.arange(8)[:, None]
    n = np.arange(8)[None, :]
    c = np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8))
    return (c * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)


def time_ms(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda")
    sizes = [256, 512, 1024, 2048, 4096, 8192]
    rows = []
    plans = {"default": m.Plan(), "dense": m.Plan(T=dct2_matrix())}
    for k in (6, 7, 8, 9, 10):
        plans[f"k{k}"] = m.Plan(keep=m.zigzag_mask(k))

    try:
        import refgpu
        from oracle import oracle as o

        have_ref = {v: refgpu.available(v) for v in refgpu.VARIANTS}
    except Exception:
        have_ref = {}

    def add(N, name, ms, bytes_per_px, launches, note=""):
        px = N * N
        rows.append({"N": N, "variant": name, "ms": ms, "gpixel_s": px / ms / 1e6, "gb_s": bytes_per_px * px / ms / 1e6,
                     "bytes_per_px": bytes_per_px, "launches": launches, "l2_resident": bool(2 * px * 4 <= 126e6), "note": note})
        print(f"N={N:6d} {name:22s} {ms * 1e3:10.1f} us {px / ms / 1e6:9.1f} Gpx/s {bytes_per_px * px / ms / 1e6:9.1f} GB/s {note}", flush=True)

    for N in sizes + ([] if args.quick else [16384]):
        iters = max(20, min(400, int(3e9 / (N * N))))
        nb = 4 if N <= 8192 else 2
        f32 = [torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float() for _ in range(nb)]
        out = [torch.empty_like(x) for x in f32]
        if N <= 8192:
            add(N, "new/fused f32", time_ms(lambda i: m.roundtrip(f32[i % nb], out=out[i % nb], plan=plans["default"]), iters), 8, 1, m.api.last_path())
            # same call looped from C (b200dct_time_calls): no Python between launches -- this is
            # the number to read at the small, launch-latency-bound sizes
            add(N, "new/fused f32 (C loop)", m.api.time_calls("roundtrip", f32[0], out[0], plan=plans["default"], iters=iters), 8, 1, "one buffer pair")
            c0 = torch.empty(N, N, device=dev)
            add(N, "new/split f32 (C loop)", m.api.time_calls("split", f32[0], out[0], c0, plan=plans["default"], iters=iters), 16, 2, "one buffer pair")
            del c0
            coef = torch.empty(N, N, device=dev)
            add(N, "new/split f32", time_ms(lambda i: (m.forward(f32[i % nb], coef=coef, plan=plans["default"]),
                                                         m.inverse(coef, img=out[i % nb], plan=plans["default"])), iters), 16, 2, m.api.last_path())
            add(N, "new/fused+coef f32", time_ms(lambda i: m.roundtrip(f32[i % nb], out=out[i % nb], coef=coef, plan=plans["default"]), iters), 12, 1)
            c16 = torch.empty(N, N, device=dev, dtype=torch.int16)
            add(N, "new/fused+coef i16", time_ms(lambda i: m.roundtrip(f32[i % nb], out=out[i % nb], coef=c16, plan=plans["default"]), iters), 10, 1)
            del coef, c16
            u8 = [x.to(torch.uint8) for x in f32]
            o8 = [torch.empty_like(x) for x in u8]
            add(N, "new/fused u8", time_ms(lambda i: m.roundtrip(u8[i % nb], out=o8[i % nb], plan=plans["default"]), iters), 2, 1, m.api.last_path())
            if N in (2048, 8192):
                for k in (6, 7, 8, 9, 10):
                    add(N, f"new/mask k={k} f32", time_ms(lambda i: m.roundtrip(f32[i % nb], out=out[i % nb], plan=plans[f"k{k}"]), iters), 8, 1, m.api.last_path())
                for k in (6, 7, 8, 9, 10):
                    add(N, f"new/mask k={k} u8", time_ms(lambda i: m.roundtrip(u8[i % nb], out=o8[i % nb], plan=plans[f"k{k}"]), iters), 2, 1, m.api.last_path())
            del u8, o8
        if have_ref.get("newappr") is not None and N <= 2048:
            # config 1 / README "DCT on CPU (Sequential)": the oracle, one thread, same generator
            try:
                cpu_img = o.rand_image(N, N, 42)
                add(N, "cpu/oracle 1 thread", o.time_roundtrip(cpu_img, reps=3, threads=1) * 1e3, 8, 0, "sequential C restatement, srand(42) image")
            except Exception as e:  # pragma: no cover
                print("cpu baseline skipped:", e)
        add(N, "new/dense(DCT-II) f32", time_ms(lambda i: m.roundtrip(f32[i % nb], out=out[i % nb], plan=plans["dense"]), iters), 8, 1, m.api.last_path())
        # ---- reference kernels, their own event timers
        if have_ref.get("newappr") and N <= 8192:
            T = torch.from_numpy(o.haweel_T()).cuda()
            img, coef = f32[0].clone(), torch.empty(N, N, device=dev)
            for variant, label in (("newappr", "ref/HpApprDCT"), ("fastappr", "ref/fastApprDCT"), ("cublas2", "ref/cublasDCTv2"), ("cublas", "ref/cublasDCT")):
                if not have_ref.get(variant):
                    continue
                if variant == "cublas" and N > 1024:
                    continue          # (N/8)^2 * 4 cuBLAS launches per round trip: minutes at 8192^2
                if variant == "cublas2" and N > 8192:
                    continue
                refgpu.set_quant(variant, o.jpeg_Q())
                ts = []
                for _ in range(3 if N >= 4096 else 5):
                    img.copy_(f32[0])
                    _, t1 = refgpu.dct(variant, img, T, coef)
                    _, t2 = refgpu.idct(variant, coef, T, out[0])
                    ts.append(t1 + t2)
                add(N, label, min(ts[1:]), 48, 6 if "Appr" in label else -1, "own cudaEvent times, DCT+IDCT")
            del img, coef
        del f32, out
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".json", "w") as f:
        json.dump(rows, f, indent=1)
    variants = []
    for r in rows:
        if r["variant"] not in variants:
            variants.append(r["variant"])
    Ns = sorted({r["N"] for r in rows})
    with open(args.out + ".md", "w") as f:
        f.write("| variant | " + " | ".join(f"{n}² ms" for n in Ns) + " | Gpx/s @ largest |\n|---|" + "---|" * (len(Ns) + 1) + "\n")
        for v in variants:
            cells, last = [], ""
            for n in Ns:
                hit = [r for r in rows if r["variant"] == v and r["N"] == n]
                cells.append(f"{hit[0]['ms']:.4f}" if hit else "")
                if hit:
                    last = f"{hit[0]['gpixel_s']:.1f}"
            f.write(f"| {v} | " + " | ".join(cells) + f" | {last} |\n")
        f.write("\nSizes <= 2048² f32 (in+out <= 126 MB) are L2-resident across iterations.\n")
    print("wrote", args.out + ".json", args.out + ".md")


if __name__ == "__main__":
    main()
