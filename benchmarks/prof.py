#!/usr/bin/env python
"""Driver for ncu captures: runs a few launches of ONE kernel flavour and exits.

    python benchmarks/prof.py <case> [N] [launches]
    cases: f32 (headline, TMA family) | f32_direct | u8 | u8_exact | u8_k10 | dense_sym | dense_chain | dense_mma |
           rgb | any_f32 | any_u8 | metrics_f32 | metrics_u8 | fwd | inv | zigzag_fwd | coded_bits

Used as
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/<name> python benchmarks/prof.py <case>
(see profiles/README in DESIGN.md section 5); tools/ncu_summary.py turns the report into profiles/*.txt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cuda_dct_idct_b200 as m


def dct2():
    k, n = np.mgrid[0:8, 0:8]
    c = np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8))
    return (c * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)


def main():
    case = sys.argv[1]
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dev = "cuda"
    x32 = torch.randint(0, 256, (N, N), device=dev, dtype=torch.int32).float()
    f = {}
    if case in ("f32", "f32_direct", "dense_sym", "dense_chain", "dense_mma", "metrics_f32", "fwd", "inv", "zigzag_fwd"):
        a = [x32.clone() for _ in range(2)]
        b = [torch.empty_like(x32) for _ in range(2)]
    if case == "f32":
        plan = m.Plan()
        step = lambda i: m.roundtrip(a[i % 2], out=b[i % 2], plan=plan)
    elif case == "f32_direct":
        plan = m.Plan(path=m.api.PATH_DIRECT)
        step = lambda i: m.roundtrip(a[i % 2], out=b[i % 2], plan=plan)
    elif case in ("dense_sym", "dense_chain", "dense_mma"):
        plan = m.Plan(T=dct2(), dense={"dense_chain": m.api.DENSE_CHAIN, "dense_mma": m.api.DENSE_MMA}.get(case, m.api.DENSE_AUTO),
                      path=m.api.PATH_DIRECT)
        step = lambda i: m.roundtrip(a[i % 2], out=b[i % 2], plan=plan)
    elif case == "fwd":
        plan = m.Plan()
        step = lambda i: m.forward(a[i % 2], coef=b[i % 2], plan=plan)
    elif case == "inv":
        plan = m.Plan()
        m.forward(a[0], coef=a[1], plan=plan)
        step = lambda i: m.inverse(a[1], img=b[i % 2], plan=plan)
    elif case == "zigzag_fwd":
        plan = m.Plan()
        zz = m.api.empty_zigzag(N, N, dev)
        step = lambda i: m.forward(a[i % 2], coef=zz, plan=plan, zigzag=True)
    elif case == "metrics_f32":
        plan = m.Plan()
        step = lambda i: m.roundtrip_with_metrics(a[i % 2], out=b[i % 2], plan=plan)
    elif case in ("u8", "u8_exact", "u8_k10", "metrics_u8"):
        a = [x32.to(torch.uint8) for _ in range(2)]
        b = [torch.empty_like(a[0]) for _ in range(2)]
        plan = m.Plan(inverse=m.api.INVERSE_EXACT if case == "u8_exact" else m.api.INVERSE_AUTO,
                      keep=m.zigzag_mask(10) if case == "u8_k10" else m.ALL_COEFFS)
        if case == "metrics_u8":
            step = lambda i: m.roundtrip_with_metrics(a[i % 2], out=b[i % 2], plan=plan)
        else:
            step = lambda i: m.roundtrip(a[i % 2], out=b[i % 2], plan=plan)
    elif case == "rgb":
        a = [torch.randint(0, 256, (N, N, 3), device=dev, dtype=torch.uint8) for _ in range(2)]
        b = [torch.empty_like(a[0]) for _ in range(2)]
        plan = m.Plan()
        step = lambda i: m.roundtrip_rgb(a[i % 2], out=b[i % 2], plan=plan)
    elif case in ("any_f32", "any_u8"):
        M = N - 1
        x = x32[:M, :M].contiguous() if case == "any_f32" else x32[:M, :M].to(torch.uint8).contiguous()
        a = [x.clone() for _ in range(2)]
        b = [torch.empty_like(x) for _ in range(2)]
        plan = m.Plan()
        step = lambda i: m.roundtrip_any(a[i % 2], out=b[i % 2], plan=plan)
    elif case == "coded_bits":
        plan = m.Plan()
        zz = m.api.empty_zigzag(N, N, dev)
        m.forward(x32, coef=zz, plan=plan, zigzag=True)
        step = lambda i: m.coded_bits(zz)
    else:
        raise SystemExit(f"unknown case {case}")
    for i in range(L):
        step(i)
    torch.cuda.synchronize()
    print("ok", case, N, m.api.last_path())


if __name__ == "__main__":
    main()
