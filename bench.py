#!/usr/bin/env python
"""bench.py -- Gpixel/s of the fused DCT -> quantise -> IDCT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]

One "step" = one fused pass (b200dct_roundtrip, HpApprDCT: Haweel T, JPEG Q, all
coefficients kept) over one 8192 x 8192 fp32 image per GPU -- BASELINE.json configs[1].
For N > 1 (launched with torchrun, one rank per GPU) every rank owns one 8192-row stripe
of an (N*8192) x 8192 image: block-row striping, no halo, no data-path collective, weak
scaling.  Rank 0 prints ONE JSON line.

Legs of the default arm, all on the same workload:
  value        device-resident: inputs already in HBM, K launches timed with CUDA events on
               the launching stream, barrier + synchronize on both sides, max over ranks;
               the step rotates over 4 input/output buffer pairs (2 GiB) so nothing is L2
               resident (one image pair alone is 512 MiB against a 126 MB L2).
  roofline     algorithmic bytes per launch (8 B/px: 4 read + 4 written) / average launch
               time from the same events, against the measured HBM copy peak
               (MEASURED_PEAKS.json, else the profiling guide's fallback).
  e2e          the same metric through the host-buffer entry point
               (b200dct_roundtrip_host): pinned host image in, pinned host image out, H2D
               and D2H inside the timed region, every step.
  cpu_baseline the sequential CPU restatement of the reference (oracle/, 1 thread) on the
               whole 8192^2 image, rank 0 / N=1 only.  A reported baseline, not the target.
  reference_gpu the UNMODIFIED reference kernels (oracle/_ref, HpApprDCT recompiled for
               sm_100a) on the same device buffers: its own printed event times.

--impl reference: the reference arm of the contract.  The reference ships no CPU code
(SURVEY.md S1), so this is the oracle port on all host threads, rank 0 only.
--impl reference-gpu: the unmodified reference GPU program flow (host image -> H2D ->
dct_all_blocks_cuda -> idct_all_blocks_cuda -> D2H) from oracle/_ref, for context.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SIDE = 8192
METRIC = "Gpixel/s DCT+quant+IDCT at 8192^2 fp32 (HpApprDCT fused round trip)"
UNIT = "Gpixel/s"
BYTES_PER_PX = 8.0  # algorithmic: 4 B read + 4 B written per pixel


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs
    (NVML, the same counters `nvidia-smi --query-gpu=clocks.sm,...` prints)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None
            return self
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def _run(self):
        nv = self._nvml
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        d = {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
             "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        return d


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel from the
    committed ncu capture (profiles/*traffic.json), or None."""
    try:
        best = None
        pd = os.path.join(ROOT, "profiles")
        for f in sorted(os.listdir(pd)):
            if f.endswith("traffic.json"):
                with open(os.path.join(pd, f)) as fh:
                    best = json.load(fh)
        return None if best is None else float(best["dram_bytes_per_launch"])
    except Exception:
        return None


# ------------------------------------------------------------------ reference arms
def run_reference_cpu(args):
    """Contract's reference arm: CPU, host cores only, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import oracle as o

    o.build()
    rows = 2048                                     # bounded sample: a 2048-row stripe of the 8192^2 image
    img = o.rand_image(rows, N_SIDE, 42)
    threads_all = max(1, len(os.sched_getaffinity(0)))
    # pick the thread count that is actually faster on this box (containers often cap CPU time)
    t1 = o.time_roundtrip(img, reps=1, threads=1)
    tn = o.time_roundtrip(img, reps=1, threads=threads_all) if threads_all > 1 else t1
    threads = threads_all if tn < t1 else 1
    # keep the whole run bounded (~90 s) whatever K the driver passes: shrink the stripe
    per_step = min(t1, tn)
    budget = 90.0 / max(1, args.steps + args.warmup)
    if per_step > budget:
        rows = max(64, int(rows * budget / per_step) // 8 * 8)
        img = np.ascontiguousarray(img[:rows])
    out = np.empty_like(img)
    for _ in range(args.warmup):
        o.time_roundtrip(img, reps=1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.lib().oracle_roundtrip(img, rows, N_SIDE, o.haweel_T(), o.jpeg_Q(), o.ALL_COEFFS, None, out, threads)
    dt = time.perf_counter() - t0
    gpx = rows * N_SIDE * args.steps / dt / 1e9
    sample = f"{rows}x{N_SIDE} f32 stripe of the 8192^2 image per step, srand(42) rand()%256"
    line = {
        "impl": "reference", "metric": METRIC, "value": gpx, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "8192x8192 fp32 HpApprDCT DCT+quant+IDCT (BASELINE configs[1]); CPU sample: " + sample,
                   "note": "the reference ships no CPU implementation (SURVEY.md S1); this is the oracle port of its "
                           "arithmetic (oracle/dct_oracle.c, gcc -O2, OpenMP over block-rows)"},
        "cpu_baseline": {"value": gpx, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "t_1thread_s": t1, "t_allthreads_s": tn, "host_threads_available": threads_all},
        "e2e": {"value": gpx, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_reference_gpu(args):
    """The unmodified reference GPU flow, for context (not the contract's reference arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refgpu
    from oracle import oracle as o

    if not refgpu.available("newappr") or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libref_newappr.so not built or no GPU"}))
        return 0
    torch.cuda.set_device(0)
    refgpu.set_quant("newappr", o.jpeg_Q())
    T = torch.from_numpy(o.haweel_T()).cuda()
    h_img = torch.randint(0, 256, (N_SIDE, N_SIDE), dtype=torch.int32).float().pin_memory()
    h_out = torch.empty_like(h_img).pin_memory()
    d_img, d_coef, d_rec = (torch.empty(N_SIDE, N_SIDE, device="cuda") for _ in range(3))
    kms = []

    def step():
        d_img.copy_(h_img, non_blocking=True)
        _, t1 = refgpu.dct("newappr", d_img, T, d_coef)
        _, t2 = refgpu.idct("newappr", d_coef, T, d_rec)
        h_out.copy_(d_rec, non_blocking=True)
        torch.cuda.synchronize()
        kms.append(t1 + t2)

    for _ in range(args.warmup):
        step()
    kms.clear()
    steps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    px = N_SIDE * N_SIDE
    line = {
        "impl": "reference-gpu", "metric": METRIC, "value": px * steps / dt / 1e9, "unit": UNIT, "n_gpus": 1,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "8192x8192 fp32, reference main() flow: pinned host -> H2D -> dct_all_blocks_cuda -> "
                               "idct_all_blocks_cuda -> D2H (main_newAppr.cu:88-124), HpApprDCT recompiled for sm_100a"},
        "kernel_only": {"ms": statistics.mean(kms), "gpixel_s": px / (statistics.mean(kms) * 1e-3) / 1e9,
                        "what": "sum of the reference's own printed DCT and IDCT event times (6 launches, 48 B/px)"},
        "e2e": {"value": px * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": px * 4, "d2h_bytes_per_step": px * 4},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch

    import cuda_dct_idct_b200 as m

    rank, local_rank, world = m.dist.init()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    m.lib()
    plan = m.Plan()
    px = N_SIDE * N_SIDE                       # pixels per rank per step
    H0, H1 = m.stripe_rows(N_SIDE * world, world, rank)   # this rank's stripe of the (N*8192) x 8192 image
    assert H1 - H0 == N_SIDE

    # ---- device-resident leg: 4 rotating buffer pairs, synthetic integers 0..255
    NBUF = 4
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    ins = [torch.randint(0, 256, (N_SIDE, N_SIDE), device=dev, generator=g, dtype=torch.int32).float() for _ in range(NBUF)]
    outs = [torch.empty(N_SIDE, N_SIDE, device=dev) for _ in range(NBUF)]
    stream = torch.cuda.current_stream()
    launches = 0

    def step(i):
        nonlocal launches
        m.roundtrip(ins[i % NBUF], out=outs[i % NBUF], plan=plan, stream=stream)
        launches += m.api.last_launch_count()

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    kernel_path = m.api.last_path()

    sampler = ClockSampler(local_rank).start()
    m.dist.barrier()
    torch.cuda.synchronize()
    launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    torch.cuda.synchronize()
    m.dist.barrier()
    ms_local = ev0.elapsed_time(ev1)
    timed_launches = launches
    # keep the identical load running briefly if the timed region was too short to sample clocks
    extra = 0
    t_end = time.perf_counter() + (0.0 if ms_local > 400 else 0.6)
    while time.perf_counter() < t_end:
        for i in range(64):
            m.roundtrip(ins[i % NBUF], out=outs[i % NBUF], plan=plan, stream=stream)
        torch.cuda.synchronize()
        extra += 64
    clocks = sampler.stop()
    clocks["sampled_over"] = "timed region" if extra == 0 else f"timed region + {extra} identical launches (region < 0.4 s)"
    ms_total = m.dist.max_over_ranks(ms_local, dev)
    value = px * world * args.steps / (ms_total * 1e-3) / 1e9
    ms_per_step = ms_total / args.steps

    # per-launch duration of the dominant kernel: the timed region is nothing but K
    # back-to-back launches of it on this stream, so event time / K is its average duration
    peak, peak_src = measured_peak()
    k_ms = ms_local / args.steps
    achieved = BYTES_PER_PX * px / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": f"k_{kernel_path}<RT,sparse,Q_IMM,f32>",
                "algorithmic_bytes_per_launch": BYTES_PER_PX * px, "avg_launch_ms": k_ms,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- e2e leg: pinned host buffers through the host-buffer entry point
    h_in = torch.randint(0, 256, (N_SIDE, N_SIDE), dtype=torch.int32).float().pin_memory()
    h_out = torch.empty(N_SIDE, N_SIDE).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        m.roundtrip_host(h_in, h_out, plan=plan)
    m.dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.roundtrip_host(h_in, h_out, plan=plan)     # synchronous: returns with h_out complete
    torch.cuda.synchronize()
    e2e_local = time.perf_counter() - t0
    m.dist.barrier()
    e2e_s = m.dist.max_over_ranks(e2e_local, dev)
    e2e = {"value": px * world * e2e_steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": px * 4 * world,
           "d2h_bytes_per_step": px * 4 * world, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "api": "b200dct_roundtrip_host (pinned host in/out, chunked H2D/kernel/D2H pipeline)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{N_SIDE}x{N_SIDE} fp32 image per GPU, HpApprDCT (Haweel T, JPEG luminance Q, 64 coefficients kept), "
                               "fused DCT+quant+dequant+IDCT (BASELINE configs[1])",
                   "striping": f"{world} block-row stripe(s) of a {N_SIDE * world}x{N_SIDE} image, no halo, no collective",
                   "l2": "inputs larger than L2: 4 rotating in/out pairs, 2 GiB working set vs 126 MB L2",
                   "launch": "K back-to-back launches on one stream; consecutive launches overlap tail and set-up through "
                             "programmatic dependent launch (each kernel waits for its predecessor before touching memory)",
                   "kernel_path": kernel_path},
        "roofline": roofline, "e2e": e2e, "gpu_launches": timed_launches, "clocks": clocks,
    }

    # ---- baselines, rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_baselines:
        line["cpu_baseline"] = cpu_baseline()
        ref = reference_gpu_kernels(ins[0], outs[0])
        if ref:
            line["reference_gpu"] = ref
        # parity spot check of what was just timed (oracle as the checker only)
        try:
            from oracle import oracle as o

            band = ins[0][:16].cpu().numpy()
            line["parity_spot_check"] = bool(np.array_equal(outs[0][:16].cpu().numpy().view(np.uint32),
                                                            o.roundtrip(band).view(np.uint32)))
        except Exception as e:  # pragma: no cover
            line["parity_spot_check"] = f"skipped: {e}"
    if rank == 0:
        print(json.dumps(line))
    m.dist.barrier()
    m.dist.shutdown()
    return 0


def cpu_baseline():
    from oracle import oracle as o

    o.build()
    img = o.rand_image(N_SIDE, N_SIDE, 42)
    t = o.time_roundtrip(img, reps=3, threads=1)     # ~2 s per pass: ~10 s of CPU work in total
    return {"value": N_SIDE * N_SIDE / t / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "the whole 8192x8192 f32 image (srand(42) rand()%256), best of 3 passes, 1 thread",
            "seconds_per_pass": t, "host_threads_available": len(os.sched_getaffinity(0)),
            "what": "sequential C restatement of the reference's arithmetic (oracle/dct_oracle.c); the reference "
                    "itself has no CPU implementation (SURVEY.md S1)"}


def reference_gpu_kernels(d_img, d_out):
    """The unmodified reference kernels on the same B200: its own printed event times."""
    try:
        import torch

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refgpu
        from oracle import oracle as o

        if not refgpu.available("newappr"):
            return None
        refgpu.set_quant("newappr", o.jpeg_Q())
        T = torch.from_numpy(o.haweel_T()).cuda()
        img = d_img.clone()
        coef = torch.empty_like(img)
        res = {}
        for variant in ("newappr", "fastappr"):
            if not refgpu.available(variant):
                continue
            ts = []
            for _ in range(4):
                img.copy_(d_img)
                _, t1 = refgpu.dct(variant, img, T, coef)
                _, t2 = refgpu.idct(variant, coef, T, d_out)
                ts.append(t1 + t2)
            ms = min(ts[1:])
            res[variant] = {"dct_plus_idct_ms": ms, "gpixel_s": N_SIDE * N_SIDE / (ms * 1e-3) / 1e9}
        res["what"] = ("reference kernels compiled unmodified for sm_100a (oracle/_ref), 8192^2 f32 resident in HBM, "
                       "its own cudaEvent times for dct_all_blocks_cuda + idct_all_blocks_cuda (6 launches, 48 B/px)")
        return res
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the CPU / reference-GPU baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 2000:
            args.steps, args.warmup = 10, 3
        return run_reference_cpu(args)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
