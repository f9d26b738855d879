#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call left in gpurun_out/ into the committed summaries
under profiles/ (run here, no GPU needed):

    python tools/ncu_summary.py r01 gpurun_out/prof_tma_f32_dyn8.ncu-rep gpurun_out/launches.csv [--no-traffic] [--tag 16384]

writes profiles/<round>_ncu_<kernel>.txt (key metrics of every captured launch),
profiles/<round>_launches.txt (per-kernel share of the launch list) and
profiles/<round>_traffic.json (dram bytes per launch, read by bench.py's roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
    "smsp__average_warp_latency_issue_stalled_sleeping.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = ""
    argv = list(sys.argv[1:])
    if "--tag" in argv:          # suffix of the summary's file name (same kernel at two sizes)
        i = argv.index("--tag")
        tag = "_" + argv[i + 1]
        del argv[i:i + 2]
    args = [a for a in argv if a != "--no-traffic"]
    write_traffic = "--no-traffic" not in sys.argv     # only the headline kernel feeds bench.py's roofline.traffic
    rnd, rep = args[0], args[1]
    launches = args[2] if len(args) > 2 else None
    os.makedirs("profiles", exist_ok=True)
    hdr, units, data = raw(rep)
    kname = data[0][hdr.index("Kernel Name")]
    short = kname.split("(")[0].replace("void ", "").replace("b200dct::", "").replace("<", "_").replace(">", "").replace(", ", "_").replace(" ", "")
    lines = [f"# ncu --set full --clock-control none --import-source on, {len(data)} launch(es) of {kname}",
             f"# source report: {os.path.basename(rep)} (gpurun_out/, not committed: 7 MB)", ""]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k:75s} {units[i]:14s} " + "  ".join(r[i] for r in data))
    short += tag
    with open(f"profiles/{rnd}_ncu_{short}.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    rd = [float(r[hdr.index("dram__bytes_read.sum")]) for r in data]
    wr = [float(r[hdr.index("dram__bytes_write.sum")]) for r in data]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index("dram__bytes_read.sum")]]
    traffic = {"kernel": kname, "dram_bytes_read_per_launch": sum(rd) / len(rd) * scale,
               "dram_bytes_write_per_launch": sum(wr) / len(wr) * scale,
               "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd) * scale,
               "note": "one ncu --set full capture, 8192x8192 f32; writes still dirty in the 126 MB L2 when the kernel "
                       "ends are not counted by dram__bytes_write, hence write < 268.4 MB",
               "report": os.path.basename(rep)}
    if write_traffic:
        with open(f"profiles/{rnd}_traffic.json", "w") as f:
            json.dump(traffic, f, indent=1)
    print("wrote", f"profiles/{rnd}_ncu_{short}.txt", f"profiles/{rnd}_traffic.json" if write_traffic else "")
    if launches:
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
        h = rows[0]
        ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
            a = agg.setdefault(r[ki], [0, 0.0, 1e30, 0.0])
            a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
        tot = sum(a[1] for a in agg.values())
        out = ["# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 20 --warmup 3 --no-baselines`",
               "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
               f"{'launches':>8s} {'total us':>10s} {'share':>7s} {'min us':>8s} {'max us':>8s}  kernel"]
        for n, (c, t, lo, hi) in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.append(f"{c:8d} {t:10.1f} {100 * t / tot:6.1f}% {lo:8.1f} {hi:8.1f}  {n[:110]}")
        with open(f"profiles/{rnd}_launches.txt", "w") as f:
            f.write("\n".join(out) + "\n")
        print("wrote", f"profiles/{rnd}_launches.txt")


if __name__ == "__main__":
    main()
