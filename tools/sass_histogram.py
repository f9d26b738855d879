#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass), committed as
profiles/<round>_sass_histogram.txt: the evidence that the hot kernels are TMA / packed-FP32
Blackwell code (UTMALDG, UTMASTG, FFMA2, ...) and use no tensor-core or library instructions.

    python tools/sass_histogram.py r02 [path/to/libb200dct.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["total", "FFMA2", "FFMA", "FADD2", "FADD", "FMUL2", "FMUL", "FRND", "F2I", "F2IP", "I2F", "PRMT", "LOP3", "IMAD", "IDP", "MOV",
        "LDG", "STG", "LDS", "STS", "UTMALDG", "UTMASTG", "SYNCS", "LDC", "LDCU", "ATOMG", "ACQBULK", "HMMA", "UTCMMA", "BAR"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [o.replace("b200dct::", "").replace("void ", "") for o in out[: len(names)]]


def main():
    rnd = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "cuda-dct-idct_b200", "libb200dct.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            cur[op] += 1
    names = demangle(list(kernels))
    lines = [f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instructions per kernel (static counts; one thread = one 8x8 block per pass,",
             "# or per tile iteration in the persistent k_tma kernels).  Template arguments: k_direct<MODE 0 fwd/1 inv/2 round trip, TK 0 dense chain/1 Haweel/2 dense symmetric,",
             "# QMODE 0 immediates/1 param tables/2 param tables + IEEE division/3..7 compile-time masks k=6..10, PIX 0 f32/1 u8, METRICS, FINV factored inverse>", ""]
    lines.append(f"{'kernel':58s} " + " ".join(f"{c:>7s}" for c in COLS))
    tot = collections.Counter()
    for (mangled, cnt), name in sorted(zip(kernels.items(), names), key=lambda x: x[1]):
        short = re.sub(r"\((?:int|bool|unsigned long)\)", "", name)
        short = re.sub(r"\((?:b200dct::)?\w+Params\)$|\([^()]*\)$", "", short)
        lines.append(f"{short[:58]:58s} " + " ".join(f"{cnt.get(c, 0):7d}" for c in COLS))
        tot.update(cnt)
    lines.append("")
    lines.append(f"{'ALL KERNELS':58s} " + " ".join(f"{tot.get(c, 0):7d}" for c in COLS))
    other = sorted(((k, v) for k, v in tot.items() if k not in COLS), key=lambda x: -x[1])
    lines.append("")
    lines.append("other mnemonics: " + ", ".join(f"{k} {v}" for k, v in other))
    path = os.path.join(ROOT, "profiles", f"{rnd}_sass_histogram.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
