"""Randomised stress of the early paths: a random programme of library round trips (both families, in place
and out of place, single stream and two event-chained streams) interleaved with foreign kernels (torch copies,
fills, arithmetic) that write the buffers the next calls read.  The programme runs twice: with a device
synchronisation after every operation (nothing can overlap: the reference result) and asynchronously
(dependent launches, early loads); every buffer must end up bit-identical."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cuda_dct_idct_b200 as m


def programme(seed, n_ops, kinds):
    rnd = random.Random(seed)
    ops = []
    for _ in range(n_ops):
        kind = rnd.choice(kinds)
        nb = NBUF[kind]
        r = rnd.random()
        a, b = rnd.randrange(nb), rnd.randrange(nb)
        if r < 0.55:
            ops.append(("rt", kind, a, b))            # may be in place (a == b)
        elif r < 0.70:
            ops.append(("copy", kind, a, b))
        elif r < 0.80:
            ops.append(("fill", kind, a, rnd.randrange(256)))
        elif r < 0.86:
            ops.append(("add", kind, a, b))
        elif r < 0.93 and kind in COEF:
            c = rnd.randrange(NCOEF)
            ops.append(("fwd", kind, a, c) if rnd.random() < 0.5 else ("inv", kind, c, b))   # the split API through coefficient planes
        else:
            ops.append(("batch", kind, a, b))
    return ops


NBUF = {"u8": 5, "f32d": 4, "f32t": 4, "rgb": 4, "anyf": 4}
SHAPE = {"u8": (8192, 8192, torch.uint8), "f32d": (4096, 4096, torch.float32), "f32t": (6144, 6144, torch.float32),
         "rgb": (4096, 4096, 3, torch.uint8), "anyf": (6001, 6007, torch.float32)}


COEF = {"u8": torch.int16, "f32d": torch.float32, "f32t": torch.float32}   # coefficient plane dtype per kind
NCOEF = 2


def run(ops, sync, plans, init):
    bufs = {k: [t.clone() for t in init[k]] for k in init}
    coefs = {k: [torch.zeros(SHAPE[k][:2], device="cuda", dtype=COEF[k]) for _ in range(NCOEF)] for k in COEF if k in init}
    s = torch.cuda.current_stream()
    for op in ops:
        kind = op[1]
        B = bufs[kind]
        if op[0] == "rt":
            if kind == "rgb":
                m.roundtrip_rgb(B[op[2]], out=B[op[3]], plan=plans[kind])
            elif kind == "anyf":
                m.roundtrip_any(B[op[2]], out=B[op[3]], plan=plans[kind])
            else:
                m.roundtrip(B[op[2]], out=B[op[3]], plan=plans[kind])
        elif op[0] == "copy":
            if op[2] != op[3]:
                B[op[3]].copy_(B[op[2]])
        elif op[0] == "fill":
            B[op[2]].fill_(op[3])
        elif op[0] == "add":
            if B[op[2]].dtype == torch.uint8:
                B[op[2]].bitwise_xor_(B[op[3]])
            else:
                B[op[2]].add_(B[op[3]]).clamp_(0, 255).floor_()
        elif op[0] == "fwd":
            m.forward(B[op[2]], coef=coefs[kind][op[3]], plan=plans[kind])
        elif op[0] == "inv":
            m.inverse(coefs[kind][op[2]], img=B[op[3]], plan=plans[kind])
        elif op[0] == "batch":
            i, j = op[2], op[3]
            if i != j and kind not in ("rgb", "anyf"):
                m.roundtrip_batch([B[i], B[j]], outs=[B[i], B[j]], plan=plans[kind])
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    for k in coefs:
        bufs[k] = bufs[k] + coefs[k]   # coefficient planes are compared as well
    return bufs


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    plans = {"u8": m.Plan(), "f32d": m.Plan(), "f32t": m.Plan(), "rgb": m.Plan(), "anyf": m.Plan()}
    g = torch.Generator(device="cuda").manual_seed(3)
    init = {k: [torch.randint(0, 256, SHAPE[k][:-1], device="cuda", generator=g, dtype=torch.int32).to(SHAPE[k][-1]) for _ in range(NBUF[k])]
            for k in NBUF}
    bad = 0
    for trial in range(trials):
        kinds = [["u8"], ["f32t"], ["f32d"], ["rgb"], ["anyf"], ["u8", "f32t", "f32d", "rgb", "anyf"]][trial % 6]
        ops = programme(100 + trial, n_ops, kinds)
        want = run(ops, True, plans, init)
        for rep in range(3):
            got = run(ops, False, plans, init)
            diff = sum(int(not torch.equal(x, y)) for k in want for x, y in zip(want[k], got[k]))
            bad += diff
            print(f"trial {trial} ({'+'.join(kinds)}, {n_ops} ops) async run {rep}: {diff} buffers differ", flush=True)
    print("STRESS", "FAILED" if bad else "ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
