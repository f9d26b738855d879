"""Regenerates tests/golden/oracle_*.npz from the CPU oracle (oracle/dct_oracle.c).

The reference has no tests or golden vectors of its own, and it is a
CUDA program, so it cannot be executed in the CPU-only build container.  Two kinds of
fixtures therefore exist:
  * oracle_*.npz   -- written HERE by this script from the CPU restatement;
  * refgpu_*.npz   -- written ON THE B200 BOX by make_ref_golden.py from the unmodified
                      reference kernels (oracle/_ref); they pin the oracle to the
                      reference's real outputs and are checked by the CPU suite too.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as o  # noqa: E402
import inputs  # noqa: E402


def main():
    o.build()
    # 1) the reference benchmark's generator, 256x256 (config 1)
    img = o.rand_image(256, 256, 42)
    out, coef = o.roundtrip(img, want_coef=True)
    np.savez_compressed(os.path.join(HERE, "oracle_rand256.npz"),
                        img=img.astype(np.uint8), coef=coef.astype(np.int16), out=out,
                        out_u8=o.to_u8(out))
    # 2) adversarial strip
    adv = inputs.adversarial(16)
    out, coef = o.roundtrip(adv, want_coef=True)
    np.savez_compressed(os.path.join(HERE, "oracle_adversarial.npz"),
                        img=adv.astype(np.uint8), coef=coef.astype(np.int16), out=out)
    # 3) retained-coefficient masks k = 6..10 on a 64x64 crop
    crop = img[:64, :64].copy()
    d = {"img": crop.astype(np.uint8)}
    for k in (6, 7, 8, 9, 10):
        out, coef = o.roundtrip(crop, keep=o.zigzag_mask(k), want_coef=True)
        d[f"coef_k{k}"] = coef.astype(np.int16)
        d[f"out_k{k}"] = out
    np.savez_compressed(os.path.join(HERE, "oracle_masks.npz"), **d)
    # 4) exact DCT-II matrix (dense variants), 64x64 crop
    T = o.dct2_T()
    out, coef = o.roundtrip(crop, T=T, want_coef=True)
    np.savez_compressed(os.path.join(HERE, "oracle_dense_dct2.npz"), img=crop.astype(np.uint8), T=T,
                        coef=coef.astype(np.int16), out=out)
    # 5) colour path and coded size (round 2): 48x64 RGB
    rng = np.random.default_rng(2026)
    yy, xx = np.mgrid[0:48, 0:64]
    rgb = np.stack([(128 + 100 * np.sin(xx / 7.0 + c) * np.cos(yy / 9.0 - c) + rng.normal(0, 6, (48, 64))) for c in range(3)], -1)
    rgb = rgb.clip(0, 255).astype(np.uint8)
    out, planes, coef3 = o.roundtrip_rgb(rgb, want_planes=True, want_coef=True)
    bits = np.array([o.coded_bits(o.zigzag_i16(coef3[c]), 0 if c == 0 else 1) for c in range(3)], np.int64)
    np.savez_compressed(os.path.join(HERE, "oracle_rgb.npz"), rgb=rgb, out=out, planes=planes,
                        coef=coef3.astype(np.int16), coded_bits=bits)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
