#!/usr/bin/env python
"""Writes tests/golden/libjpeg_color.npz: what the REAL libjpeg (libjpeg-turbo inside Pillow, the
reference's image dependency -- jpeglib.h, utils.cu:6) computes for the two colour conversions and
which Huffman tables it writes, so the CPU suite can pin oracle/dct_oracle.c's restatements.

  * ycc_in / rgb_out : one JPEG decoded twice by libjpeg, once with out_color_space = JCS_YCbCr
    (Pillow's draft("YCbCr")) and once as RGB: pairs (Y,Cb,Cr) -> (R,G,B) of jdcolor.c.
  * rgb_in / ycc_out : an image of constant-colour 8x8 blocks encoded at quality 100 / 4:4:4 and
    decoded as YCbCr.  A constant block has only a DC term, which survives Q = 1 and the inverse
    DCT exactly, so the decoded sample IS jccolor.c's RGB -> YCbCr result for that colour.
  * dht_<class><id>_bits / _vals : the DHT segments of a baseline file written with optimize=False
    (jstdhuff.c = ITU-T T.81 Annex K.3).
Run in the build container (needs Pillow); the output is committed."""
import io
import os

import numpy as np
from PIL import Image, features

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(2026)
    # decoder side
    noise = rng.integers(0, 256, (160, 160, 3), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(noise).save(buf, "JPEG", quality=100, subsampling=0)
    a = Image.open(io.BytesIO(buf.getvalue()))
    a.draft("YCbCr", a.size)
    assert a.mode == "YCbCr"
    ycc = np.array(a).reshape(-1, 3)
    rgb = np.array(Image.open(io.BytesIO(buf.getvalue()))).reshape(-1, 3)
    ycc, idx = np.unique(ycc, axis=0, return_index=True)
    rgb = rgb[idx]
    # encoder side
    nb = 160
    tri = rng.integers(0, 256, (nb, nb, 3), dtype=np.uint8)
    tri[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255]]
    big = np.repeat(np.repeat(tri, 8, 0), 8, 1)
    buf2 = io.BytesIO()
    Image.fromarray(big).save(buf2, "JPEG", quality=100, subsampling=0)
    b = Image.open(io.BytesIO(buf2.getvalue()))
    b.draft("YCbCr", b.size)
    got = np.array(b).reshape(nb, 8, nb, 8, 3)
    assert (got == got[:, :1, :, :1]).all(), "constant blocks must decode to constant blocks"
    out = {"ycc_in": ycc, "rgb_out": rgb, "rgb_in": tri.reshape(-1, 3), "ycc_out": got[:, 0, :, 0].reshape(-1, 3),
           "libjpeg": np.array(f"{features.version('jpg')} turbo={features.check_feature('libjpeg_turbo')}")}
    # Huffman tables
    buf3 = io.BytesIO()
    Image.fromarray(noise).save(buf3, "JPEG", quality=75, optimize=False)
    d = buf3.getvalue()
    i = 2
    while i < len(d):
        m, L = d[i + 1], (d[i + 2] << 8) | d[i + 3]
        if m == 0xC4:
            seg, j = d[i + 4:i + 2 + L], 0
            while j < len(seg):
                bits = np.frombuffer(seg[j + 1:j + 17], np.uint8)
                n = int(bits.sum())
                out[f"dht_{seg[j] >> 4}{seg[j] & 15}_bits"] = bits.copy()
                out[f"dht_{seg[j] >> 4}{seg[j] & 15}_vals"] = np.frombuffer(seg[j + 17:j + 17 + n], np.uint8).copy()
                j += 17 + n
        if m == 0xDA:
            break
        i += 2 + L
    np.savez_compressed(os.path.join(HERE, "libjpeg_color.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
