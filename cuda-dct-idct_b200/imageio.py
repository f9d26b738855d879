"""Image files at the edges of the path: the host-side
replacement of the reference's libjpeg helpers `load_jpeg_as_matrix` / `save_grayscale_jpeg`
(utils.cu:38,98).  File decoding/encoding stays on the host (north_star), here through Pillow;
binary PGM (P5) is read and written without any dependency.  Pixels travel as uint8, so the
kernels' u8 path absorbs convertToFloat / convertToUnsignedChar (utils.cu:10-24).
"""
from __future__ import annotations

import numpy as np


def load_gray(path: str) -> np.ndarray:
    """H x W uint8, grayscale.  JPEG/PNG/... via Pillow, binary PGM natively."""
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"P5":
        return _read_pgm(path)
    from PIL import Image

    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("L"), dtype=np.uint8))


def save_gray(path: str, img: np.ndarray, quality: int = 100) -> None:
    """Writes an H x W uint8 image; JPEG quality 100 is the reference's setting
    (main_newAppr.cu:136)."""
    img = np.ascontiguousarray(img, np.uint8)
    if path.lower().endswith(".pgm"):
        with open(path, "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
            f.write(img.tobytes())
        return
    from PIL import Image

    Image.fromarray(img, mode="L").save(path, quality=quality)


def _read_pgm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        tokens.append(data[pos:end])
        pos = end
    if tokens[0] != b"P5" or int(tokens[3]) != 255:
        raise ValueError("only 8-bit binary PGM (P5, maxval 255) is supported")
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, np.uint8, count=w * h, offset=pos + 1).reshape(h, w).copy()


def crop_to_blocks(img: np.ndarray) -> np.ndarray:
    """The kernels need H and W multiples of 8 (the reference silently mis-computes otherwise,
    main_newAppr.cu:261-262): drop the ragged right/bottom remainder."""
    h, w = img.shape
    return np.ascontiguousarray(img[: h - h % 8, : w - w % 8])


def transform_file(src: str, dst: str, plan=None, quality: int = 100):
    """File -> DCT -> quantise -> IDCT -> file on the current GPU, the reference program's
    whole flow (main_newAppr.cu:26-165).  Returns (MSE, PEEN%) of the reconstruction."""
    from . import api

    img = crop_to_blocks(load_gray(src))
    out = api.roundtrip_host(img, plan=plan)
    save_gray(dst, out, quality)
    d = img.astype(np.float64) - out.astype(np.float64)
    sse, en = float((d * d).sum()), float((img.astype(np.float64) ** 2).sum())
    return sse / img.size, (100.0 * (sse / en) ** 0.5 if en > 0 else 0.0)
