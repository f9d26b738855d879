"""CPU: the colour path and the compression-factor definition of the oracle, pinned against the
real libjpeg (tests/golden/libjpeg_color.npz, written by make_libjpeg_color_golden.py)."""
import io
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libjpeg_color.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_colour_conversions_match_libjpeg(oracle, gold):
    rgb_in = gold["rgb_in"].reshape(-1, 1, 3)
    y, cb, cr = oracle.rgb_to_ycc(rgb_in)
    got = np.stack([y, cb, cr], -1).reshape(-1, 3)
    assert np.array_equal(got, gold["ycc_out"]), "RGB -> YCbCr differs from libjpeg's jccolor.c"
    ycc = gold["ycc_in"]
    rgb = oracle.ycc_to_rgb(ycc[:, 0].reshape(-1, 1), ycc[:, 1].reshape(-1, 1), ycc[:, 2].reshape(-1, 1))
    assert np.array_equal(rgb.reshape(-1, 3), gold["rgb_out"]), "YCbCr -> RGB differs from libjpeg's jdcolor.c"


def test_huffman_tables_are_libjpegs(oracle, gold):
    for which in (0, 1):
        for table in (0, 1):
            bits, vals = oracle.huffman_spec(which, table)
            assert np.array_equal(bits, gold[f"dht_{which}{table}_bits"])
            assert np.array_equal(vals, gold[f"dht_{which}{table}_vals"])
    # Annex K.3: DC luminance category 0 has a 2-bit code, EOB 4 bits (luma) / 2 bits (chroma), ZRL 11 / 10
    assert oracle.huffman_lengths(0, 0)[0] == 2 and oracle.huffman_lengths(1, 0)[0x00] == 4
    assert oracle.huffman_lengths(1, 1)[0x00] == 2 and oracle.huffman_lengths(1, 0)[0xF0] == 11 and oracle.huffman_lengths(1, 1)[0xF0] == 10


class BitWriter:
    def __init__(self):
        self.out, self.acc, self.n, self.bits = bytearray(), 0, 0, 0

    def put(self, code, length):
        self.bits += length
        self.acc = (self.acc << length) | code
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)      # byte stuffing
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)
            self.bits -= 0  # padding counted below by the caller


def huff_codes(bits, vals):
    codes, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(int(bits[length - 1])):
            codes[int(vals[k])] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return codes


def write_baseline_jpeg(stream, H, W, Q, oracle):
    """A complete baseline JPEG file (grayscale, Annex K tables) around a zig-zag coefficient stream."""
    zz = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    dcb, dcv = oracle.huffman_spec(0, 0)
    acb, acv = oracle.huffman_spec(1, 0)
    dc, ac = huff_codes(dcb, dcv), huff_codes(acb, acv)

    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + bytes(payload)

    f = bytearray(b"\xff\xd8")
    f += seg(0xDB, bytes([0]) + bytes(int(Q[zz[k]]) for k in range(64)))
    f += seg(0xC0, bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([1, 1, 0x11, 0]))
    f += seg(0xC4, bytes([0x00]) + bytes(dcb.tolist()) + bytes(dcv.tolist()))
    f += seg(0xC4, bytes([0x10]) + bytes(acb.tolist()) + bytes(acv.tolist()))
    f += seg(0xDA, bytes([1, 1, 0x00, 0, 63, 0]))
    bw, prev = BitWriter(), 0

    def put_value(v, size):
        if size:
            bw.put(v if v >= 0 else v + (1 << size) - 1, size)

    for blk in stream.reshape(-1, 64).astype(int):
        d = int(blk[0]) - prev
        prev = int(blk[0])
        s = abs(d).bit_length()
        bw.put(*dc[s])
        put_value(d, s)
        run = 0
        for k in range(1, 64):
            v = int(blk[k])
            if v == 0:
                run += 1
                continue
            while run > 15:
                bw.put(*ac[0xF0])
                run -= 16
            s = abs(v).bit_length()
            bw.put(*ac[(run << 4) | s])
            put_value(v, s)
            run = 0
        if run:
            bw.put(*ac[0x00])
    coded_bits = bw.bits
    bw.flush()
    f += bw.out + b"\xff\xd9"
    return bytes(f), coded_bits


def test_coded_bits_is_the_size_of_a_real_jpeg_scan(oracle):
    """The zig-zag stream + Annex K tables form a valid baseline JPEG: the real libjpeg decodes the
    file written around OUR coefficients into (nearly) our reconstruction, and the number of
    entropy-coded bits in that file is exactly oracle_coded_bits()."""
    from PIL import Image

    H = W = 64
    yy, xx = np.mgrid[0:H, 0:W]
    img = (128 + 60 * np.sin(xx / 9.0) + 50 * np.cos(yy / 7.0) + 20 * np.sin((xx + yy) / 3.0)).clip(0, 255).astype(np.uint8)
    img[8:24, 8:24] = np.random.default_rng(0).integers(0, 256, (16, 16))          # some busy blocks
    rec, coef = oracle.roundtrip(img, want_coef=True)
    stream = oracle.zigzag_i16(coef)
    data, bits = write_baseline_jpeg(stream, H, W, oracle.jpeg_Q(), oracle)
    assert bits == oracle.coded_bits(stream, 0)
    dec = np.array(Image.open(io.BytesIO(data)))
    assert dec.shape == (H, W)
    # Haweel's T approximates the DCT, so libjpeg's exact inverse DCT lands close to our own reconstruction
    smooth = np.ones((H, W), bool)
    smooth[8:24, 8:24] = False
    err_dec = np.mean((dec.astype(float) - img)[smooth] ** 2)
    err_own = np.mean((oracle.to_u8(rec).astype(float) - img)[smooth] ** 2)
    print(f"libjpeg decode of our stream: MSE {err_dec:.1f} vs own inverse {err_own:.1f} (smooth part)")
    assert err_dec < 150.0 and err_dec < 0.02 * np.mean((img.astype(float) - 128)[smooth] ** 2) + 100
    cf = oracle.compression_factor(coef, 0)
    assert abs(cf - 8.0 * H * W / bits) < 1e-12 and cf > 1.0
    # fewer retained coefficients -> fewer bits
    b10 = oracle.coded_bits(oracle.zigzag_i16(oracle.roundtrip(img, keep=oracle.zigzag_mask(10), want_coef=True)[1]))
    b6 = oracle.coded_bits(oracle.zigzag_i16(oracle.roundtrip(img, keep=oracle.zigzag_mask(6), want_coef=True)[1]))
    assert b6 <= b10 <= bits
    # edge cases: an all-zero plane is DC category 0 + EOB per block; a run of 16+ zeros needs ZRL
    z = np.zeros((3, 64), np.int16)
    assert oracle.coded_bits(z, 0) == 3 * (2 + 4) and oracle.coded_bits(z, 1) == 3 * (2 + 2)
    z[0, 40] = 5          # 39 zeros in front: two ZRL + (7,3) code + 3 bits, then EOB
    ac = oracle.huffman_lengths(1, 0)
    assert oracle.coded_bits(z[:1], 0) == 2 + 2 * 11 + int(ac[0x73]) + 3 + 4
    z[0, 63] = -1         # last coefficient non-zero: no EOB
    assert oracle.coded_bits(z[:1], 0) == 2 + 2 * 11 + int(ac[0x73]) + 3 + (11 + int(ac[0x61]) + 1)


def test_rgb_round_trip_structure(oracle):
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (32, 48, 3), dtype=np.uint8)
    out, planes, coef = oracle.roundtrip_rgb(rgb, want_planes=True, want_coef=True)
    y, cb, cr = oracle.rgb_to_ycc(rgb)
    for c, (p, q) in enumerate(((y, oracle.jpeg_Q()), (cb, oracle.jpeg_Q_chroma()), (cr, oracle.jpeg_Q_chroma()))):
        rec, cf = oracle.roundtrip(p, Q=q, want_coef=True)
        assert np.array_equal(cf.view(np.uint32), coef[c].view(np.uint32))
        assert np.array_equal(oracle.to_u8(rec), planes[c])
    assert np.array_equal(out, oracle.ycc_to_rgb(planes[0], planes[1], planes[2]))
    # a grey image stays grey-ish and its chroma planes carry (almost) nothing
    grey = np.repeat(rng.integers(0, 256, (16, 16, 1), dtype=np.uint8), 3, 2)
    _, _, cg = oracle.roundtrip_rgb(grey, want_planes=True, want_coef=True)
    assert np.count_nonzero(cg[1]) == 0 and np.count_nonzero(cg[2]) == 0


def test_rgb_golden_fixture(oracle):
    """tests/golden/oracle_rgb.npz (make_golden.py): the colour path and the coded sizes stay what they were."""
    g = np.load(os.path.join(os.path.dirname(GOLD), "oracle_rgb.npz"))
    out, planes, coef = oracle.roundtrip_rgb(g["rgb"], want_planes=True, want_coef=True)
    assert np.array_equal(out, g["out"]) and np.array_equal(planes, g["planes"])
    assert np.array_equal(coef.astype(np.int16), g["coef"])
    bits = [oracle.coded_bits(oracle.zigzag_i16(coef[c]), 0 if c == 0 else 1) for c in range(3)]
    assert bits == g["coded_bits"].tolist()
