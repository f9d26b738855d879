"""GPU: the colour entry point (b200dct_roundtrip_rgb) and the coded-size kernel
(b200dct_zigzag_coded_bits) against the oracle, which is pinned against the real libjpeg
(tests/test_color_cpu.py).  EXACT inverse: bit-exact planes, streams and RGB bytes; the library
default (factored inverse): planes within 1 LSB, hence RGB within 3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def images(shape, seed):
    H, W = shape
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = np.stack([128 + 100 * np.sin(xx / 11.0 + c) * np.cos(yy / 13.0 - c) + rng.normal(0, 4, (H, W)) for c in range(3)], -1)
    sat = rng.choice(np.array([0, 255], np.uint8), (H, W, 3))          # saturated primaries: range_limit on every pixel
    return {"noise": rng.integers(0, 256, (H, W, 3), dtype=np.uint8), "smooth": smooth.clip(0, 255).astype(np.uint8), "saturated": sat}


@pytest.mark.parametrize("shape", [(8, 8), (64, 96), (256, 256), (520, 1064)])
def test_rgb_round_trip_bit_exact(dct, oracle, shape):
    H, W = shape
    for name, rgb in images(shape, 3).items():
        want, planes, coef = oracle.roundtrip_rgb(rgb, want_planes=True, want_coef=True)
        d = torch.from_numpy(rgb).cuda()
        zz = torch.empty(3, H // 8, W // 8, 64, dtype=torch.int16, device="cuda")
        out = dct.roundtrip_rgb(d, streams=zz)
        assert dct.api.last_path() == "rgb"
        torch.cuda.synchronize()
        for c in range(3):
            assert np.array_equal(zz[c].cpu().numpy(), oracle.zigzag_i16(coef[c])), f"{name}: coefficient stream of plane {c}"
        assert np.array_equal(out.cpu().numpy(), want), name
        assert np.array_equal(d.cpu().numpy(), rgb)                   # input untouched
        out2 = dct.roundtrip_rgb(d)                                   # without the streams
        assert np.array_equal(out2.cpu().numpy(), want)


def test_rgb_tables_masks_and_views(dct, oracle):
    rgb = images((72, 200), 9)["noise"]
    for Q, Qc, keep in ((oracle.jpeg_Q() * 2, oracle.jpeg_Q_chroma(), dct.ALL_COEFFS),
                        (oracle.jpeg_Q(), oracle.jpeg_Q_chroma(), oracle.zigzag_mask(10)),
                        (oracle.jpeg_Q() * 0.37, oracle.jpeg_Q_chroma() * 1.3, oracle.zigzag_mask(21)),   # non-integer: IEEE division kernels
                        (np.full(64, 1.0, np.float32), np.full(64, 255.0, np.float32), dct.ALL_COEFFS)):
        want = oracle.roundtrip_rgb(rgb, Q=Q, Qc=Qc, keep=keep)
        plan = dct.Plan(Q=Q, Qc=Qc, keep=keep)
        assert np.array_equal(plan.chroma_quant(), np.asarray(Qc, np.float32))
        out = dct.roundtrip_rgb(torch.from_numpy(rgb).cuda(), plan=plan)
        assert np.array_equal(out.cpu().numpy(), want)
    # pitched views: a window of a larger image, written into a window; the frame keeps its sentinel
    big = torch.from_numpy(images((128, 256), 4)["smooth"]).cuda()
    frame = torch.full((128, 256, 3), 77, dtype=torch.uint8, device="cuda")
    src, dst = big[16:80, 32:160], frame[8:72, 64:192]
    dct.roundtrip_rgb(src, out=dst)
    assert np.array_equal(dst.cpu().numpy(), oracle.roundtrip_rgb(src.cpu().numpy()))
    mask = torch.ones(128, 256, dtype=torch.bool, device="cuda")
    mask[8:72, 64:192] = False
    assert bool((frame[mask] == 77).all())
    # in place
    x = big.clone()
    dct.roundtrip_rgb(x, out=x)
    assert np.array_equal(x.cpu().numpy(), oracle.roundtrip_rgb(big.cpu().numpy()))


def test_rgb_argument_errors(dct, oracle):
    rgb = torch.zeros(16, 16, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_rgb(rgb, plan=dct.Plan(T=oracle.dct2_T()))                    # Haweel's T only
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_rgb(torch.zeros(16, 12, 3, dtype=torch.uint8, device="cuda"))  # W % 8
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_rgb(torch.zeros(16, 16, 3, device="cuda"))                     # f32
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_rgb(rgb, out=torch.zeros(16, 8, 3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_rgb(rgb, streams=torch.zeros(3, 2, 2, 32, dtype=torch.int16, device="cuda"))


@pytest.mark.factored
def test_rgb_default_inverse_within_tolerance(dct, oracle):
    rgb = images((512, 512), 11)
    worst = 0
    for name, x in rgb.items():
        want = oracle.roundtrip_rgb(x)
        got = dct.roundtrip_rgb(torch.from_numpy(x).cuda()).cpu().numpy()
        diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
        worst = max(worst, int(diff.max()))
        assert diff.max() <= 3 and np.mean(diff > 0) < 0.01, name
        m_w, p_w = oracle.metrics(x.reshape(-1), want.reshape(-1))
        m_g, p_g = oracle.metrics(x.reshape(-1), got.reshape(-1))
        assert abs(m_g - m_w) <= 1e-3 * m_w and abs(p_g - p_w) <= 1e-3 * p_w


@pytest.mark.parametrize("shape", [(8, 8), (64, 96), (1024, 2048)])
def test_coded_bits_and_compression_factor(dct, oracle, shape):
    H, W = shape
    rng = np.random.default_rng(17)
    yy, xx = np.mgrid[0:H, 0:W]
    img = (128 + 90 * np.sin(xx / 17.0) * np.cos(yy / 23.0) + rng.normal(0, 6, (H, W))).clip(0, 255).astype(np.uint8)
    d = torch.from_numpy(img).cuda()
    for k in (64, 10, 6):
        keep = oracle.zigzag_mask(k)
        _, coef = oracle.roundtrip(img, keep=keep, want_coef=True)
        zz = dct.api.empty_zigzag(H, W, "cuda")
        dct.roundtrip(d, coef=zz, plan=dct.Plan(keep=keep), zigzag=True)
        for table in (0, 1):
            want = oracle.coded_bits(oracle.zigzag_i16(coef), table)
            assert dct.coded_bits(zz, table) == want
        assert abs(dct.compression_factor(zz) - oracle.compression_factor(coef)) < 1e-12
    # synthetic streams: long zero runs (ZRL), large magnitudes, last coefficient set, negative DC steps
    s = np.zeros((H // 8, W // 8, 64), np.int16)
    s[..., 0] = rng.integers(-1024, 1024, s.shape[:2])
    idx = rng.integers(1, 64, s.shape[:2])
    np.put_along_axis(s, idx[..., None], rng.integers(-1023, 1024, s.shape[:2] + (1,)).astype(np.int16), 2)
    s[0, 0, 63] = -1
    assert dct.coded_bits(torch.from_numpy(s).cuda(), 0) == oracle.coded_bits(s, 0)
    assert dct.coded_bits(torch.from_numpy(s).cuda(), 1) == oracle.coded_bits(s, 1)
    # colour: CF of the three streams together
    rgb = np.stack([img, np.roll(img, 5, 0), np.roll(img, 9, 1)], -1)
    zz3 = torch.empty(3, H // 8, W // 8, 64, dtype=torch.int16, device="cuda")
    dct.roundtrip_rgb(torch.from_numpy(rgb).cuda(), streams=zz3)
    _, coef3 = oracle.roundtrip_rgb(rgb, want_coef=True)
    bits = sum(oracle.coded_bits(oracle.zigzag_i16(coef3[c]), 0 if c == 0 else 1) for c in range(3))
    assert abs(dct.compression_factor(zz3) - 24.0 * H * W / bits) < 1e-12


def test_coded_bits_on_a_padded_stream_and_accumulation(dct, oracle):
    """The coded-size kernel walks the stream with its block-row pitch (DC prediction crosses the
    padding correctly) and ADDS into the caller's counter."""
    H, W = 40, 136
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    _, coef = oracle.roundtrip(img, want_coef=True)
    want = oracle.coded_bits(oracle.zigzag_i16(coef), 0)
    buf = torch.full((H // 8, W // 8 + 5, 64), 12345, dtype=torch.int16, device="cuda")
    view = buf[:, : W // 8, :]
    dct.roundtrip(torch.from_numpy(img).cuda(), coef=view, zigzag=True)
    assert dct.coded_bits(view, 0) == want
    acc = torch.full((1,), 1000, dtype=torch.int64, device="cuda")
    L = dct.lib()
    for _ in range(2):
        assert L.b200dct_zigzag_coded_bits(view.data_ptr(), view.stride(0) * 2, H, W, 0, acc.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert int(acc.item()) == 1000 + 2 * want
    # argument errors
    assert L.b200dct_zigzag_coded_bits(view.data_ptr(), 64, H, W, 0, acc.data_ptr(), None) != 0      # pitch too small
    assert L.b200dct_zigzag_coded_bits(view.data_ptr(), view.stride(0) * 2, H, W, 2, acc.data_ptr(), None) != 0  # no such table
