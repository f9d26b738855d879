"""GPU: the UNMODIFIED reference kernels (oracle/_ref, compiled from /root/reference by
oracle/Makefile) against (a) the CPU oracle -- this is what pins the oracle -- and
(b) the new CUDA path.  Bit-exact for HpApprDCT and fastApprDCT.  The two cuBLAS variants
accumulate in an order cuBLAS does not document, so for them the mismatch COUNT is
reported and bounded, pixels within +-1 LSB."""
import os

import numpy as np
import pytest
import torch

import inputs
import refgpu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def need(variant):
    if not refgpu.available(variant):
        pytest.skip(f"oracle/_ref/{refgpu.VARIANTS[variant]} not built (needs /root/reference at build time)")


@pytest.mark.parametrize("variant", ["newappr", "fastappr"])
@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
def test_reference_kernels_pin_the_oracle(oracle, dct, variant, N):
    need(variant)
    img = oracle.rand_image(N, N, 42)
    T = dev(oracle.haweel_T())
    refgpu.set_quant(variant, oracle.jpeg_Q())
    d_img = dev(img)
    ref_coef, _ = refgpu.dct(variant, d_img, T)
    ref_rec, _ = refgpu.idct(variant, ref_coef, T)
    want_coef, want_shift = oracle.dct(img, want_shifted=True)
    # (a) oracle == reference, bit for bit (including the sign of zero coefficients)
    assert np.array_equal(bits(host(ref_coef)), bits(want_coef))
    assert np.array_equal(host(d_img), want_shift)                   # input left as image-128
    assert np.array_equal(bits(host(ref_rec)), bits(oracle.idct(want_coef)))
    # (b) new CUDA path == reference
    d2 = dev(img)
    coef = torch.empty_like(d2)
    out = dct.roundtrip(d2, coef=coef)
    assert torch.equal(coef.view(torch.int32), ref_coef.view(torch.int32))
    assert torch.equal(out.view(torch.int32), ref_rec.view(torch.int32))
    # MSE / PEEN of the u8 result agree (spec: 1e-3 relative)
    u8 = d2.to(torch.uint8)
    m_new = dct.metrics(u8, out.clamp(0, 255).to(torch.uint8))
    m_ref = dct.metrics(u8, ref_rec.clamp(0, 255).to(torch.uint8))
    assert m_new[0] == pytest.approx(m_ref[0], rel=1e-12) and m_new[1] == pytest.approx(m_ref[1], rel=1e-12)


def test_headline_size_whole_image_against_the_reference_kernel(oracle, dct):
    """BASELINE configs[1] compared WHOLE: every coefficient and every f32 pixel of the 8192^2
    image against the unmodified reference kernels (3.3 ms on the box), plus SURVEY.md
    Appendix B's N=8192 known answers on the GPU results."""
    need("newappr")
    N = 8192
    img = oracle.rand_image(N, N, 42)                       # srand(42); rand()%256
    T = dev(oracle.haweel_T())
    refgpu.set_quant("newappr", oracle.jpeg_Q())
    work = dev(img)
    ref_coef, _ = refgpu.dct("newappr", work, T)
    ref_rec, _ = refgpu.idct("newappr", ref_coef, T)
    d = dev(img)
    for path in (2, 1):
        plan = dct.Plan(path=path)
        coef = torch.empty_like(d)
        out = dct.roundtrip(d, coef=coef, plan=plan)
        assert torch.equal(coef.view(torch.int32), ref_coef.view(torch.int32)), path
        assert torch.equal(out.view(torch.int32), ref_rec.view(torch.int32)), path
        del coef
    # the split entry points (the drop-in two-call API) as well
    c2 = dct.forward(d)
    assert torch.equal(c2.view(torch.int32), ref_coef.view(torch.int32))
    assert torch.equal(dct.inverse(c2).view(torch.int32), ref_rec.view(torch.int32))
    # Appendix B, N = 8192
    assert int(ref_coef.double().sum().item()) == -268447
    assert int(ref_coef.abs().double().sum().item()) == 117045095
    assert int((ref_coef != 0).sum().item()) == 47462419
    u8 = out.clamp(0, 255).to(torch.uint8)
    assert int(u8.long().sum().item()) == 8524699637
    mse, peen = dct.metrics(d.to(torch.uint8), u8)
    assert mse == pytest.approx(344.379744306, rel=1e-9) and peen == pytest.approx(12.593068395, rel=1e-9)


def test_reference_on_adversarial_and_rectangular(oracle, dct):
    need("newappr")
    T = dev(oracle.haweel_T())
    refgpu.set_quant("newappr", oracle.jpeg_Q())
    for img in (inputs.adversarial(32), inputs.float_noise(64, 256), oracle.rand_image(48, 640, 7)):
        d_img = dev(img)
        ref_coef, _ = refgpu.dct("newappr", d_img, T)
        ref_rec, _ = refgpu.idct("newappr", ref_coef, T)
        assert np.array_equal(bits(host(ref_coef)), bits(oracle.dct(img)))
        assert np.array_equal(bits(host(ref_rec)), bits(oracle.idct(oracle.dct(img))))
        out = dct.roundtrip(dev(img))
        assert torch.equal(out.view(torch.int32), ref_rec.view(torch.int32))


def test_reference_custom_quant(oracle, dct):
    need("newappr")
    img = oracle.rand_image(256, 256, 5)
    Q = oracle.jpeg_Q() * 0.5
    T = dev(oracle.haweel_T())
    assert refgpu.set_quant("newappr", Q) == 0
    ref_coef, _ = refgpu.dct("newappr", dev(img), T)
    refgpu.set_quant("newappr", oracle.jpeg_Q())
    assert np.array_equal(bits(host(ref_coef)), bits(oracle.dct(img, Q=Q)))
    assert torch.equal(dct.forward(dev(img), plan=dct.Plan(Q=Q)).view(torch.int32), ref_coef.view(torch.int32))


@pytest.mark.parametrize("variant", ["cublas2", "cublas"])
def test_cublas_variants_mismatch_count(oracle, dct, variant):
    """cuBLAS's k-accumulation order is opaque: report how many coefficients differ from
    the FMA-chain result, require pixels within +-1 LSB and MSE/PEEN within 1e-3."""
    need(variant)
    N = 256
    img = oracle.rand_image(N, N, 42)
    refgpu.set_quant(variant, oracle.jpeg_Q())
    for name, Tm in (("haweel", oracle.haweel_T()), ("dct2", oracle.dct2_T())):
        T = dev(Tm)
        d_img = dev(img)
        ref_coef, _ = refgpu.dct(variant, d_img, T)
        keep = ref_coef.clone()
        ref_rec, _ = refgpu.idct(variant, ref_coef, T)     # v2 dequantises ref_coef in place
        plan = dct.Plan(T=Tm)
        coef = dct.forward(dev(img), plan=plan)
        rec = dct.inverse(keep, plan=plan)                  # same coefficients in: isolates the inverse
        n_diff = int((coef != keep).sum())
        max_cdiff = float((coef - keep).abs().max())
        pix_diff = float((rec - ref_rec).abs().max())
        print(f"[{variant}/{name}] coefficient mismatches vs cuBLAS: {n_diff}/{N * N} (max |diff| {max_cdiff}); "
              f"max pixel diff for identical coefficients: {pix_diff:.3e}")
        assert max_cdiff <= 1 and n_diff <= N * N * 2e-3
        assert pix_diff < 1e-3                               # float reassociation noise only
        u8 = dev(img).to(torch.uint8)
        a = dct.metrics(u8, rec.clamp(0, 255).to(torch.uint8))
        b = dct.metrics(u8, ref_rec.clamp(0, 255).to(torch.uint8))
        assert a[0] == pytest.approx(b[0], rel=1e-3) and a[1] == pytest.approx(b[1], rel=1e-3)
        assert int((rec.clamp(0, 255).to(torch.uint8).int() - ref_rec.clamp(0, 255).to(torch.uint8).int()).abs().max()) <= 1


@pytest.mark.factored
@pytest.mark.parametrize("N", [1024, 2048])
def test_cublas2_mismatch_counts_at_larger_sizes(oracle, dct, N):
    """cublasDCTv2 (whole-image Sgemm, main_cublass_2.cu:228-235,288-295) at 1024^2 and 2048^2 with
    the true DCT-II matrix: mismatch count of the quantised coefficients for BOTH dense modes
    (ordered chains, and the default even/odd evaluation) -- the even/odd kernel must be no
    further from cuBLAS than the chain kernel is, pixels within 1 LSB, MSE/PEEN within 1e-3."""
    need("cublas2")
    img = oracle.rand_image(N, N, 42)
    Tm = oracle.dct2_T()
    refgpu.set_quant("cublas2", oracle.jpeg_Q())
    T = dev(Tm)
    ref_coef, _ = refgpu.dct("cublas2", dev(img), T)
    keep = ref_coef.clone()
    ref_rec, _ = refgpu.idct("cublas2", ref_coef, T)       # dequantises ref_coef in place
    counts = {}
    for name, mode in (("chain", dct.api.DENSE_CHAIN), ("symmetric", dct.api.DENSE_AUTO)):
        plan = dct.Plan(T=Tm, dense=mode)
        coef = dct.forward(dev(img), plan=plan)
        counts[name] = int((coef != keep).sum())
        assert float((coef - keep).abs().max()) <= 1
        rec = dct.inverse(keep, plan=plan)
        assert float((rec - ref_rec).abs().max()) < 1e-3
        u8 = dev(img).to(torch.uint8)
        a = dct.metrics(u8, rec.clamp(0, 255).to(torch.uint8))
        b = dct.metrics(u8, ref_rec.clamp(0, 255).to(torch.uint8))
        assert a[0] == pytest.approx(b[0], rel=1e-3) and a[1] == pytest.approx(b[1], rel=1e-3)
        assert int((rec.clamp(0, 255).to(torch.uint8).int() - ref_rec.clamp(0, 255).to(torch.uint8).int()).abs().max()) <= 1
    # the tensor-core arm (fused round trip only): same criterion against the live cuBLAS result
    cm = torch.empty(N, N, device="cuda")
    om = dct.roundtrip(dev(img), coef=cm, plan=dct.Plan(T=Tm, dense=dct.api.DENSE_MMA))
    assert dct.api.last_path() == "mma"
    counts["mma"] = int((cm != keep).sum())
    assert float((cm - keep).abs().max()) <= 1
    u8 = dev(img).to(torch.uint8)
    a = dct.metrics(u8, om.clamp(0, 255).to(torch.uint8))
    b = dct.metrics(u8, ref_rec.clamp(0, 255).to(torch.uint8))
    assert a[0] == pytest.approx(b[0], rel=1e-3) and a[1] == pytest.approx(b[1], rel=1e-3)
    same = (cm == keep).view(N // 8, 8, N // 8, 8).all(dim=3).all(dim=1)                # blocks with identical coefficients
    pd = (om - ref_rec).abs().view(N // 8, 8, N // 8, 8).amax(dim=3).amax(dim=1)
    assert float(pd[same].max()) < 2e-3
    print(f"[cublas2/dct2 {N}^2] coefficient mismatches vs live cuBLAS: chain {counts['chain']}, "
          f"even/odd {counts['symmetric']}, tensor-core arm {counts['mma']} of {N * N}")
    assert counts["symmetric"] <= max(2 * counts["chain"], 1e-3 * N * N)
    assert counts["mma"] <= 2e-3 * N * N


def test_committed_reference_fixtures_match_live_reference(oracle):
    """tests/golden/refgpu_*.npz were produced by these same reference kernels."""
    need("newappr")
    import glob

    files = sorted(glob.glob(os.path.join(GOLD, "refgpu_*.npz")))
    if not files:
        pytest.skip("no refgpu fixtures committed yet")
    T = dev(oracle.haweel_T())
    refgpu.set_quant("newappr", oracle.jpeg_Q())
    for f in files:
        g = np.load(f)
        coef, _ = refgpu.dct("newappr", dev(g["img"].astype(np.float32)), T)
        assert np.array_equal(bits(host(coef)), bits(g["coef"]))


def test_cpp_caller_links_against_compat_library(oracle):
    """tests/cpp/dropin_main.cu is written like the reference's programs (forward-declared
    dct_all_blocks_cuda / idct_all_blocks_cuda, (height, width) order, device buffers) and is
    linked against libb200dct_compat.so: it must print the oracle's (= the reference's) numbers."""
    import re
    import subprocess

    exe = os.path.join(os.path.dirname(GOLD.rstrip("/")), "..", "cuda-dct-idct_b200", "dropin_demo")
    exe = os.path.abspath(exe)
    if not os.path.exists(exe):
        pytest.skip("dropin_demo not built")
    for variant in (0, 1):
        out = subprocess.run([exe, "256", str(variant)], capture_output=True, text=True, timeout=120).stdout
        m = re.search(r"COEF sum=(-?\d+) sumabs=(\d+) nonzero=(\d+)", out)
        assert m, out
        assert tuple(map(int, m.groups())) == (-306, 114414, 46317)         # SURVEY.md Appendix B, N=256
        assert int(re.search(r"PIX sumu8=(\d+)", out).group(1)) == 8329775
        assert "DCT (256,256):" in out and "IDCT (256,256):" in out          # the reference's timing lines
        assert "INPUT_AFTER first=-58 (was 70)" in out                       # input left as image-128
