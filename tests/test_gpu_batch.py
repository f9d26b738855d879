"""b200dct_roundtrip_batch: a list of separately allocated images of one shape in one launch per 64
images (BASELINE configs[4]: "a batch of 64 8192^2 images"; the reference runs one image per program,
main_newAppr.cu:99,120).  Criterion: every image equals the CPU oracle -- and therefore the single-image
entry point -- BIT FOR BIT (coefficients never leave the kernel here; pixels are compared: f32 as bit
patterns, u8 with the exact inverse; the library-default factored inverse within 1 LSB)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


@pytest.mark.parametrize("n", [1, 3, 64, 65, 130])
@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_batch_equals_oracle_per_image(dct, oracle, n, dtype):
    H, W = 40, 96 + 8 * (n % 3)   # partial CTAs in both directions, partial warps at the right edge
    gen = oracle.rand_image if dtype == "f32" else oracle.rand_image_u8
    imgs = [gen(H, W, 100 + k) for k in range(n)]
    # separately allocated, with unrelated tensors in between so that the images are not back to back
    d, pad = [], []
    for a in imgs:
        d.append(dev(a))
        pad.append(torch.empty(1000 + 8 * len(pad), dtype=torch.uint8, device="cuda"))
    outs = dct.roundtrip_batch(d)
    assert dct.api.last_path() == "direct"
    assert dct.api.last_launch_count() == (n + 63) // 64
    assert len(outs) == n
    for k in range(n):
        want = oracle.roundtrip(imgs[k])
        got = host(outs[k])
        if dtype == "f32":
            assert np.array_equal(bits(got), bits(want)), f"image {k}"
        else:
            assert np.array_equal(got, want), f"image {k}"
        assert np.array_equal(host(d[k]), imgs[k])   # inputs untouched


def test_batch_in_place_pitched_and_masked(dct, oracle):
    """Views inside larger tensors (shared pitch), results written over the inputs, a retained-coefficient
    mask as runtime data (k = 5) and as a compile-time kernel (k = 8), a custom quantiser."""
    H, W, n = 64, 128, 5
    for plan, kw in ((dct.Plan(keep=dct.zigzag_mask(5)), dict(keep=dct.zigzag_mask(5))),
                     (dct.Plan(keep=dct.zigzag_mask(8)), dict(keep=dct.zigzag_mask(8))),
                     (dct.Plan(Q=np.full(64, 7.0, np.float32)), dict(Q=np.full(64, 7.0, np.float32)))):
        imgs = [oracle.rand_image(H, W, 7 + k) for k in range(n)]
        big = [torch.full((H + 16, W + 64), -7.0, device="cuda") for _ in range(n)]
        views = [b[8:8 + H, 32:32 + W] for b in big]
        for v, a in zip(views, imgs):
            v.copy_(dev(a))
        outs = dct.roundtrip_batch(views, outs=views, plan=plan)
        for k in range(n):
            want = oracle.roundtrip(imgs[k], **kw)
            assert np.array_equal(bits(host(outs[k])), bits(want)), f"image {k}"
            full = host(big[k])
            full[8:8 + H, 32:32 + W] = -7.0
            assert (full == -7.0).all()      # nothing outside the views was written


@pytest.mark.factored
def test_batch_default_u8_inverse_within_one_lsb(dct, oracle):
    n, H, W = 9, 256, 256
    imgs = [oracle.rand_image_u8(H, W, 3 * k) for k in range(n)]
    outs = dct.roundtrip_batch([dev(a) for a in imgs])
    single = [dct.roundtrip(dev(a)) for a in imgs]
    for k in range(n):
        want = oracle.roundtrip(imgs[k]).astype(np.int32)
        got = host(outs[k]).astype(np.int32)
        assert np.abs(got - want).max() <= 1
        assert np.array_equal(host(outs[k]), host(single[k]))   # same kernel, same bits as the single-image call


def test_batch_dense_transform(dct, oracle):
    k, n = np.mgrid[0:8, 0:8]
    T = (np.where(k == 0, np.sqrt(1 / 8), np.sqrt(2 / 8)) * np.cos((2 * n + 1) * k * np.pi / 16)).astype(np.float32)
    imgs = [oracle.rand_image(48, 72, 11 + i) for i in range(4)]
    outs = dct.roundtrip_batch([dev(a) for a in imgs], plan=dct.Plan(T=T))
    for i in range(4):
        assert np.array_equal(bits(host(outs[i])), bits(oracle.roundtrip(imgs[i], T=T)))


def test_batch_argument_errors(dct):
    a = torch.zeros(64, 64, device="cuda")
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_batch([a, torch.zeros(64, 72, device="cuda")])          # shapes differ
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_batch([a, torch.zeros(64, 64, device="cuda", dtype=torch.uint8)])   # dtypes differ
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_batch([a, a], outs=[a])                                 # one output per image
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_batch([torch.zeros(60, 64, device="cuda")])             # not a multiple of 8
    assert dct.roundtrip_batch([]) == []


def test_batch_in_a_captured_graph(dct, oracle):
    imgs = [oracle.rand_image(64, 64, 50 + k) for k in range(70)]
    d = [dev(a) for a in imgs]
    outs = [torch.empty_like(t) for t in d]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dct.roundtrip_batch(d, outs=outs, stream=s)   # warm-up outside the capture
        s.synchronize()
        for o in outs:
            o.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            dct.roundtrip_batch(d, outs=outs, stream=s)
        g.replay()
    s.synchronize()
    for k in range(70):
        assert np.array_equal(bits(host(outs[k])), bits(oracle.roundtrip(imgs[k])))


def test_batch_of_full_size_images_matches_single_calls(dct):
    """Four separately allocated 4096^2 f32 images and four 8192^2 u8 images: the batch launch against
    the single-image entry point (whatever family AUTO picks for it), whole images, bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(5)
    # 8192^2 f32 images are large enough for the persistent TMA kernels: launched one by one on that family
    for N, dt, path, launches in ((4096, torch.float32, "direct", 1), (8192, torch.uint8, "direct", 1), (8192, torch.float32, "tma", 4)):
        imgs = [torch.randint(0, 256, (N, N), device="cuda", generator=g, dtype=torch.int32).to(dt) for _ in range(4)]
        outs = dct.roundtrip_batch(imgs)
        assert (dct.api.last_path(), dct.api.last_launch_count()) == (path, launches)
        for a, b in zip(imgs, outs):
            ref = dct.roundtrip(a)
            assert torch.equal(ref.view(torch.int32) if dt == torch.float32 else ref, b.view(torch.int32) if dt == torch.float32 else b)
