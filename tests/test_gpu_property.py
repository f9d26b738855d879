"""GPU property tests (hypothesis): random shapes, pitches, dtypes, masks, quantisers and
kernel families against the CPU oracle, bit-exact; plus structural properties that hold at
any size (block independence, batch == tall image, split == fused)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(hb=st.integers(1, 12), wb=st.integers(1, 80), pad=st.sampled_from([0, 4, 8, 32, 100]),
       u8=st.booleans(), path=st.sampled_from([0, 1]), k=st.sampled_from([64, 64, 10, 6, 1]),
       qscale=st.sampled_from([1.0, 1.0, 2.0, 0.5, 0.37]), with_coef=st.sampled_from(["none", "f32", "i16"]),
       seed=st.integers(0, 2 ** 16))
def test_random_configurations_match_oracle(dct, oracle, hb, wb, pad, u8, path, k, qscale, with_coef, seed):
    H, W = hb * 8, wb * 8
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8 if u8 else np.float32)
    if not u8 and seed % 3 == 0:
        img = (img + rng.random((H, W), dtype=np.float32)).astype(np.float32)   # non-integer pixels
    Q = (oracle.jpeg_Q() * qscale).astype(np.float32)
    keep = oracle.zigzag_mask(k)
    want_out, want_coef = oracle.roundtrip(img, Q=Q, keep=keep, want_coef=True)
    plan = dct.Plan(Q=Q, keep=keep, path=path)
    pad_elems = pad if not u8 else pad * 4          # keep rows 16-byte aligned in both dtypes
    big = torch.zeros(H, W + pad_elems, dtype=torch.uint8 if u8 else torch.float32, device="cuda")
    big[:, :W] = torch.from_numpy(img).cuda()
    outb = torch.full((H, W + pad_elems), 7, dtype=big.dtype, device="cuda")
    coef = None
    if with_coef != "none":
        coef = torch.empty(H, W, dtype=torch.float32 if with_coef == "f32" else torch.int16, device="cuda")
    dct.roundtrip(big[:, :W], out=outb[:, :W], coef=coef, plan=plan)
    got = host(outb)
    assert np.array_equal(got[:, :W] if u8 else bits(got[:, :W]), want_out if u8 else bits(want_out))
    assert (got[:, W:] == 7).all()                   # padding columns untouched
    if coef is not None:
        assert np.array_equal(host(coef).astype(np.float32), want_coef)


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(hb=st.integers(2, 40), wb=st.sampled_from([4, 32, 36, 64, 128]), cut=st.integers(1, 39), seed=st.integers(0, 999))
def test_stripes_batches_and_split_equal_fused(dct, hb, wb, cut, seed):
    H, W = hb * 8, wb * 8
    cut = min(cut, hb - 1) * 8
    g = torch.Generator(device="cuda").manual_seed(seed)
    img = torch.randint(0, 256, (H, W), device="cuda", generator=g, dtype=torch.int32).float()
    full = dct.roundtrip(img)
    top, bottom = dct.roundtrip(img[:cut]), dct.roundtrip(img[cut:])
    assert torch.equal(torch.cat([top, bottom]).view(torch.int32), full.view(torch.int32))
    rec = dct.inverse(dct.forward(img))
    assert torch.equal(rec.view(torch.int32), full.view(torch.int32))
    rec16 = dct.inverse(dct.forward(img, coef_dtype=torch.int16))
    assert torch.equal(rec16.view(torch.int32), full.view(torch.int32))
    batch = img.view(2, H // 2, W) if (H // 8) % 2 == 0 else None
    if batch is not None:
        assert torch.equal(dct.roundtrip(batch).view(torch.int32).view(H, W), full.view(torch.int32))


def test_time_calls_helper(dct, oracle):
    img = torch.from_numpy(oracle.rand_image(256, 256, 42)).cuda()
    out, coef = torch.empty_like(img), torch.empty_like(img)
    for which in ("roundtrip", "forward", "inverse", "split"):
        ms = dct.api.time_calls(which, img if which != "inverse" else coef, out, coef if which in ("roundtrip", "split") else None, iters=10)
        assert 0 < ms < 5
    dct.api.time_calls("split", img, out, coef, iters=3)
    assert np.array_equal(bits(host(out)), bits(oracle.roundtrip(host(img))))


def test_concurrent_streams_share_nothing(dct, oracle):
    """Launches on different streams use different scheduler slots: results stay exact."""
    imgs = [torch.from_numpy(oracle.rand_image(512, 512, s)).cuda() for s in range(4)]
    outs = [torch.empty_like(x) for x in imgs]
    streams = [torch.cuda.Stream() for _ in imgs]
    torch.cuda.synchronize()
    for rep in range(20):
        for x, y, s in zip(imgs, outs, streams):
            dct.roundtrip(x, out=y, stream=s)
    torch.cuda.synchronize()
    for i, (x, y) in enumerate(zip(imgs, outs)):
        assert np.array_equal(bits(host(y)), bits(oracle.roundtrip(oracle.rand_image(512, 512, i))))


@pytest.mark.parametrize("path,expect", [(0, "tma"), (1, "direct"), (2, "tma")])
def test_cuda_graph_capture(dct, oracle, path, expect):
    """A captured launch owns a ticket-counter pair of its own (left zeroed by every run), so the
    dynamically scheduled TMA kernels are capturable; replays stay exact, also when ordinary
    launches run in between."""
    N = 6144                                          # large enough for AUTO to pick TMA outside capture
    img = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float()
    out = torch.empty_like(img)
    plan = dct.Plan(path=path)
    dct.roundtrip(img, out=out, plan=plan)           # warm-up outside capture
    assert dct.api.last_path() == expect
    want = out.clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    slots_before = int(dct.lib().b200dct_capture_slots_left())
    with torch.cuda.graph(g, stream=s):
        dct.roundtrip(img, out=out, plan=plan, stream=torch.cuda.current_stream())
        captured_path = dct.api.last_path()
    assert captured_path == expect
    # a captured TMA launch keeps one scheduler slot for good; the count is visible to the caller
    assert int(dct.lib().b200dct_capture_slots_left()) == slots_before - (1 if expect == "tma" else 0)
    out.zero_()
    other = torch.empty_like(img)
    for _ in range(3):
        g.replay()
        dct.roundtrip(img, out=other, plan=plan)      # ordinary launches between replays (ring slots)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), want.view(torch.int32))
    assert torch.equal(other.view(torch.int32), want.view(torch.int32))
    band = host(img[:16])
    assert np.array_equal(bits(host(out[:16])), bits(oracle.roundtrip(band)))


def test_host_threads_are_reentrant(dct, oracle):
    """The C ABI is re-entrant: four host threads, each with its own plan and stream, hammer the
    library concurrently (ctypes releases the GIL); every result stays exact."""
    import threading

    imgs = [oracle.rand_image(1024, 1024, 100 + i) for i in range(4)]
    wants = [oracle.roundtrip(x, keep=oracle.zigzag_mask(6 + i)) for i, x in enumerate(imgs)]
    errors = []

    def worker(i):
        try:
            torch.cuda.set_device(0)
            plan = dct.Plan(keep=oracle.zigzag_mask(6 + i), path=1 + (i % 2))
            s = torch.cuda.Stream()
            x = torch.from_numpy(imgs[i]).cuda()
            y = torch.empty_like(x)
            for _ in range(50):
                dct.roundtrip(x, out=y, plan=plan, stream=s)
            s.synchronize()
            if not np.array_equal(bits(y.cpu().numpy()), bits(wants[i])):
                errors.append(f"thread {i}: mismatch")
        except Exception as e:  # pragma: no cover
            errors.append(f"thread {i}: {e!r}")

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_api_rejects_mismatched_planes(dct):
    """The C side trusts H, W and the pitch: the Python binding must refuse tensors that do not
    describe the same image instead of letting the kernels write out of bounds."""
    img = torch.zeros(64, 64, device="cuda")
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip(img, out=torch.empty(32, 64, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip(img, coef=torch.empty(64, 32, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip(img, out=torch.empty(64, 64, device="cuda", dtype=torch.uint8))
    with pytest.raises(dct.B200DCTError):
        dct.forward(img, coef=torch.empty(64, 128, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.forward(img, shifted=torch.empty(64, 64, device="cuda", dtype=torch.uint8))
    with pytest.raises(dct.B200DCTError):
        dct.forward(img, shifted=torch.empty(64, 128, device="cuda")[:, :64])   # other pitch
    with pytest.raises(dct.B200DCTError):
        dct.inverse(img, img=torch.empty(8, 64, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.metrics(img, torch.zeros(64, 32, device="cuda"))
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip(torch.zeros(3, 12, 64, device="cuda"))   # blocks would straddle the images of the batch
    with pytest.raises(dct.B200DCTError):
        dct.roundtrip_with_metrics(img, out=torch.empty(16, 64, device="cuda"))
    T = torch.zeros(64, device="cuda")
    with pytest.raises(dct.B200DCTError):     # would exit() inside the compat wrapper
        dct.dct_all_blocks_cuda(torch.zeros(100 * 100, device="cuda"), 100, 100, T, torch.zeros(100 * 100, device="cuda"))


def test_calls_on_a_side_stream(dct, oracle):
    """Allocation, zero-fill, launch and read-back of one call are ordered on the caller's stream
    (the metrics accumulators used to be zeroed on torch's current stream)."""
    img = oracle.rand_image(512, 512, 7)
    want = oracle.roundtrip(img)
    w_mse, w_peen = oracle.metrics(img, want)
    d = torch.from_numpy(img).cuda()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(5):
        # keep the current stream busy so that a zero-fill issued there would lose the race
        busy = torch.randn(4096, 4096, device="cuda") @ torch.randn(4096, 4096, device="cuda")
        out, (mse, peen, nnz) = dct.roundtrip_with_metrics(d, stream=side)
        side.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
        assert abs(mse - w_mse) <= 1e-6 * w_mse and abs(peen - w_peen) <= 1e-6 * w_peen
        m2, p2 = dct.metrics(d, out, stream=side)
        assert abs(m2 - w_mse) <= 1e-6 * w_mse and abs(p2 - w_peen) <= 1e-6 * w_peen
        out3 = dct.roundtrip(d, stream=side)
        side.synchronize()
        assert np.array_equal(out3.cpu().numpy().view(np.uint32), want.view(np.uint32))
        del busy


def test_early_loads_never_break_dependent_call_chains(dct, oracle):
    """Tile loads of a TMA-family launch may start before its predecessor has completed when the
    library has established that the predecessor does not write what this launch reads.  Chains in
    which it DOES (forward -> inverse through one coefficient plane, ping-pong round trips) must stay
    exact, launch after launch, and independent launches in between must not disturb them."""
    N = 6144                                   # >= 28 Mpixel: AUTO takes the TMA family, grid = every SM
    x = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32).float()
    y, z, c = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    plan = dct.Plan()
    want1 = dct.roundtrip(x, plan=plan).clone()
    want2 = dct.roundtrip(want1, plan=plan).clone()
    want3 = dct.roundtrip(want2, plan=plan).clone()
    torch.cuda.synchronize()
    band = x[:16].cpu().numpy()
    assert np.array_equal(bits(host(want1[:16])), bits(oracle.roundtrip(band)))
    for rep in range(25):
        dct.roundtrip(x, out=y, plan=plan)     # y <- f(x)
        dct.roundtrip(y, out=z, plan=plan)     # z <- f(y): reads what the previous launch wrote
        dct.roundtrip(z, out=y, plan=plan)     # y <- f(z): reads AND overwrites across the boundary
        assert dct.api.last_path() == "tma"
        dct.forward(x, coef=c, plan=plan)      # split API through one coefficient plane
        dct.inverse(c, img=z, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int32), want3.view(torch.int32))
    assert torch.equal(z.view(torch.int32), want1.view(torch.int32))
    # independent launches (rotating buffers): early loads apply; results identical
    outs = [torch.empty_like(x) for _ in range(3)]
    ins = [x, want1, want2]
    for rep in range(10):
        for i in range(3):
            dct.roundtrip(ins[i], out=outs[i], plan=plan)
    torch.cuda.synchronize()
    for o, w in zip(outs, (want1, want2, want3)):
        assert torch.equal(o.view(torch.int32), w.view(torch.int32))


@pytest.mark.parametrize("family", ["direct-u8", "direct-u8-ragged", "direct-f32", "tma-f32"])
def test_early_path_with_foreign_kernels_in_between(dct, oracle, family):
    """What the CALLER enqueues between two calls is invisible to the library's host side: a foreign
    kernel that writes the next call's input (here: torch copies / fills) must never be overtaken by the
    early loads of that call.  The device-side completion counter decides: a launch only loads early
    while its library predecessor is provably still running -- then nothing can stand between them.
    Also the direct family's version of dependent chains (reads what the predecessor wrote)."""
    N = 8192 if family != "tma-f32" else 6144   # direct family: 8192 CTAs, more than the machine holds at once
    M = N
    if family == "direct-u8-ragged":            # CTAs with exited threads (partial rows and a 1-lane last column) at the early path's barrier
        N, M = 8136, 8200
    dt = torch.uint8 if family.startswith("direct-u8") else torch.float32
    plan = dct.Plan(path=dct.api.PATH_TMA if family == "tma-f32" else dct.api.PATH_DIRECT, inverse=dct.api.INVERSE_EXACT)
    g = torch.Generator(device="cuda").manual_seed(11)
    src = [torch.randint(0, 256, (N, M), device="cuda", generator=g, dtype=torch.int32).to(dt) for _ in range(3)]
    want = [dct.roundtrip(a, plan=plan).clone() for a in src]
    want2 = dct.roundtrip(want[0], plan=plan).clone()
    torch.cuda.synchronize()
    band = src[0][:16].cpu().numpy()
    got = host(want[0][:16])
    assert np.array_equal(got, oracle.roundtrip(band)) if dt == torch.uint8 else np.array_equal(bits(got), bits(oracle.roundtrip(band)))
    x = torch.empty_like(src[0])
    outs = [torch.empty_like(x) for _ in range(2)]
    bad = 0
    for it in range(30):
        k = it % 3
        x.copy_(src[k])                                  # foreign kernel writes the input of the next call ...
        dct.roundtrip(x, out=outs[it % 2], plan=plan)    # ... whose predecessor in the library's books wrote outs[(it-1) % 2]
        if it % 5 == 4:
            x.fill_(0)                                   # and clobbers it again right behind the call
        bad += int(not torch.equal(outs[it % 2], want[k]))
    assert bad == 0
    # dependent chain: every launch reads what its predecessor wrote, and overwrites what that one read
    a, b = torch.empty_like(x), torch.empty_like(x)
    for rep in range(10):
        dct.roundtrip(src[0], out=a, plan=plan)
        dct.roundtrip(a, out=b, plan=plan)
        dct.roundtrip(src[0], out=a, plan=plan)          # overwrites what the previous launch is still reading
    torch.cuda.synchronize()
    assert torch.equal(a, want[0]) and torch.equal(b, want2)
    # independent launches back to back (the early path proper): identical results
    ins3, outs3 = src, [torch.empty_like(x) for _ in range(3)]
    for rep in range(10):
        for i in range(3):
            dct.roundtrip(ins3[i], out=outs3[i], plan=plan)
    torch.cuda.synchronize()
    for o, w in zip(outs3, want):
        assert torch.equal(o, w)


def test_randomised_programmes_async_equals_synchronised():
    """benchmarks/experiments/stress_early.py: random programmes of round trips (both families, in place
    and out of place, batches) interleaved with foreign kernels that write what the next calls read; run
    once with a device synchronisation after every operation and three times asynchronously -- every
    buffer must end up bit-identical."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("B200DCT_INVERSE", "B200DCT_DENSE")}   # library defaults
    r = subprocess.run([sys.executable, os.path.join(root, "benchmarks", "experiments", "stress_early.py"), "6", "100"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "STRESS ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("family", ["tma-f32", "direct-u8"])
def test_early_loads_with_threads_and_streams(dct, oracle, family):
    """The early path is decided per (device, stream) under one lock with the launch itself: two
    host threads hammering ONE stream with dependent pairs, and two streams chained by events, stay
    exact (TMA family: early tile loads; direct family: load + transform before the wait)."""
    import threading

    N = 6144 if family == "tma-f32" else 8192
    x = torch.randint(0, 256, (N, N), device="cuda", dtype=torch.int32)
    x = x.float() if family == "tma-f32" else x.to(torch.uint8)
    plan = dct.Plan(inverse=dct.api.INVERSE_EXACT)
    want1 = dct.roundtrip(x, plan=plan).clone()
    want2 = dct.roundtrip(want1, plan=plan).clone()
    torch.cuda.synchronize()
    # (1) one stream, two threads, each running its own dependent chain x -> a -> b
    s = torch.cuda.Stream()
    bufs = [(torch.empty_like(x), torch.empty_like(x)) for _ in range(2)]
    errs = []

    def work(i):
        try:
            a, b = bufs[i]
            for _ in range(15):
                dct.roundtrip(x, out=a, plan=plan, stream=s)
                dct.roundtrip(a, out=b, plan=plan, stream=s)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    s.synchronize()
    assert not errs
    for a, b in bufs:
        assert torch.equal(a.view(torch.int32), want1.view(torch.int32))
        assert torch.equal(b.view(torch.int32), want2.view(torch.int32))
    # (2) two streams chained by events: s2 consumes what s1 produced, s1 keeps producing into the other buffer
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    prod = [torch.empty_like(x) for _ in range(2)]
    cons = [torch.empty_like(x) for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    for it in range(12):
        k = it % 2
        if it >= 2:
            s1.wait_event(freed[k])
        dct.roundtrip(x, out=prod[k], plan=plan, stream=s1)
        done[k].record(s1)
        s2.wait_event(done[k])
        dct.roundtrip(prod[k], out=cons[k], plan=plan, stream=s2)
        freed[k].record(s2)
    torch.cuda.synchronize()
    for k in range(2):
        assert torch.equal(prod[k].view(torch.int32), want1.view(torch.int32))
        assert torch.equal(cons[k].view(torch.int32), want2.view(torch.int32))
