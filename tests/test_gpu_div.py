"""GPU: the kernels' three-FMA division by a known divisor equals div.rn.f32
(utils_kernels.cu:42) -- swept over EVERY float bit pattern for each JPEG divisor, and
over a dense sample for every integer divisor 1..255 (the set the library accepts for the
fast path; anything else uses __fdiv_rn)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sweep(dct, d, first, count):
    acc = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = dct.lib().b200dct_selftest_division(float(d), first, count, acc.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    return acc.tolist()


def test_exhaustive_for_jpeg_divisors(dct, oracle):
    for d in sorted(set(oracle.jpeg_Q().tolist())):
        bad_q, bad_c = sweep(dct, d, 0, 1 << 32)
        assert bad_c == 0, f"quantised value differs for divisor {d}: {bad_c} inputs"
        assert bad_q == 0, f"quotient bits differ for divisor {d} at |x| >= 2^-120: {bad_q} inputs"


def test_all_integer_divisors_1_255(dct):
    # normal-range dividends 2^-20 .. 2^20 of both signs: every exponent the transform can produce
    lo, hi = np.float32(2.0 ** -20).view(np.uint32), np.float32(2.0 ** 20).view(np.uint32)
    for d in range(1, 256):
        for sign in (0, 0x80000000):
            bad_q, bad_c = sweep(dct, d, int(lo) | sign, int(hi) - int(lo))
            assert bad_q == 0 and bad_c == 0, (d, sign, bad_q, bad_c)


@pytest.mark.slow
def test_exhaustive_all_integer_divisors(dct):
    for d in range(1, 256):
        bad_q, bad_c = sweep(dct, d, 0, 1 << 32)
        assert bad_c == 0 and bad_q == 0, (d, bad_q, bad_c)
