"""Jump-ahead of std::mt19937_64 (msweep_b200/csrc/mt64_jump.cu) against an independent Python MT19937-64 (tests/mt64.py).

The reference seeds one generator for all bootstrap replicates (src/BootstrapSample.cpp:48-60); with the replicates spread
over GPUs, mswb_bootstrap_run moves each rank to the start of its own replicates with this function.  Host arithmetic only:
runs without a GPU."""
import numpy as np
import pytest

from tests.mt64 import MT19937_64


def outputs_from(words, n):
    g = MT19937_64(0)
    g.mt = [int(x) for x in words]
    g.mti = 312
    return [g.next() for _ in range(n)]


def seeded(seed):
    return np.array(MT19937_64(seed).mt, np.uint64)


@pytest.mark.parametrize("J", [0, 1, 155, 156, 311, 312, 313, 624, 19937, 54321])
def test_jump_equals_stepping(mswb, J):
    st = seeded(7)
    g = MT19937_64(7)
    for _ in range(J):
        g.next()
    want = [g.next() for _ in range(700)]                       # crosses two refills
    assert outputs_from(mswb.mt64_jump(st, J), 700) == want


def test_jumps_compose_far_beyond_what_can_be_stepped(mswb):
    """z^a z^b = z^(a+b): a replicate stride of 1e7 draws taken 87 times equals one jump of 8.7e8, and
    a jump of 2^63 + 12345 equals its pieces."""
    st = seeded((-7) & ((1 << 64) - 1))
    D = 10_017_675
    a = st
    for _ in range(5):
        a = mswb.mt64_jump(a, D)
    b = mswb.mt64_jump(st, 5 * D)
    assert outputs_from(a, 400) == outputs_from(b, 400)
    far = mswb.mt64_jump(mswb.mt64_jump(st, 1 << 63), 12345)
    assert outputs_from(far, 400) == outputs_from(mswb.mt64_jump(st, (1 << 63) + 12345), 400)
    # and the words themselves agree apart from the 31 bits of word 0 that are not state
    assert np.array_equal(a[1:], b[1:]) and (int(a[0]) >> 31) == (int(b[0]) >> 31)


def test_libstdcxx_discard_agrees(mswb, oracle):
    """The oracle's bootstrap stream IS std::mt19937_64: replicate r of its resampler starts r x (number of draws) outputs
    in.  Rebuild replicate 3's draws from a jumped state and compare the resampled counts."""
    counts = np.array([5, 1, 0, 7, 2, 9, 4], np.uint32)
    total = int(counts.sum())
    want = oracle.bootstrap_resample(counts, 4242, 4)[3]
    words = mswb.mt64_jump(seeded(4242), 3 * total)
    cp = np.cumsum(counts.astype(np.float64) / float(counts.sum()))
    cp[-1] = 1.0
    got = np.zeros_like(counts)
    for u in outputs_from(words, total):
        p = float(u) * 2.0 ** -64
        p = np.nextafter(1.0, 0.0) if p >= 1.0 else p
        got[int(np.searchsorted(cp, p, side="left"))] += 1
    assert np.array_equal(got, want)
