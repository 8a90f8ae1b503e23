"""The identity the KS2 device test rests on (fringe_b200/csrc/nmap_kernels.cu, ks_bad4):

    max_v |#{a <= v} - #{b <= v}| <= k   <=>   b[i-k] <= a[i] and a[i-k] <= b[i] for all i in [k, n)

for ascending a, b of equal length n, ties included.  Checked exhaustively on small alphabets
(many ties) and on random draws, and against the oracle's KS statistic (KS2sample.hpp walk)."""
import itertools

import numpy as np
import pytest


def max_count_diff(a, b):
    v = np.union1d(a, b)
    ca = np.searchsorted(a, v, side="right")
    cb = np.searchsorted(b, v, side="right")
    return int(np.abs(ca - cb).max())


def within(a, b, k):
    n = len(a)
    if k >= n:
        return True
    return bool(np.all(b[: n - k] <= a[k:]) and np.all(a[: n - k] <= b[k:]))


def test_exhaustive_small_alphabet():
    n = 5
    combos = list(itertools.combinations_with_replacement(range(4), n))
    for a in combos:
        a = np.array(a)
        for b in combos:
            b = np.array(b)
            d = max_count_diff(a, b)
            for k in range(0, n + 1):
                assert within(a, b, k) == (d <= k), (a, b, k, d)


@pytest.mark.parametrize("n", [2, 7, 30, 100])
def test_random(n):
    rng = np.random.default_rng(n)
    for trial in range(400):
        hi = rng.choice([3, 10, 1000, 1 << 30])
        a = np.sort(rng.integers(0, hi, n))
        b = np.sort(rng.integers(0, hi, n) + rng.integers(0, max(1, hi // 4)))
        d = max_count_diff(a, b)
        for k in {0, 1, d - 1, d, d + 1, n - 1, n}:
            if k < 0:
                continue
            assert within(a, b, k) == (d <= k)


@pytest.mark.parametrize("n,pvalue", [(30, 0.05), (20, 0.05), (12, 0.2), (64, 0.01)])
def test_against_oracle_probability(n, pvalue):
    """Reference decision (KS2sample.hpp probability >= threshold, via the oracle) == the
    inequality test with the host-computed bound (fringe_ks2_critical_count, no GPU needed)."""
    import ctypes as C

    from fringe_b200._lib import lib
    from oracle import load
    orc = load()
    k = C.c_int(0)
    assert lib.fringe_ks2_critical_count(n, pvalue, C.byref(k), None) == 0
    rng = np.random.default_rng(n)
    accepted = 0
    for trial in range(300):
        a = np.sort(rng.rayleigh(1.0, n).astype(np.float32))
        b = np.sort((rng.rayleigh(1.0, n) * rng.choice([1.0, 1.3, 2.0])).astype(np.float32))
        if trial % 5 == 0:
            b[: n // 3] = a[: n // 3]            # ties
            b.sort()
        ref = orc.ks2_prob(a, b) >= pvalue
        assert within(a, b, k.value) == ref
        accepted += ref
    assert 0 < accepted < 300
