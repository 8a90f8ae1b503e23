"""CPU: the C-ABI library loads, exports every symbol include/fringe_b200.h declares, and its
host-side threshold arithmetic (no device needed) agrees with the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from fringe_b200 import _lib, engine


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 20 and "fringe_nmap_block" in names and "fringe_evd_block_device" in names
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n
    assert _lib.lib.fringe_abi_version() == 2
    # profiling header: context-bound hooks live in the main library, the microbenchmarks in their own
    prof = C.CDLL(_lib.PROF_LIB_PATH)
    for n in _lib.declared_symbols(_lib.PROF_HEADER_PATH):
        assert hasattr(prof if n.startswith("fringe_prof_") and n != "fringe_prof_force_generic" else raw, n), n
    # nothing measurement-related is left in the drop-in header
    assert not [n for n in names if n.endswith(("_peak", "_rate", "_stats", "_cycles", "_kernel_ms")) or "prof" in n]


def test_nmap_cuda_h_shim_is_exported_with_the_reference_linkage():
    """src/nmap/nmap_cuda.h:13-17 declares plain C++ functions: the library must export exactly those mangled names
    so that nmap.cpp built with -DBUILD_NMAP_WITH_CUDA links against it."""
    raw = C.CDLL(_lib.LIB_PATH)
    for sym in ("_Z16nmapProcessBlockPfPhiiiPiPjidii", "_Z7lockGPUv", "_Z9unlockGPUv"):
        assert hasattr(raw, sym), sym


def test_no_torch_or_oracle_in_the_product_library():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "openblas" not in out


def test_status_strings_and_nulong():
    assert _lib.lib.fringe_status_string(0) == b"ok"
    assert b"CPU fallback" in _lib.lib.fringe_status_string(_lib.ERR_NO_DEVICE)
    for Nx, Ny in [(5, 2), (5, 5), (10, 10), (11, 5), (0, 0)]:
        assert engine.nulong(Nx, Ny) == oracle.nulong(Nx, Ny)


def test_no_cpu_fallback_without_a_device():
    if engine.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.FringeError) as e:
        engine.Context(0)
    assert e.value.status == _lib.ERR_NO_DEVICE


def test_ks2_critical_counts():
    # SURVEY.md appendix A.2 (computed there with the reference's own double code)
    expect = {10: 6, 11: 6, 15: 7, 20: 8, 29: 10, 30: 10, 59: 14, 100: 19}
    o = oracle.load()
    for n, k in expect.items():
        kc, margin = engine.ks2_critical_count(n, 0.05)
        assert kc == k
        assert margin > 1e-3          # a few ulp of drift in the accumulated distance cannot flip it
        # the critical count really is the boundary of the reference's p-value
        assert o.kolmogorov_prob(kc / n * np.sqrt(n / 2)) >= 0.05 > o.kolmogorov_prob((kc + 1) / n * np.sqrt(n / 2))
    assert engine.ks2_critical_count(30, 0.0)[0] == 30      # everything accepted
    assert engine.ks2_critical_count(30, 1.5)[0] == -1      # nothing accepted


def _ks_walk(a, b):
    """numpy mirror of ks_max_count_diff() in nmap_kernels.cu"""
    n = len(a)
    A = np.append(a, np.inf); B = np.append(b, np.inf)
    ia = ib = kmax = 0
    va, vb = A[0], B[0]
    for _ in range(2 * n):
        ta = (ib >= n) or (ia < n and va <= vb)
        x = va if ta else vb
        if ta:
            ia += 1; va = A[ia]
        else:
            ib += 1; vb = B[ib]
        if va > x and vb > x:
            kmax = max(kmax, abs(ib - ia))
    return kmax


def _ad_walk(a, b, T):
    """numpy mirror of ad_inner_sum() in nmap_kernels.cu"""
    n = len(a)
    A = np.append(a, np.inf); B = np.append(b, np.inf)
    ia = ib = 0
    va, vb = A[0], B[0]
    S = 0.0
    for j in range(2 * n - 1):
        ta = (ib >= n) or (ia < n and va < vb)
        if ta:
            ia += 1; va = A[ia]
        else:
            ib += 1; vb = B[ib]
        S = S + T[j, abs(2 * ia - (j + 1))]
    return S


def _ad_table(n):
    L = 2 * n
    T = np.zeros((L - 1, n + 1))
    for j in range(L - 1):
        bj = j + 1.0
        for u in range(n + 1):
            t = float(n) * u
            T[j, u] = t * t / (bj * (L - bj))
    return T


@pytest.mark.parametrize("n", [5, 10, 20, 30])
@pytest.mark.parametrize("pvalue", [0.01, 0.05, 0.3])
def test_integer_ks_and_table_ad_decisions_equal_reference_pvalue_test(n, pvalue):
    """The device kernels never evaluate a p-value: KS compares an integer with
    fringe_ks2_critical_count, AD compares a table sum with fringe_ad2_critical_sum.  Both
    decisions must equal `p >= threshold` of the reference on every pair, ties included."""
    o = oracle.load()
    rng = np.random.default_rng(100 * n + int(pvalue * 100))
    kc, _ = engine.ks2_critical_count(n, pvalue)
    sc = engine.ad2_critical_sum(n, pvalue)
    assert engine.ad2_sigma(n) == o.ad2_sigma(n)
    T = _ad_table(n)
    for t in range(300):
        a = np.sort(rng.rayleigh(rng.choice([1.0, 1.0, 1.5, 2.5]), n)).astype(np.float32)
        b = np.sort(rng.rayleigh(rng.choice([1.0, 1.0, 1.5, 2.5]), n)).astype(np.float32)
        if t % 3 == 0:
            a = np.sort(np.round(a * 4) / 4 + 0.25).astype(np.float32)
            b = np.sort(np.round(b * 4) / 4 + 0.25).astype(np.float32)
        assert (_ks_walk(a, b) <= kc) == (o.ks2_prob(a, b) >= pvalue)
        assert (_ad_walk(a, b, T) <= sc) == (o.ad2_prob(a, b) >= pvalue)
