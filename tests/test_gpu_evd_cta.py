"""GPU parity of the CTA-per-pixel phase_link kernel (fringe_b200/csrc/evd_cta.cu, 32 < bands <= 104) against the CPU
oracle (phase_link.cpp:479-666), through the C ABI.  Same gates as everywhere: sentinel codes equal, wrapped phase
<= 1e-3 rad where the oracle's temporal coherence > 0.3, |delta tcorr| <= 1e-4.

Covered: every instantiated order and its boundary, the three ways a pixel leaves the kernel -- EVD fall-back solved in
place (|C| not positive definite: more dates than SHPs), MLE branch handed to the warp-per-pixel kernel through the work
list (|C| positive definite: more SHPs than dates), sentinel -1 (a band that is zero in every SHP) -- compressed-SLC
offsets, line ranges, a wide window whose SHP list is staged in several chunks, and agreement with the generic kernel."""
import time

import numpy as np
import pytest

from conftest import wrapped_diff
from fringe_b200 import synth
from test_gpu_evd import _compare, _nmap

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bands", [33, 40, 48, 49, 64, 65, 80, 81, 92, 93, 100, 101, 104])
def test_every_order_fall_back_path(ctx, oracle_lib, bands):
    """11x5 window: at most 55 SHPs, so from ~56 dates on |C| is never positive definite (EVD fall-back in the CTA kernel);
    below that both paths occur."""
    slc = synth.make_stack(bands, 12, 36, seed=200 + bands, region=12)
    wts = _nmap(oracle_lib, slc, 5, 2)
    ref = oracle_lib.evd_block(slc, wts, 5, 2, method=1, variant=1, min_neighbors=5)
    gpu = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=1, min_neighbors=5)
    _compare(ref, gpu, borderline=3)
    st = ctx.evd_stats()
    print(f"bands {bands}: {st}")


def test_positive_definite_pixels_take_the_work_list(ctx, oracle_lib):
    """23x11 window, homogeneous scene: ~250 SHPs for 40 dates, |C| positive definite almost everywhere -- the MLE branch,
    solved by the warp-per-pixel kernel from the work list the CTA kernel leaves behind."""
    slc = synth.make_stack(40, 26, 48, seed=41, region=48)
    wts = _nmap(oracle_lib, slc, 11, 5)
    ref = oracle_lib.evd_block(slc, wts, 11, 5, method=1, variant=1, min_neighbors=5)
    gpu = ctx.evd_block(slc, wts, 11, 5, method="MLE", variant=1, min_neighbors=5)
    _compare(ref, gpu, borderline=3)
    ctx.force_generic(True)
    try:
        generic = ctx.evd_block(slc, wts, 11, 5, method="MLE", variant=1, min_neighbors=5)
    finally:
        ctx.force_generic(False)
    # the deferred pixels are solved by the very same code as in the generic run
    same = np.all(generic[0] == gpu[0], axis=0)
    print(f"identical to the generic kernel's output on {same.mean():.3f} of the pixels")
    assert same.mean() > 0.5


def test_agrees_with_the_generic_kernel_and_is_faster(ctx, oracle_lib):
    slc = synth.make_stack(100, 40, 160, seed=5)
    wts = _nmap(oracle_lib, slc, 5, 2)
    ref = oracle_lib.evd_block(slc, wts, 5, 2, method=1, variant=1, min_neighbors=5)
    t0 = time.perf_counter()
    fast = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=1, min_neighbors=5)
    t1 = time.perf_counter()
    st = ctx.evd_stats()
    ctx.force_generic(True)
    try:
        generic = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=1, min_neighbors=5)
    finally:
        ctx.force_generic(False)
    t2 = time.perf_counter()
    _compare(ref, fast, borderline=3)
    _compare(ref, generic, borderline=3)
    assert not np.array_equal(fast[0], generic[0])
    ok = (fast[1] > 0) & (generic[1] > 0)
    d = wrapped_diff(fast[0][:, ok], generic[0][:, ok]).max()
    print(f"CTA kernel {t1 - t0:.3f} s, generic {t2 - t1:.3f} s (host calls incl. copies); max phase difference between the two {d:.2e}; {st}")


def test_compressed_bands_and_line_range(ctx, oracle_lib):
    slc = synth.make_stack(70, 20, 40, seed=9, region=20)
    wts = _nmap(oracle_lib, slc, 5, 2)
    kw = dict(variant=1, min_neighbors=5, mini_stack_count=4)
    ref = oracle_lib.evd_block(slc, wts, 5, 2, method=1, **kw)
    gpu = ctx.evd_block(slc, wts, 5, 2, method="MLE", **kw)
    _compare(ref, gpu, borderline=3)
    assert np.all(gpu[0][3][gpu[1] > 0] == 1.0)
    part = ctx.evd_block(slc, wts, 5, 2, method="MLE", first_line=6, n_lines=9, **kw)
    assert np.array_equal(part[0][:, 6:15], gpu[0][:, 6:15]) and np.array_equal(part[1][6:15], gpu[1][6:15])
    assert np.array_equal(part[2][6:15], gpu[2][6:15])


def test_zero_band_sentinel(ctx, oracle_lib):
    slc = synth.make_stack(40, 30, 64, seed=117, region=16, zero_fraction=0.0)
    wts = _nmap(oracle_lib, slc, 5, 2)
    slc = slc.copy()
    slc[2, 8:22, 10:40] = 0
    ref = oracle_lib.evd_block(slc, wts, 5, 2, method=1, variant=1, min_neighbors=3)
    gpu = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=1, min_neighbors=3)
    assert (ref[1] == -1.0).sum() > 50
    assert np.array_equal(ref[1] == -1.0, gpu[1] == -1.0)
    _compare(ref, gpu, borderline=3)


def test_wide_window_staged_in_chunks(ctx, oracle_lib):
    """59x19 window (sequential.py's default), 36 dates: up to 1121 SHPs, staged ~20 at a time."""
    slc = synth.make_stack(36, 26, 64, seed=96, region=32)
    wts = _nmap(oracle_lib, slc, 29, 9)
    ref = oracle_lib.evd_block(slc, wts, 29, 9, method=1, variant=1, min_neighbors=5)
    gpu = ctx.evd_block(slc, wts, 29, 9, method="MLE", variant=1, min_neighbors=5)
    _compare(ref, gpu, borderline=3)
