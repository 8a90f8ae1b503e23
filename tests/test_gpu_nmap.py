"""GPU parity: SHP selection through the C ABI (fringe_nmap_block) vs the CPU oracle.
Gate: neighbour count and bitmask words bit-exact (BASELINE.json north_star)."""
import numpy as np
import pytest

from fringe_b200 import synth

pytestmark = pytest.mark.gpu


def _check(ctx, oracle_lib, slc, Nx, Ny, method, pvalue=0.05, mask=None, alpha=None):
    mcode = {"KS2": 0, "AD2": 1}[method]
    c_ref, w_ref = oracle_lib.nmap_block(slc, Nx, Ny, method=mcode, thresh=pvalue, mask=mask, alpha=alpha)
    c_gpu, w_gpu = ctx.nmap_block(slc, Nx, Ny, method=method, pvalue=pvalue, mask=mask, alpha=alpha)
    assert np.array_equal(c_gpu, c_ref), f"count mismatches: {(c_gpu != c_ref).sum()}"
    assert np.array_equal(w_gpu, w_ref), f"mask word mismatches: {(w_gpu != w_ref).sum()}"
    return c_ref


def test_ks2_config1_window_11x5(ctx, oracle_lib):
    # BASELINE.json configs[0]: 20 dates, KS2, 11x5 window (Nx=5, Ny=2); 192x256 crop of the 512x512 case
    slc = synth.make_stack(20, 192, 256, seed=1)
    c = _check(ctx, oracle_lib, slc, 5, 2, "KS2")
    assert 1 < c.mean() < 55


@pytest.mark.parametrize("bands", [5, 10, 30, 59])
def test_ks2_band_counts(ctx, oracle_lib, bands):
    slc = synth.make_stack(bands, 40, 70, seed=bands, region=16)
    _check(ctx, oracle_lib, slc, 5, 5, "KS2")


@pytest.mark.parametrize("pvalue", [0.0, 0.01, 0.2, 0.95, 1.5])
def test_ks2_thresholds(ctx, oracle_lib, pvalue):
    slc = synth.make_stack(15, 33, 47, seed=3, region=8)
    _check(ctx, oracle_lib, slc, 3, 2, "KS2", pvalue=pvalue)


def test_ks2_heavy_ties_and_inf(ctx, oracle_lib):
    rng = np.random.default_rng(7)
    slc = synth.make_stack(12, 30, 41, seed=4, region=8, zero_fraction=0.0)
    # quantise amplitudes to force ties inside and across pixels
    amp = np.round(np.abs(slc) * 2) / 2 + 0.5
    slc = (amp * np.exp(1j * np.angle(slc))).astype(np.complex64)
    slc[3, 5, 5] = np.inf
    slc[7, 5, 6] = complex(np.inf, 1.0)
    slc[2, 9, 9] = complex(np.nan, 0.0)      # NaN amplitude -> pixel invalid
    slc[:, 20, 20] = 1.0 + 0j                # all dates identical
    _check(ctx, oracle_lib, slc, 4, 3, "KS2")


def test_ks2_mask_and_calibration(ctx, oracle_lib):
    rng = np.random.default_rng(11)
    slc = synth.make_stack(16, 37, 53, seed=5, region=16)
    mask = (rng.random((37, 53)) > 0.2).astype(np.uint8)
    alpha = np.concatenate([[1.0], rng.uniform(0.5, 2.0, 15)])
    _check(ctx, oracle_lib, slc, 5, 2, "KS2", mask=mask, alpha=alpha)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (7, 1), (3, 4), (9, 40)])
def test_ks2_tiny_images(ctx, oracle_lib, shape):
    slc = synth.make_stack(8, shape[0], shape[1], seed=6, region=4, zero_fraction=0.0)
    _check(ctx, oracle_lib, slc, 5, 2, "KS2")


def test_ad2_small_window(ctx, oracle_lib):
    slc = synth.make_stack(20, 48, 64, seed=8, region=16)
    _check(ctx, oracle_lib, slc, 5, 2, "AD2")


def test_ad2_wide_window_21x21(ctx, oracle_lib):
    # BASELINE.json configs[4]: AD2, 21x21 window, 30 dates (crop)
    slc = synth.make_stack(30, 40, 56, seed=9, region=16)
    _check(ctx, oracle_lib, slc, 10, 10, "AD2")


def test_ad2_ties(ctx, oracle_lib):
    slc = synth.make_stack(10, 24, 30, seed=10, region=8, zero_fraction=0.0)
    amp = np.round(np.abs(slc) * 2) / 2 + 0.5
    slc = (amp * np.exp(1j * np.angle(slc))).astype(np.complex64)
    _check(ctx, oracle_lib, slc, 3, 3, "AD2")


@pytest.mark.parametrize("pvalue", [0.01, 0.05, 0.3])
def test_ad2_thresholds(ctx, oracle_lib, pvalue):
    slc = synth.make_stack(12, 30, 30, seed=12, region=8)
    _check(ctx, oracle_lib, slc, 3, 3, "AD2", pvalue=pvalue)


def test_large_bands_100(ctx, oracle_lib):
    slc = synth.make_stack(100, 20, 48, seed=13, region=16)
    _check(ctx, oracle_lib, slc, 5, 2, "KS2")
    _check(ctx, oracle_lib, slc, 2, 2, "AD2")


def test_unknown_method_status(ctx):
    from fringe_b200._lib import FringeError
    slc = synth.make_stack(5, 4, 4, seed=1)
    with pytest.raises(FringeError) as e:
        ctx.nmap_block(slc, 1, 1, method="XYZ")
    assert e.value.status == 1      # same code nmap_process returns for an unknown method


# ---- windows and band counts beyond the shared-memory tile / the 32-word mask (the reference's own
# sequential defaults, 59 x 19 = 1121 pixels = 36 words, used to be rejected) -------------------------------
@pytest.mark.parametrize("Nx,Ny,method", [(29, 9, "KS2"), (11, 5, "KS2"), (11, 5, "AD2"), (29, 9, "AD2")])
def test_wide_windows_sequential_defaults(ctx, oracle_lib, Nx, Ny, method):
    slc = synth.make_stack(12, 30, 90, seed=5, region=16)
    c = _check(ctx, oracle_lib, slc, Nx, Ny, method)
    if (Nx, Ny) == (29, 9):
        from fringe_b200 import engine
        assert engine.nulong(Nx, Ny) == 36


@pytest.mark.parametrize("method", ["KS2", "AD2"])
def test_200_bands_global_memory_kernel(ctx, oracle_lib, method):
    """BASELINE configs[3]: the SHP input of the 200-date sequential case.  With the 59 x 19 window no tile of 200
    ranks fits shared memory: the pair tests run from global memory, results unchanged."""
    slc = synth.make_stack(200, 24, 70, seed=9, region=16)
    _check(ctx, oracle_lib, slc, 29, 9, method)
    _check(ctx, oracle_lib, slc, 5, 2, method)


def test_nmap_process_block_shim(oracle_lib):
    """The reference's own C-style boundary (src/nmap/nmap_cuda.h:13-17): amplitudes in, KS2, host pointers."""
    import ctypes as C
    from fringe_b200 import _lib
    raw = C.CDLL(_lib.LIB_PATH)
    fn = getattr(raw, "_Z16nmapProcessBlockPfPhiiiPiPjidii")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
    slc = synth.make_stack(20, 48, 80, seed=11, region=16)
    bands, lines, cols = slc.shape
    c_ref, w_ref, amp_sorted = oracle_lib.nmap_block(slc, 5, 2, want_amp=True)
    # what nmap.cpp:370-381 hands over: unsorted amplitudes [pixel][band] and the validity mask
    amp = np.ascontiguousarray(np.abs(slc).transpose(1, 2, 0).astype(np.float32))
    valid = np.ascontiguousarray((np.abs(slc) != 0).all(axis=0).astype(np.uint8))
    amp[valid == 0] = 0
    cnt = np.zeros((lines, cols), np.int32)
    wts = np.zeros((lines, cols, 2), np.uint32)
    getattr(raw, "_Z7lockGPUv")()
    fn(amp.ctypes.data, valid.ctypes.data, cols, lines, bands, cnt.ctypes.data, wts.ctypes.data, 2, 0.05, 5, 2)
    getattr(raw, "_Z9unlockGPUv")()
    assert np.array_equal(cnt, c_ref) and np.array_equal(wts, w_ref)


@pytest.mark.parametrize("Nx,Ny", [(6, 1), (7, 3), (9, 2), (2, 0)])
@pytest.mark.parametrize("method", ["KS2", "AD2"])
def test_tma_tile_alignment_shifts(ctx, oracle_lib, Nx, Ny, method):
    """The TMA box of a tile has to start on a 16-byte boundary, so the tile grid is shifted by (-Nx) mod 4 columns
    (k_nmap): every residue of Nx, an image width that is not a multiple of 4 (padded row pitch of the rank planes) and
    an image narrower than one tile."""
    for cols in (61, 23):
        slc = synth.make_stack(9, 21, cols, seed=10 * Nx + Ny, region=8)
        _check(ctx, oracle_lib, slc, Nx, Ny, method)


def test_more_than_64_bands_takes_the_insertion_sort(ctx, oracle_lib):
    # the register sorting network covers bands <= 64; beyond that the shared-memory insertion sort
    slc = synth.make_stack(70, 18, 37, seed=70, region=8)
    _check(ctx, oracle_lib, slc, 3, 2, "KS2")
