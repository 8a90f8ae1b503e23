"""The oracle pinned to the reference's OWN block drivers.

`oracle/ref_drivers/` compiles src/nmap/nmap.cpp, src/evd/evd.cpp, src/phase_link/phase_link.cpp, src/despeck/despeck.cpp,
src/ampdispersion/ampdispersion.cpp and src/calamp/calamp.cpp unmodified from /root/reference against stand-ins for GDAL
and Armadillo (oracle/shims/: file I/O and column-major storage, no arithmetic of the path) and runs them file in, file
out, single-threaded.  Every restated loop of oracle/loops.hpp -- what the GPU parity tests compare against -- must give
what those drivers write: bit for bit for the integer / float32 products, to LAPACK-call reproducibility for the
eigenvectors (same OpenBLAS, same call sequence)."""
import os

import numpy as np
import pytest

import oracle as oracle_pkg
from conftest import wrapped_diff
from fringe_b200 import stackio, synth

pytestmark = pytest.mark.skipif(not all(oracle_pkg.ref_driver_available(n) for n in oracle_pkg.REF_DRIVERS),
                                reason="reference drivers not built (needs /root/reference: make -C oracle refdrivers)")


def b(s):
    return None if s is None else str(s).encode()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.fixture(scope="module")
def stack(tmp_path_factory, oracle_lib):
    root = str(tmp_path_factory.mktemp("refstack"))
    slc = synth.make_stack(12, 150, 96, seed=31, region=32)
    vrt = stackio.make_stack_on_disk(root, slc)
    w = os.path.join(root, "nmap_ks2")
    c = os.path.join(root, "count_ks2")
    # memsize 1 MB, 64-line boxes -> 128-line blocks: two overlapping blocks for 150 lines (nmap.cpp:166-183)
    rc = oracle_pkg.ref_driver("nmap").ref_nmap(b(vrt), b(w), b(c), None, 5, 2, b"KS2", 0.05, 1, 64)
    assert rc == 0
    return root, slc, vrt, w, c


def test_nmap_driver_ks2_and_ad2(stack, oracle_lib):
    root, slc, vrt, w, c = stack
    count, wts = stackio.read_envi(c), stackio.read_envi(w)
    assert count.dtype == np.int16 and wts.dtype == np.uint32 and wts.shape == (150, 96, 2)
    hdr = stackio.read_envi_header(w)
    assert hdr["halfwindowx"] == "5" and hdr["halfwindowy"] == "2"
    c_ref, w_ref = oracle_lib.nmap_block(slc, 5, 2, method=0, thresh=0.05)
    assert np.array_equal(count.astype(np.int32), c_ref) and np.array_equal(wts, w_ref)
    w2, c2 = os.path.join(root, "nmap_ad2"), os.path.join(root, "count_ad2")
    assert oracle_pkg.ref_driver("nmap").ref_nmap(b(vrt), b(w2), b(c2), None, 4, 3, b"AD2", 0.05, 1, 64) == 0
    c_ref, w_ref = oracle_lib.nmap_block(slc, 4, 3, method=1, thresh=0.05)
    assert np.array_equal(stackio.read_envi(c2).astype(np.int32), c_ref) and np.array_equal(stackio.read_envi(w2), w_ref)


def test_nmap_driver_mask_and_calibration(stack, oracle_lib, tmp_path):
    root, slc, vrt, w, c = stack
    rng = np.random.default_rng(3)
    mask = (rng.random((150, 96)) > 0.2).astype(np.uint8)
    mpath = str(tmp_path / "mask.bin")
    stackio.write_envi(mpath, mask)
    alpha = np.concatenate([[2.0], rng.uniform(0.5, 3.0, 11)])
    croot = str(tmp_path / "cal")
    cvrt = stackio.make_stack_on_disk(croot, slc, extra_md={d: {"amplitudeConstant": repr(float(a))}
                                                            for d, a in zip(stackio.default_dates(12), alpha)})
    w2, c2 = str(tmp_path / "nmap"), str(tmp_path / "count")
    assert oracle_pkg.ref_driver("nmap").ref_nmap(b(cvrt), b(w2), b(c2), b(mpath), 5, 2, b"KS2", 0.05, 1, 64) == 0
    c_ref, w_ref = oracle_lib.nmap_block(slc, 5, 2, method=0, thresh=0.05, mask=mask, alpha=alpha / alpha[0])
    assert np.array_equal(stackio.read_envi(c2).astype(np.int32), c_ref) and np.array_equal(stackio.read_envi(w2), w_ref)


# (STBAS is not run through the reference driver: its temporal-coherence loop indexes Covar and evddata past the matrix,
# evd.cpp:770-786 with ulim = ti+BW+1 -- unchecked reads hundreds of kB beyond the arrays, i.e. undefined behaviour; the
# oracle clamps that sum, oracle/loops.hpp, and the STBAS eigen solve is the EVD one on a band-limited matrix.)
@pytest.mark.parametrize("method,code,kw", [("EVD", 0, {}), ("MLE", 1, {}), ("MLE", 1, {"mini_stack_count": 3}),
                                             ("EVD", 0, {"mini_stack_count": 12})])
def test_evd_driver(stack, oracle_lib, tmp_path, method, code, kw):
    root, slc, vrt, w, c = stack
    out = str(tmp_path / "EVD")
    rc = oracle_pkg.ref_driver("evd").ref_evd(b(vrt), b(w), b(out), b(out), b"compslc.bin", 5, 2, b(method), kw.get("bandwidth", -1),
                                              kw.get("mini_stack_count", 1), 2, 1, 64)
    assert rc == 0
    wts = stackio.read_envi(w)
    o_ref, t_ref, c_ref = oracle_lib.evd_block(slc, wts, 5, 2, method=code, **kw)
    tcorr = stackio.read_envi(os.path.join(out, "tcorr.bin"))
    comp = stackio.read_envi(os.path.join(out, "compslc.bin"))
    dates = stackio.default_dates(12)
    phasors = np.stack([stackio.read_envi(os.path.join(out, d + ".slc")) for d in dates])
    # sentinels and skipped pixels: identical codes
    assert np.array_equal(np.where(t_ref <= 0, t_ref, 0), np.where(tcorr <= 0, tcorr, 0))
    ok = t_ref > 0
    assert ok.sum() > 5000
    assert np.abs(tcorr - t_ref)[ok].max() <= 1e-6
    assert wrapped_diff(phasors[:, ok], o_ref[:, ok]).max() <= 1e-5
    assert np.abs(comp - c_ref)[ok].max() <= 1e-5 * np.abs(c_ref).max()
    assert np.all(phasors[:, ~ok] == 0)


def test_phase_link_driver(stack, oracle_lib, tmp_path):
    root, slc, vrt, w, c = stack
    out = str(tmp_path / "PL")
    rc = oracle_pkg.ref_driver("phase_link").ref_phase_link(b(vrt), b(w), b(out), b(out), b"compslc.bin", 5, 2, b"MLE", -1, 1, 5, 1, 64)
    assert rc == 0
    wts = stackio.read_envi(w)
    o_ref, t_ref, c_ref = oracle_lib.evd_block(slc, wts, 5, 2, method=1, variant=1, min_neighbors=5)
    tcorr = stackio.read_envi(os.path.join(out, "tcorr.bin"))
    phasors = np.stack([stackio.read_envi(os.path.join(out, d + ".slc")) for d in stackio.default_dates(12)])
    assert np.array_equal(np.where(t_ref <= 0, t_ref, 0), np.where(tcorr <= 0, tcorr, 0))
    ok = t_ref > 0
    assert np.abs(tcorr - t_ref)[ok].max() <= 1e-6
    assert wrapped_diff(phasors[:, ok], o_ref[:, ok]).max() <= 1e-5


@pytest.mark.parametrize("band2,coh", [(0, False), (4, False), (4, True), (0, True)])
def test_despeck_driver(stack, oracle_lib, tmp_path, band2, coh):
    root, slc, vrt, w, c = stack
    out = str(tmp_path / "despeck.bin")
    assert oracle_pkg.ref_driver("despeck").ref_despeck(b(vrt), b(w), b(out), 5, 2, 1, band2, int(coh), 1, 64) == 0
    got = stackio.read_envi(out)
    wts = stackio.read_envi(w)
    want = oracle_lib.despeck_block(slc[0], wts, 5, 2, z2=slc[band2 - 1] if band2 else None, coherence=coh)
    if got.dtype != np.complex64:                     # single band: Float32 raster (despeck.cpp:191-195)
        want = want.real.astype(np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_ampdispersion_driver(stack, oracle_lib, tmp_path):
    root, slc, vrt, w, c = stack
    da, mean = str(tmp_path / "da.bin"), str(tmp_path / "mean.bin")
    assert oracle_pkg.ref_driver("ampdispersion").ref_ampdispersion(b(vrt), b(da), b(mean), 1, 1, 64) == 0
    r_da, r_mean = oracle_lib.ampdispersion_block(slc)
    assert np.array_equal(bits(stackio.read_envi(da)), bits(r_da)) and np.array_equal(bits(stackio.read_envi(mean)), bits(r_mean))
    assert stackio.read_envi_header(da)["n"] == "12"


def test_calamp_driver(stack, oracle_lib, tmp_path):
    root, slc, vrt, w, c = stack
    import ctypes as C
    consts = (C.c_double * 12)()
    rc = oracle_pkg.ref_driver("calamp").ref_calamp(b(vrt), None, b(str(tmp_path / "cal.vrt")), 1.0, 0, 1, 64, 12, consts)
    assert rc == 0
    sums, counts = oracle_lib.calamp_block(slc)
    want = sums / counts
    got = np.array(list(consts))
    # the constant goes into the VRT as text through an ostringstream: six significant digits (calamp.cpp:260-264)
    assert np.array_equal(got, np.array([float("%g" % v) for v in want]))
