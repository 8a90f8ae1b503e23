"""CPU: pins the oracle (test infrastructure) before anything is compared against it.

* the restated port (oracle/restated.hpp) agrees bit-for-bit with the reference's own headers
  compiled in place (oracle/_ref, only where that build is present);
* both reproduce the committed golden fixtures (tests/golden, generated from the reference-header
  build by tests/golden/make_golden.py);
* the few known answers the reference's tests print: the 3x3 matrix of tests/eigen/test_eig.cpp,
  the bit layout of tests/bitmask/test_bits.cpp, "K-statistic equals scipy's"
  (tests/KS2/README.md:39).
"""
import os

import numpy as np
import pytest

import oracle
from fringe_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KINDS = [k for k in ("port", "reference") if k == "port" or oracle.available("reference")]


@pytest.fixture(scope="module", params=KINDS)
def lib(request):
    return oracle.load(request.param)


def test_reference_build_present_in_authoring_container():
    if os.path.isdir("/root/reference"):
        assert oracle.available("reference"), "run `make -C oracle ref`"


def test_known_answers(lib):
    a = np.array([1, 2, 3, 4, 5], np.float32)
    b = np.array([1.5, 2.5, 3.5, 9, 10], np.float32)
    assert lib.ks2_prob(a, b) == 0.8186211748061034          # SURVEY.md 8(c) probe values
    assert lib.ad2_prob(a, b) == 0.7162112461293072
    assert lib.ad2_sigma(30) == 0.738262623146096
    A = np.array([[3, -2, 4], [-2, 8, 2], [4, 2, 3]], float)  # tests/eigen/test_eig.cpp:55-64
    info, val, vec = lib.eig_extreme(A, largest=True)
    assert info == 0 and abs(val - 8.81507314463) < 1e-9
    w, v = np.linalg.eigh(A)
    assert abs(abs(np.vdot(v[:, -1], vec)) - 1) < 1e-9
    info, val, _ = lib.eig_extreme(A, largest=False)
    assert info == 0 and abs(val - w[0]) < 1e-6                # abstol 1e-6 (EigenLapack.hpp:145)
    info, inv = lib.pd_inverse(A + 2 * np.eye(3))              # tests/eigen/test_eig.cpp:97-99
    assert info == 0 and np.allclose(inv, np.linalg.inv(A + 2 * np.eye(3)))
    info, _ = lib.pd_inverse(A - 10 * np.eye(3))
    assert info != 0                                           # not positive definite


@pytest.mark.parametrize("Ny,Nx", [(2, 2), (7, 7), (2, 5), (10, 10)])
def test_bit_layout(lib, Ny, Nx):
    # tests/bitmask/test_bits.cpp:18-24: flat=(ii+Ny)*(2Nx+1)+jj+Nx ; word=flat/32 ; bit=flat%32
    nu = oracle.nulong(Nx, Ny)
    for dy, dx in [(-Ny, -Nx), (0, 0), (Ny, Nx), (-1, 1), (Ny, -Nx)]:
        words = np.zeros(nu, np.uint32)
        lib.mask_setbit(words, Ny, Nx, dy, dx, True)
        flat = (dy + Ny) * (2 * Nx + 1) + dx + Nx
        expect = np.zeros(nu, np.uint32)
        expect[flat // 32] = np.uint32(1) << np.uint32(flat % 32)
        assert np.array_equal(words, expect)
        assert lib.mask_getbit(words, Ny, Nx, dy, dx)
        lib.mask_setbit(words, Ny, Nx, dy, dx, False)
        assert not words.any()


def test_ks_statistic_matches_scipy(lib):
    stats = pytest.importorskip("scipy.stats")
    rng = np.random.default_rng(5)
    for n in (10, 25, 30, 50):
        a = np.sort(rng.rayleigh(1.0, n)).astype(np.float32)
        b = np.sort(rng.exponential(1.3, n)).astype(np.float32)
        d = round(float(stats.ks_2samp(a, b).statistic) * n) / n      # scipy returns it in float32 here
        assert abs(lib.ks2_prob(a, b) - lib.kolmogorov_prob(d * np.sqrt(n / 2.0))) < 1e-12


def test_golden_pairs(lib):
    g = np.load(os.path.join(GOLD, "pairs.npz"))
    for a, b, ks, ad in zip(g["a"], g["b"], g["ks_p"], g["ad_p"]):
        n = int(np.isfinite(a).sum())
        assert lib.ks2_prob(a[:n], b[:n]) == ks
        assert lib.ad2_prob(a[:n], b[:n]) == ad
    assert [lib.ad2_sigma(n) for n in (5, 10, 20, 30, 100)] == list(g["ad_sigma"])


def test_golden_blocks(lib):
    g = np.load(os.path.join(GOLD, "block_12x24x40.npz"))
    slc = g["slc"]
    c, w = lib.nmap_block(slc, 5, 2, method=oracle.KS2)
    assert np.array_equal(c, g["count_ks"]) and np.array_equal(w, g["wts_ks"])
    c, w = lib.nmap_block(slc, 3, 3, method=oracle.AD2)
    assert np.array_equal(c, g["count_ad"]) and np.array_equal(w, g["wts_ad"])
    cases = (("evd", dict(method=oracle.EVD)), ("mle", dict(method=oracle.MLE)),
             ("stbas", dict(method=oracle.STBAS, bandwidth=4)),
             ("pl", dict(method=oracle.MLE, variant=oracle.VARIANT_PHASE_LINK, min_neighbors=5)),
             ("seq", dict(method=oracle.MLE, mini_stack_count=3, first_line=2, n_lines=20)))
    for name, kw in cases:
        o, t, cp = lib.evd_block(slc, g["wts_ks"], 5, 2, **kw)
        assert np.array_equal(t < 0, g[name + "_tcorr"] < 0), name
        assert np.allclose(t, g[name + "_tcorr"], atol=1e-6), name
        ok = g[name + "_tcorr"] > 0.3
        assert np.abs(np.angle(o[:, ok] * np.conj(g[name + "_out"][:, ok]))).max() < 1e-5, name
        assert np.allclose(cp[ok], g[name + "_comp"][ok], atol=1e-4), name


def test_golden_post_rows(lib):
    """SURVEY 8f rows against the committed fixtures (tests/golden/make_golden.py: make_post)."""
    g = np.load(os.path.join(GOLD, "block_12x24x40.npz"))
    p = np.load(os.path.join(GOLD, "post_12x24x40.npz"))
    slc, wts = g["slc"], g["wts_ks"]
    same = lambda a, b: np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
    da, mean = lib.ampdispersion_block(slc, p["alpha"])
    assert same(da, p["ampdisp_da"]) and same(mean, p["ampdisp_mean"])
    assert same(lib.despeck_block(slc[2], wts, 5, 2), p["despeck_amp"])
    assert same(lib.despeck_block(slc[2], wts, 5, 2, z2=slc[9]), p["despeck_ifg"])
    assert same(lib.despeck_block(slc[2], wts, 5, 2, z2=slc[9], coherence=True), p["despeck_coh"])
    assert same(lib.cmul(slc[4], slc[7]), p["cmul"])


@pytest.mark.skipif(not oracle.available("reference"), reason="reference-header build not present")
def test_port_equals_reference_bit_for_bit():
    port, ref = oracle.load("port"), oracle.load("reference")
    rng = np.random.default_rng(9)
    for n in (7, 20, 30, 61):
        for t in range(40):
            a = np.sort(rng.rayleigh(1.0, n)).astype(np.float32)
            b = np.sort(rng.rayleigh(rng.choice([1.0, 2.0]), n)).astype(np.float32)
            if t % 2:
                a = np.sort(np.round(a * 3) / 3).astype(np.float32) + 1
                b = np.sort(np.round(b * 3) / 3).astype(np.float32) + 1
            assert port.ks2_prob(a, b) == ref.ks2_prob(a, b)
            assert port.ad2_prob(a, b) == ref.ad2_prob(a, b)
        assert port.ad2_sigma(n) == ref.ad2_sigma(n)
    for z in np.linspace(0, 7.5, 400):
        assert port.kolmogorov_prob(z) == ref.kolmogorov_prob(z)
    for a2 in np.linspace(-2, 13, 300):
        assert port.ad2_pvalue_of_stat(a2, 30) == ref.ad2_pvalue_of_stat(a2, 30)
    slc = synth.make_stack(14, 20, 36, seed=3, region=8)
    for method in (oracle.KS2, oracle.AD2):
        cp, wp = port.nmap_block(slc, 4, 2, method=method)
        cr, wr = ref.nmap_block(slc, 4, 2, method=method)
        assert np.array_equal(cp, cr) and np.array_equal(wp, wr)
    for kw in (dict(method=oracle.EVD), dict(method=oracle.MLE), dict(method=oracle.STBAS, bandwidth=3),
               dict(method=oracle.MLE, variant=oracle.VARIANT_PHASE_LINK, min_neighbors=4)):
        for x, y in zip(port.evd_block(slc, wp, 4, 2, **kw), ref.evd_block(slc, wp, 4, 2, **kw)):
            assert np.array_equal(x, y)
    # the SURVEY 8f rows: despeck reads the mask through the reference's own Ulongmask in the reference build
    for kw in (dict(), dict(z2=slc[5]), dict(z2=slc[5], coherence=True)):
        a, b = port.despeck_block(slc[2], wp, 4, 2, **kw), ref.despeck_block(slc[2], wp, 4, 2, **kw)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for x, y in zip(port.ampdispersion_block(slc), ref.ampdispersion_block(slc)):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_block_results_independent_of_block_schedule(lib):
    """The reference streams overlapping row blocks (nmap.cpp:487-573, evd.cpp:414-443,861-868) and
    states results do not depend on the block height; check that on the oracle."""
    from fringe_b200.partition import block_schedule
    slc = synth.make_stack(8, 40, 24, seed=12, region=8)
    c_all, w_all = lib.nmap_block(slc, 3, 2)
    o_all, t_all, cp_all = lib.evd_block(slc, w_all, 3, 2, method=oracle.EVD)
    c_blk = np.zeros_like(c_all); w_blk = np.zeros_like(w_all)
    o_blk = np.zeros_like(o_all); t_blk = np.zeros_like(t_all)
    for yoff, ny, first, nwrite in block_schedule(40, 16, 2):
        c, w = lib.nmap_block(slc[:, yoff:yoff + ny], 3, 2)
        c_blk[yoff + first:yoff + first + nwrite] = c[first:first + nwrite]
        w_blk[yoff + first:yoff + first + nwrite] = w[first:first + nwrite]
    assert np.array_equal(c_blk, c_all) and np.array_equal(w_blk, w_all)
    for yoff, ny, first, nwrite in block_schedule(40, 16, 2):
        o, t, _ = lib.evd_block(slc[:, yoff:yoff + ny], w_all[yoff:yoff + ny], 3, 2, method=oracle.EVD,
                                first_line=first, n_lines=nwrite)
        o_blk[:, yoff + first:yoff + first + nwrite] = o[:, first:first + nwrite]
        t_blk[yoff + first:yoff + first + nwrite] = t[first:first + nwrite]
    assert np.array_equal(t_blk, t_all) and np.array_equal(o_blk, o_all)
