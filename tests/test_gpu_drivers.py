"""GPU: the file-level drop-in boundary -- Python bindings (nmaplib / evdlib / phase_linklib), the
C++ block drivers with the reference's block/halo schedule, VRT stack in and ENVI rasters out --
against the CPU oracle run on the same stack in memory."""
import os

import numpy as np
import pytest

from conftest import wrapped_diff
from fringe_b200 import stackio, synth
from fringe_b200.cli import evd as evd_cli
from fringe_b200.cli import nmap as nmap_cli
from fringe_b200.cli import phase_link as pl_cli
from fringe_b200.cli import sequential as seq_cli

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("stack"))
    slc = synth.make_stack(12, 150, 96, seed=31, region=32)
    vrt = stackio.make_stack_on_disk(root, slc)
    return root, slc, vrt


def test_nmap_evd_cli_multi_block(stack, oracle_lib):
    root, slc, vrt = stack
    wts_path = os.path.join(root, "KS2", "nmap")
    cnt_path = os.path.join(root, "KS2", "count")
    # 1 MB -> 64-line blocks -> 3 overlapping blocks for 150 lines (nmap.cpp:166-183 schedule)
    nmap_cli.main(["-i", vrt, "-o", wts_path, "-c", cnt_path, "-x", "5", "-y", "2", "-r", "1"])
    count = stackio.read_envi(cnt_path)
    wts = stackio.read_envi(wts_path)
    assert count.dtype == np.int16 and wts.dtype == np.uint32 and wts.shape == (150, 96, 2)
    hdr = stackio.read_envi_header(wts_path)
    assert hdr["halfwindowx"] == "5" and hdr["halfwindowy"] == "2" and hdr["interleave"] == "bip"
    c_ref, w_ref = oracle_lib.nmap_block(slc, 5, 2)
    assert np.array_equal(count.astype(np.int32), c_ref) and np.array_equal(wts, w_ref)

    out_dir = os.path.join(root, "EVD")
    evd_cli.main(["-i", vrt, "-w", wts_path, "-o", out_dir, "-x", "5", "-y", "2", "-m", "EVD", "-r", "1"])
    o_ref, t_ref, c_ref2 = oracle_lib.evd_block(slc, w_ref, 5, 2, method=0)
    tcorr = stackio.read_envi(os.path.join(out_dir, "tcorr.bin"))
    comp = stackio.read_envi(os.path.join(out_dir, "compslc.bin"))
    dates = stackio.default_dates(12)
    out = np.stack([stackio.read_envi(os.path.join(out_dir, d + ".slc")) for d in dates])
    assert all(os.path.exists(os.path.join(out_dir, d + ".slc.vrt")) for d in dates)
    ok = t_ref > 0
    assert np.abs(tcorr - t_ref)[ok].max() <= 1e-4
    good = t_ref > 0.3
    assert wrapped_diff(out[:, good], o_ref[:, good]).max() <= 1e-3
    assert np.abs(comp - c_ref2)[good].max() <= 2e-3 * np.abs(c_ref2).max()
    # rerun refuses to overwrite (evd.cpp:229-251 -> rc 113)
    with pytest.raises(RuntimeError, match="113"):
        evd_cli.main(["-i", vrt, "-w", wts_path, "-o", out_dir, "-x", "5", "-y", "2", "-m", "EVD"])


def test_phase_link_cli_and_error_codes(stack, oracle_lib):
    root, slc, vrt = stack
    wts_path = os.path.join(root, "KS2", "nmap")
    out_dir = os.path.join(root, "PL")
    pl_cli.main(["-i", vrt, "-w", wts_path, "-o", out_dir, "-x", "5", "-y", "2", "-n", "5", "-r", "2"])
    w_ref = stackio.read_envi(wts_path)
    o_ref, t_ref, _ = oracle_lib.evd_block(slc, w_ref, 5, 2, method=1, variant=1, min_neighbors=5)
    tcorr = stackio.read_envi(os.path.join(out_dir, "tcorr.bin"))
    assert np.array_equal(np.where(tcorr < 0, tcorr, 0), np.where(t_ref < 0, t_ref, 0))        # sentinel codes, pixel by pixel
    assert np.abs(tcorr - t_ref)[t_ref > 0].max() <= 1e-4
    dates = stackio.default_dates(12)
    out = np.stack([stackio.read_envi(os.path.join(out_dir, d + ".slc")) for d in dates])
    good = t_ref > 0.3
    assert wrapped_diff(out[:, good], o_ref[:, good]).max() <= 1e-3
    assert np.all(out[:, ~(t_ref > 0)] == 0)
    # window mismatch between wts metadata and request -> rc 109 (evd.cpp:131-141)
    with pytest.raises(RuntimeError, match="109"):
        evd_cli.main(["-i", vrt, "-w", wts_path, "-o", os.path.join(root, "X1"), "-x", "4", "-y", "2"])
    with pytest.raises(RuntimeError, match="102"):
        nmap_cli.main(["-i", os.path.join(root, "missing.vrt"), "-o", os.path.join(root, "a"), "-c", os.path.join(root, "b")])
    with pytest.raises(RuntimeError, match="returned 1"):
        nmap_cli.main(["-i", vrt, "-o", os.path.join(root, "a"), "-c", os.path.join(root, "b"), "-s", "XYZ"])


def test_sequential_chain(stack, oracle_lib):
    """BASELINE.json configs[3] in miniature through the command line: 12 dates in ministacks of 5 (5+5+2), several
    blocks of lines with halos, compressed-SLC hand-off on the device, datum connection, adjusted product.  Every
    stage's phasors, temporal coherence and compressed SLC are compared with the oracle run on the same inputs."""
    root, slc, vrt = stack
    wts_path = os.path.join(root, "KS2", "nmap")
    out = os.path.join(root, "seq")
    # -r 1: a memory budget that forces several blocks of 64 lines for the 150-line image
    seq_cli.main(["-i", os.path.join(root, "SLC"), "-w", wts_path, "-o", out, "-x", "5", "-y", "2", "-s", "5", "-r", "1"])
    w_ref = stackio.read_envi(wts_path)
    dates = stackio.default_dates(12)
    groups = [(0, 5), (5, 10), (10, 12)]
    comps, minis = [], []
    for k, (a, b) in enumerate(groups, start=1):
        bands = np.concatenate([np.array(comps), slc[a:b]]) if comps else slc[a:b]
        o_ref, t_ref, c_ref = oracle_lib.evd_block(bands.astype(np.complex64), w_ref, 5, 2, method=1, mini_stack_count=k)
        d = os.path.join(out, "miniStacks", dates[a] + "_" + dates[b - 1], "EVD")
        assert os.path.exists(os.path.join(out, "miniStacks", dates[a] + "_" + dates[b - 1], "stack", "stack.vrt"))
        tcorr = stackio.read_envi(os.path.join(d, "tcorr.bin"))
        mini = np.stack([stackio.read_envi(os.path.join(d, dates[i] + ".slc")) for i in range(a, b)])
        code_ref, code_gpu = np.where(t_ref < 0, t_ref, 0), np.where(tcorr < 0, tcorr, 0)
        flips = np.argwhere(code_ref != code_gpu)
        assert len(flips) <= 3, [(int(y), int(x), float(t_ref[y, x]), float(tcorr[y, x])) for y, x in flips]
        both = (t_ref > 0) & (tcorr > 0)
        assert np.abs(tcorr - t_ref)[both].max() <= 1e-4
        good = both & (t_ref > 0.3)
        assert wrapped_diff(mini[:, good], o_ref[k - 1:][:, good]).max() <= 1e-3          # the phasors themselves
        comp = stackio.read_envi(os.path.join(out, "compressedSlc", dates[b - 1], dates[b - 1] + ".slc"))
        assert np.abs(comp - c_ref)[good].max() <= 2e-3 * np.abs(c_ref[good]).max()
        comps.append(comp)                       # feed OUR compressed SLC forward: every stage sees what the device saw
        minis.append(mini)
    dc = os.path.join(out, "Datum_connection", "EVD")
    o_ref, t_ref, _ = oracle_lib.evd_block(np.array(comps).astype(np.complex64), w_ref, 5, 2, method=1, mini_stack_count=1)
    tcorr = stackio.read_envi(os.path.join(dc, "tcorr.bin"))
    datum = np.stack([stackio.read_envi(os.path.join(dc, dates[b - 1] + ".slc")) for _, b in groups])
    both = (t_ref > 0) & (tcorr > 0)
    assert (np.where(t_ref < 0, t_ref, 0) != np.where(tcorr < 0, tcorr, 0)).sum() <= 3
    assert np.abs(tcorr - t_ref)[both].max() <= 1e-4
    good = both & (t_ref > 0.3)
    assert wrapped_diff(datum[:, good], o_ref[:, good]).max() <= 1e-3
    # the adjusted series is the product of the two rasters on disk, evaluated as adjustMiniStacks.py's VRTs would
    for k, (a, b) in enumerate(groups):
        for i in range(a, b):
            adj = stackio.read_envi(os.path.join(out, "adjusted", dates[i] + ".slc"))
            want = oracle_lib.cmul(minis[k][i - a], datum[k])
            assert np.array_equal(adj.view(np.uint32), want.view(np.uint32))
    # a finished run is left alone (sequential.py:206-207 skips what exists); -f redoes it
    before = os.path.getmtime(os.path.join(dc, "tcorr.bin"))
    seq_cli.main(["-i", os.path.join(root, "SLC"), "-w", wts_path, "-o", out, "-x", "5", "-y", "2", "-s", "5", "-r", "1"])
    assert os.path.getmtime(os.path.join(dc, "tcorr.bin")) == before


def test_despeck_cli_multi_block(stack, oracle_lib):
    """despeck.py -> despecklib -> despeck_process (block schedule of despeck.cpp) -> fringe_despeck_block,
    against the oracle on the whole image: amplitude (Float32 out), interferogram and coherence (CFloat32)."""
    from fringe_b200.cli import despeck as despeck_cli
    root, slc, vrt = stack
    wts_path = os.path.join(root, "KS2d", "nmap")
    cnt_path = os.path.join(root, "KS2d", "count")
    nmap_cli.main(["-i", vrt, "-o", wts_path, "-c", cnt_path, "-x", "4", "-y", "3"])
    wts = stackio.read_envi(wts_path)
    # -r 1 -l 32: several overlapping blocks for 150 lines
    out = os.path.join(root, "despeck", "amp")
    despeck_cli.main(["-i", vrt, "-w", wts_path, "-o", out, "-x", "4", "-y", "3", "-b", "3", "-r", "1", "-l", "32"])
    amp = stackio.read_envi(out)
    assert amp.dtype == np.float32 and amp.shape == (150, 96)
    want = oracle_lib.despeck_block(slc[2], wts, 4, 3)
    assert np.array_equal(amp.view(np.uint32), want.real.astype(np.float32).view(np.uint32))
    for name, coh in (("ifg", False), ("coh", True)):
        out = os.path.join(root, "despeck", name)
        despeck_cli.main(["-i", vrt, "-w", wts_path, "-o", out, "-x", "4", "-y", "3", "-b", "2", "9", "-r", "1", "-l", "32"]
                         + (["-c"] if coh else []))
        got = stackio.read_envi(out)
        assert got.dtype == np.complex64
        want = oracle_lib.despeck_block(slc[1], wts, 4, 3, z2=slc[8], coherence=coh)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(Exception, match="coherence"):
        despeck_cli.main(["-i", vrt, "-w", wts_path, "-o", out + "x", "-b", "1", "-c"])


def test_rasters_equal_the_reference_drivers_rasters(stack, tmp_path):
    """File in, file out on both sides: the reference's own nmap.cpp / evd.cpp (compiled unmodified against the GDAL /
    Armadillo stand-ins, oracle/ref_drivers/) and this repository's drivers read the same stack VRT; the neighbour mask
    and count rasters must be byte-identical, the phase-linked rasters within the parity gates."""
    import oracle as oracle_pkg
    if not all(oracle_pkg.ref_driver_available(n) for n in ("nmap", "evd")):
        pytest.skip("reference drivers not built")
    root, slc, vrt = stack
    enc = lambda s: str(s).encode()
    r_w, r_c = str(tmp_path / "ref_nmap"), str(tmp_path / "ref_count")
    assert oracle_pkg.ref_driver("nmap").ref_nmap(enc(vrt), enc(r_w), enc(r_c), None, 5, 2, b"KS2", 0.05, 1, 64) == 0
    g_w, g_c = str(tmp_path / "gpu_nmap"), str(tmp_path / "gpu_count")
    nmap_cli.main(["-i", vrt, "-o", g_w, "-c", g_c, "-x", "5", "-y", "2", "-r", "1"])
    for a, b in ((r_w, g_w), (r_c, g_c)):
        assert open(a, "rb").read() == open(b, "rb").read()
        ha, hb = stackio.read_envi_header(a), stackio.read_envi_header(b)
        assert all(ha[k] == hb[k] for k in ("samples", "lines", "bands", "data type", "interleave", "halfwindowx", "halfwindowy"))
    dates = stackio.default_dates(12)
    for method in ("EVD", "MLE"):
        r_out, g_out = str(tmp_path / ("ref_" + method)), str(tmp_path / ("gpu_" + method))
        assert oracle_pkg.ref_driver("evd").ref_evd(enc(vrt), enc(r_w), enc(r_out), enc(r_out), b"compslc.bin", 5, 2, enc(method),
                                                    -1, 1, 2, 1, 64) == 0
        evd_cli.main(["-i", vrt, "-w", g_w, "-o", g_out, "-x", "5", "-y", "2", "-m", method, "-r", "1"])
        t_ref, t_gpu = stackio.read_envi(os.path.join(r_out, "tcorr.bin")), stackio.read_envi(os.path.join(g_out, "tcorr.bin"))
        codes_differ = np.where(t_ref <= 0, t_ref, 0) != np.where(t_gpu <= 0, t_gpu, 0)
        assert codes_differ.sum() <= (3 if method == "MLE" else 0)          # MLE: pixels on the 1e-6 eigenvalue gates
        ok = (t_ref > 0) & (t_gpu > 0)
        assert np.abs(t_ref - t_gpu)[ok].max() <= 1e-4
        good = ok & (t_ref > 0.3)
        p_ref = np.stack([stackio.read_envi(os.path.join(r_out, d + ".slc")) for d in dates])
        p_gpu = np.stack([stackio.read_envi(os.path.join(g_out, d + ".slc")) for d in dates])
        assert wrapped_diff(p_ref[:, good], p_gpu[:, good]).max() <= 1e-3
        c_ref, c_gpu = stackio.read_envi(os.path.join(r_out, "compslc.bin")), stackio.read_envi(os.path.join(g_out, "compslc.bin"))
        assert np.abs(c_ref - c_gpu)[good].max() <= 2e-3 * np.abs(c_ref).max()
