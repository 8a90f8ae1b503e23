"""GPU: calamp and PS / DS integration against the oracle, through the C ABI, the
calamplib binding / command line and the integrate_ps command line."""
import os
import re

import numpy as np
import pytest

from fringe_b200 import stackio, synth

pytestmark = pytest.mark.gpu


def test_calamp_block_against_oracle(ctx, oracle_lib):
    slc = synth.make_stack(7, 96, 130, seed=9, region=32)
    slc[1, 4, 4] = np.nan
    mask = (np.random.default_rng(2).random((96, 130)) > 0.3).astype(np.uint8)
    for m in (None, mask):
        s_ref, c_ref = oracle_lib.calamp_block(slc, m)
        s_gpu, c_gpu = ctx.calamp_block(slc, m)
        assert np.array_equal(c_ref, c_gpu)
        assert np.allclose(s_ref, s_gpu, rtol=1e-12, atol=0)


def test_calamp_cli_round_trip_into_nmap(tmp_path, oracle_lib, ctx):
    """calamp.py writes amplitudeConstant into a copy of the stack VRT; nmap reads it back (nmap.cpp:204-233) and
    the result equals the oracle run with the same constants."""
    from fringe_b200.cli import calamp as calamp_cli
    from fringe_b200.cli import nmap as nmap_cli
    root = str(tmp_path)
    slc = synth.make_stack(8, 70, 64, seed=13, region=16)
    slc *= np.linspace(0.5, 2.0, 8, dtype=np.float32)[:, None, None]          # different gains per date
    vrt = stackio.make_stack_on_disk(root, slc)
    out_vrt = os.path.join(root, "cal", "stack_cal.vrt")
    os.makedirs(os.path.dirname(out_vrt))
    calamp_cli.main(["-i", vrt, "-o", out_vrt, "-l", "32", "-r", "1"])
    consts = [float(v) for v in re.findall(r'<MDI key="amplitudeConstant">(.*?)</MDI>', open(out_vrt).read())]
    s_ref, c_ref = oracle_lib.calamp_block(slc)
    want = [float("%g" % v) for v in s_ref / c_ref]                      # calamp.cpp:258-260: default stream formatting
    assert consts == want
    wts_path, cnt_path = os.path.join(root, "cal", "nmap"), os.path.join(root, "cal", "count")
    nmap_cli.main(["-i", out_vrt, "-o", wts_path, "-c", cnt_path, "-x", "4", "-y", "2"])
    alpha = np.array(consts) / consts[0]
    alpha[0] = 1.0
    c_o, w_o = oracle_lib.nmap_block(slc, 4, 2, alpha=alpha)
    assert np.array_equal(stackio.read_envi(wts_path), w_o) and np.array_equal(stackio.read_envi(cnt_path).astype(np.int32), c_o)


def test_integrate_ps_against_oracle_and_cli(tmp_path, ctx, oracle_lib):
    from fringe_b200.cli import integrate_ps as ips_cli
    rng = np.random.default_rng(21)
    root = str(tmp_path)
    slc = synth.make_stack(5, 60, 48, seed=17, region=16)
    vrt = stackio.make_stack_on_disk(root, slc)
    dates = stackio.default_dates(5)
    ds = np.exp(1j * rng.uniform(-np.pi, np.pi, (5, 60, 48))).astype(np.complex64)
    ds_dir = os.path.join(root, "adjusted")
    os.makedirs(ds_dir)
    for d, arr in zip(dates, ds):
        stackio.write_envi(os.path.join(ds_dir, d + ".slc"), arr)
    ps = (rng.random((60, 48)) > 0.8).astype(np.uint8)
    tcorr = rng.random((60, 48)).astype(np.float32)
    stackio.write_envi(os.path.join(root, "ps.bin"), ps)
    stackio.write_envi(os.path.join(root, "tcorr.bin"), tcorr)
    got = ctx.integrate_ps(ds[0], ds[2], slc[0], slc[2], ps)
    want = oracle_lib.integrate_ps(ds[0], ds[2], slc[0], slc[2], ps)
    assert np.abs(got - want).max() <= 1e-6
    assert np.array_equal(got[ps == 0].view(np.uint32), want[ps == 0].view(np.uint32))      # DS pixels: one complex64 product
    out = os.path.join(root, "psds")
    ips_cli.main(["-s", vrt, "-d", ds_dir, "-t", os.path.join(root, "tcorr.bin"), "-p", os.path.join(root, "ps.bin"), "-o", out])
    for j in range(1, 5):
        ifg = stackio.read_envi(os.path.join(out, f"{dates[0]}_{dates[j]}.int"))
        assert np.abs(ifg - oracle_lib.integrate_ps(ds[0], ds[j], slc[0], slc[j], ps)).max() <= 1e-6
    cor = stackio.read_envi(os.path.join(out, "tcorr_ds_ps.bin"))
    assert np.array_equal(cor, np.where(ps == 1, np.float32(0.95), tcorr))


def test_integrate_ps_against_the_reference_scripts_output(ctx):
    """The kernel against the vectors the reference's own integratePS.py produced (tests/golden/make_golden_integrate_ps.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "integrate_ps_24x40.npz"))
    for j in (1, 2, 3):
        got = ctx.integrate_ps(g["ds"][0], g["ds"][j], g["slc"][0], g["slc"][j], g["ps"])
        want = g[f"ifg_0_{j}"]
        assert np.abs(got - want).max() <= 1e-6
        assert np.abs(got - want)[g["ps"] != 1].max() <= 1.5e-7          # an ulp: see tests/test_calamp_ps_cpu.py
    assert np.array_equal(ctx.ps_coherence(g["tcorr"], g["ps"]), g["coherence"])
