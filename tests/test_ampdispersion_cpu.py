"""ampdispersion, CPU side: the oracle's loop (src/ampdispersion/ampdispersion.cpp:207-247)
against an independent numpy evaluation, and the binding / error codes that precede device work."""
import os

import numpy as np
import pytest

from fringe_b200 import engine, stackio, synth
from fringe_b200.cli._common import use_bindings

use_bindings()
import ampdispersionlib  # noqa: E402


def _numpy_ampdisp(slc, alpha):
    re, im = slc.real.astype(np.float64), slc.imag.astype(np.float64)
    amp = np.sqrt(re * re + im * im).astype(np.float32).astype(np.float64)          # glibc hypotf
    valid = amp != 0
    bands = slc.shape[0]
    mean = np.zeros(slc.shape[1:]); meansq = np.zeros_like(mean); norms = np.zeros_like(mean)
    for b in range(bands):
        a = amp[b] * (valid[b] / alpha[b])
        mean += a
        meansq += a * a
        norms += valid[b]
    da = np.full(mean.shape, -1.0)
    m = np.zeros_like(mean)
    ok = norms > 1
    with np.errstate(all="ignore"):
        avg = mean / norms
        avg2 = meansq / norms
        sdev = np.sqrt(avg2 - avg * avg)
        good = ok & ~np.isnan(sdev) & (sdev > 0)
        da[good] = (sdev / avg)[good]
        m[ok] = avg[ok]
    return da.astype(np.float32), m.astype(np.float32)


def test_oracle_matches_numpy(oracle_lib):
    slc = synth.make_stack(9, 17, 23, seed=5, region=8, zero_fraction=0.2)
    slc[:, 3, 4] = 0                         # never valid
    slc[1:, 5, 6] = 0                        # one valid date only
    slc[:, 7, 8] = 2 + 0j                    # zero spread: sdev == 0 -> -1
    alpha = np.linspace(1.0, 1.4, 9)
    da, mean = oracle_lib.ampdispersion_block(slc, alpha)
    da2, mean2 = _numpy_ampdisp(slc, alpha)
    assert np.array_equal(da.view(np.uint32), da2.view(np.uint32))
    assert np.array_equal(mean.view(np.uint32), mean2.view(np.uint32))
    assert da[3, 4] == -1 and mean[3, 4] == 0 and da[5, 6] == -1 and mean[5, 6] == 0
    assert mean[7, 8] > 0 and (da > 0).mean() > 0.9
    da0, mean0 = oracle_lib.ampdispersion_block(slc, None)       # no calibration: constant amplitude -> sigma == 0 -> -1
    assert da0[7, 8] == -1 and mean0[7, 8] == 2


def test_binding_and_error_codes(tmp_path):
    a = ampdispersionlib.Ampdispersion()
    assert (a.blocksize, a.memsize, a.refband) == (64, 256, 1)
    slc = synth.make_stack(4, 10, 12, seed=2, region=4)
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc)
    a.inputDS, a.outputDS, a.meanampDS = str(tmp_path / "nope.vrt"), str(tmp_path / "da"), str(tmp_path / "mean")
    with pytest.raises(RuntimeError, match="102"):               # ampdispersion.cpp:37-46
        a.run()
    a.inputDS = vrt
    a.refband = 7
    with pytest.raises(RuntimeError, match="102"):               # ampdispersion.cpp:108-115
        a.run()
    a.refband = 1
    if engine.device_count() == 0:
        with pytest.raises(RuntimeError, match="204"):
            a.run()
        assert not os.path.exists(str(tmp_path / "da"))
