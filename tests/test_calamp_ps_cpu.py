"""CPU: the oracle's restatements of calamp (src/calamp/calamp.cpp:207-243) and of the PS / DS integration
(python/integratePS.py:97-159) against independent numpy evaluations of the same expressions.  The reference holds no
vectors for either (its drivers need GDAL / isce): parity for these two rows is pinned to numpy only."""
import numpy as np

import oracle
from fringe_b200 import synth


def test_calamp_oracle_equals_numpy():
    o = oracle.load()
    slc = synth.make_stack(6, 40, 50, seed=3, region=16)
    slc[2, 5, 7] = np.nan
    mask = (np.random.default_rng(1).random((40, 50)) > 0.2).astype(np.uint8)
    for m in (None, mask):
        sums, counts = o.calamp_block(slc, m)
        amp = np.abs(slc).astype(np.float64)
        amp[np.isnan(amp)] = 0
        valid = (amp != 0) & (True if m is None else (m > 0)[None])
        assert np.array_equal(counts, valid.sum(axis=(1, 2)))
        # numpy's float32 hypot and glibc's differ by an ulp on a few samples: float-level agreement of the sums
        assert np.allclose(sums, (amp * valid).sum(axis=(1, 2)), rtol=1e-7)


def test_integrate_ps_oracle_equals_numpy():
    o = oracle.load()
    rng = np.random.default_rng(5)
    z = lambda: (rng.standard_normal((30, 40)) + 1j * rng.standard_normal((30, 40))).astype(np.complex64)
    ds_i, ds_j, slc_i, slc_j = np.exp(1j * np.angle(z())).astype(np.complex64), np.exp(1j * np.angle(z())).astype(np.complex64), z(), z()
    ps = (rng.random((30, 40)) > 0.7).astype(np.uint8)
    slc_i[3, 3] = 0
    ps[3, 3] = 1
    got = o.integrate_ps(ds_i, ds_j, slc_i, slc_j, ps)
    want = ds_j * np.conjugate(ds_i)
    ifg = np.exp(1J * np.angle(slc_j * np.conjugate(slc_i)))
    want[ps == 1] = ifg[ps == 1]
    assert np.abs(got - want).max() <= 1e-6
    assert got[3, 3] == 1.0 + 0j                                  # angle(0) = 0
