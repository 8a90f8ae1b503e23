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


def test_integrate_ps_oracle_equals_the_reference_scripts_output():
    """tests/golden/integrate_ps_24x40.npz holds what the reference's own python/integratePS.py (integratePS2DS,
    getCoherence) produced for these inputs (tests/golden/make_golden_integrate_ps.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "integrate_ps_24x40.npz"))
    o = oracle.load()
    for j in (1, 2, 3):
        got = o.integrate_ps(g["ds"][0], g["ds"][j], g["slc"][0], g["slc"][j], g["ps"])
        want = g[f"ifg_0_{j}"]
        assert np.abs(got - want).max() <= 1e-6
        # (numpy's complex64 multiply is SIMD code that may contract into FMAs: its last bit is the machine's, so the
        # DS pixels are compared to an ulp, not bit for bit)
        assert np.abs(got - want)[g["ps"] != 1].max() <= 1.5e-7
    got = o.integrate_ps(g["ds"][0], g["ds"][2], g["slc"][0], g["slc"][2], g["ps"])
    assert got[7, 9].real == -1.0 and g["ifg_0_2"][7, 9].real == -1.0        # product on the negative real axis: angle = pi
    assert np.array_equal(np.where(g["ps"] == 1, np.float32(0.95), g["tcorr"]), g["coherence"])
