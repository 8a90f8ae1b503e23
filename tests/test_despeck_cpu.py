"""despeck, CPU side: the oracle's loop (src/despeck/despeck.cpp:321-361, 387-432)
against an independent numpy evaluation with float32 running sums in window raster order."""
import numpy as np
import pytest

from fringe_b200 import synth


def _hypotf(z):
    """glibc hypotf: (float) sqrt((double) re * re + (double) im * im); numpy's own abs on complex64 takes a
    SIMD route that differs from it by an ulp now and then."""
    re, im = z.real.astype(np.float64), z.imag.astype(np.float64)
    return np.sqrt(re * re + im * im).astype(np.float32)


def _numpy_despeck(z1, z2, wts, Nx, Ny, coherence):
    lines, cols = z1.shape
    f32 = np.float32
    if z2 is None:
        d1 = _hypotf(z1).astype(np.complex64)
        d2 = np.ones_like(z1)
    else:
        a, b = z1.astype(np.complex64), z2.astype(np.complex64)
        re = (a.real * b.real).astype(f32) + (a.imag * b.imag).astype(f32)
        im = (a.imag * b.real).astype(f32) - (a.real * b.imag).astype(f32)
        d1 = (re + 1j * im).astype(np.complex64)
        if coherence:
            p, q = _hypotf(a), _hypotf(b)
            d2 = ((p * p).astype(f32) + 1j * (q * q).astype(f32)).astype(np.complex64)
        else:
            d2 = np.ones_like(z1)
    out = np.zeros_like(z1)
    WX = 2 * Nx + 1
    center = Ny * WX + Nx
    for y in range(lines):
        for x in range(cols):
            w = wts[y, x]
            if not (w[center >> 5] >> (center & 31)) & 1:
                continue
            vr = vi = sr = si = f32(0)
            for yy in range(max(y - Ny, 0), min(lines - 1, y + Ny) + 1):
                for xx in range(max(x - Nx, 0), min(cols - 1, x + Nx) + 1):
                    f = (yy - y + Ny) * WX + xx - x + Nx
                    if (w[f >> 5] >> (f & 31)) & 1:
                        vr = f32(vr + d1[yy, xx].real); vi = f32(vi + d1[yy, xx].imag)
                        sr = f32(sr + d2[yy, xx].real); si = f32(si + d2[yy, xx].imag)
            if sr > 0:
                if coherence:
                    if si > 0:
                        den = f32(f32(np.sqrt(sr)) * f32(np.sqrt(si)))
                        out[y, x] = f32(vr / den) + 1j * f32(vi / den)
                else:
                    out[y, x] = f32(vr / sr) + 1j * f32(vi / sr)
    return out


@pytest.mark.parametrize("mode", ["amplitude", "ifg", "coherence", "amplitude+coherence"])
def test_oracle_despeck_matches_numpy(oracle_lib, mode):
    slc = synth.make_stack(4, 14, 19, seed=7, region=8)
    wts = oracle_lib.nmap_block(slc, 3, 2)[1]
    z2 = None if mode.startswith("amplitude") else slc[2]
    coh = mode.endswith("coherence")
    got = oracle_lib.despeck_block(slc[0], wts, 3, 2, z2=z2, coherence=coh)
    want = _numpy_despeck(slc[0], z2, wts, 3, 2, coh)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if mode == "amplitude+coherence":
        assert not got.any()                                 # reference quirk: weight sum has no imaginary part
    elif mode == "coherence":
        assert 0 < np.abs(got).max() <= 1.0 + 1e-6


def test_oracle_despeck_line_range(oracle_lib):
    slc = synth.make_stack(3, 12, 16, seed=9, region=8)
    wts = oracle_lib.nmap_block(slc, 2, 2)[1]
    full = oracle_lib.despeck_block(slc[0], wts, 2, 2, z2=slc[1])
    part = oracle_lib.despeck_block(slc[0], wts, 2, 2, z2=slc[1], first_line=3, n_lines=5)
    assert np.array_equal(part[3:8], full[3:8]) and not part[:3].any() and not part[8:].any()
