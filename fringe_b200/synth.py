"""Seeded synthetic SLC stacks for parity tests, smoke() and bench.py.

Statistical model follows the reference's own end-to-end test (tests/evd/test_evd.py:14-71):
a per-region coherence matrix

    Gamma_ij = ((g0 - ginf) * exp(-|t_i - t_j| / tau) + ginf) * exp(1j * (phi_i - phi_j))

and circular-Gaussian samples  z = Gamma^(1/2) @ CN(0, 1).  Unlike that test (unit amplitudes,
one homogeneous window) the image here is a checkerboard of regions with different Rayleigh
scale, coherence and phase history, so that the SHP tests have something to separate and the
temporal coherence spans a wide range; a small fraction of pixels is zeroed in one date to
exercise the validity mask.
"""
from __future__ import annotations

import numpy as np

# (amplitude scale, gamma0, gamma_inf, linear rate rad/yr, seasonal k)
REGION_TYPES = (
    (1.0, 0.60, 0.10, 1.0, 1),
    (2.0, 0.95, 0.30, -2.0, 2),
    (4.0, 0.80, 0.20, 3.0, 1),
    (1.5, 0.70, 0.15, 0.5, 2),
)


def phase_series(n: int, dt: float = 12.0, rate: float = 1.0, k: int = 1) -> np.ndarray:
    """Deterministic wrapped phase history (tests/evd/test_evd.py:50-71 without the noise term)."""
    t = np.arange(n) * dt
    ph = rate * t / 365.0
    if k > 0:
        ph = ph + np.sin(2 * np.pi * k * t / 365.0) + np.cos(2 * np.pi * k * t / 365.0)
    ph = ph - ph[0]
    return np.angle(np.exp(1j * ph))


def coherence_matrix(n: int, g0: float, ginf: float, tau: float, phase: np.ndarray,
                     dt: float = 12.0) -> np.ndarray:
    t = np.arange(n) * dt
    dtm = np.abs(t[:, None] - t[None, :])
    gam = (g0 - ginf) * np.exp(-dtm / tau) + ginf
    np.fill_diagonal(gam, 1.0)
    return gam * np.exp(1j * (phase[:, None] - phase[None, :]))


def matrix_sqrt(G: np.ndarray) -> np.ndarray:
    w, v = np.linalg.eigh(G)
    w = np.clip(w, 0.0, None)
    return (v * np.sqrt(w)) @ v.conj().T


def region_map(lines: int, cols: int, region: int) -> np.ndarray:
    r = (np.arange(lines) // region)[:, None]
    c = (np.arange(cols) // region)[None, :]
    return ((r * 2 + c + (r // 2)) % len(REGION_TYPES)).astype(np.int32)


def make_stack(bands: int, lines: int, cols: int, seed: int = 0, region: int = 64,
               tau: float = 72.0, dt: float = 12.0, zero_fraction: float = 0.01,
               return_truth: bool = False):
    """Return slc (bands, lines, cols) complex64 [, truth dict]."""
    rng = np.random.default_rng(seed)
    rmap = region_map(lines, cols, region)
    slc = np.empty((bands, lines, cols), np.complex64)
    phases = []
    for k, (sigma, g0, ginf, rate, seas) in enumerate(REGION_TYPES):
        ph = phase_series(bands, dt, rate, seas)
        phases.append(ph)
        sel = rmap == k
        n = int(sel.sum())
        if n == 0:
            continue
        L = matrix_sqrt(coherence_matrix(bands, g0, ginf, tau, ph, dt))
        noise = (rng.standard_normal((bands, n)) + 1j * rng.standard_normal((bands, n))) / np.sqrt(2)
        slc[:, sel] = (sigma * (L @ noise)).astype(np.complex64)
    if zero_fraction > 0:
        nz = int(zero_fraction * lines * cols)
        rr = rng.integers(0, lines, nz)
        cc = rng.integers(0, cols, nz)
        bb = rng.integers(0, bands, nz)
        slc[bb, rr, cc] = 0
    if return_truth:
        return slc, {"region": rmap, "phase": np.array(phases)}
    return slc


def make_stack_torch(bands: int, lines: int, cols: int, seed: int, device, region: int = 64,
                     tau: float = 72.0, dt: float = 12.0, zero_fraction: float = 0.01,
                     rows_per_chunk: int = 64, row_range=None):
    """Same model generated on `device` with torch (bench-size stacks; not bit-identical to
    make_stack).  Random numbers are drawn per 64-row chunk of the *global* image with a seed
    derived from the chunk index, so any `row_range=(r0, r1)` sub-block (a rank's tile + halo)
    holds exactly the rows the full image would.  Returns (bands, r1-r0, cols) complex64."""
    import torch

    r_lo, r_hi = (0, lines) if row_range is None else row_range
    out = torch.empty((bands, r_hi - r_lo, cols), dtype=torch.complex64, device=device)
    rmap = torch.from_numpy(region_map(lines, cols, region)).to(device)
    roots = []
    for (sigma, g0, ginf, rate, seas) in REGION_TYPES:
        L = sigma * matrix_sqrt(coherence_matrix(bands, g0, ginf, tau, phase_series(bands, dt, rate, seas), dt))
        roots.append(torch.from_numpy(L.astype(np.complex64)).to(device))
    roots = torch.stack(roots)                       # (T, bands, bands)
    g = torch.Generator(device=device)
    for c in range(r_lo // rows_per_chunk, (r_hi + rows_per_chunk - 1) // rows_per_chunk):
        c0, c1 = c * rows_per_chunk, min(lines, (c + 1) * rows_per_chunk)
        g.manual_seed(seed * 1000003 + c)
        n = (c1 - c0) * cols
        re = torch.randn((bands, n), generator=g, device=device)
        im = torch.randn((bands, n), generator=g, device=device)
        noise = torch.complex(re, im) * (0.5 ** 0.5)
        del re, im
        kinds = rmap[c0:c1].reshape(-1)
        chunk = torch.empty((bands, n), dtype=torch.complex64, device=device)
        for k in range(roots.shape[0]):
            sel = kinds == k
            if bool(sel.any()):
                chunk[:, sel] = roots[k] @ noise[:, sel]
        if zero_fraction > 0:
            nz = int(zero_fraction * n)
            pp = torch.randint(0, n, (nz,), generator=g, device=device)
            bb = torch.randint(0, bands, (nz,), generator=g, device=device)
            chunk[bb, pp] = 0
        chunk = chunk.reshape(bands, c1 - c0, cols)
        a, b = max(c0, r_lo), min(c1, r_hi)
        out[:, a - r_lo:b - r_lo, :] = chunk[:, a - c0:b - c0, :]
    return out
