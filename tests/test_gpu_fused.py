"""fringe_nmap_evd_block (one upload, mask stays on the device) against the two separate calls:
bit-identical outputs, including sub-ranges of lines, masks/calibration, and several pipeline
chunks per block."""
import os

import numpy as np
import pytest

from fringe_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from fringe_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8))


@pytest.mark.parametrize("chunk", [None, "7"])
@pytest.mark.parametrize("case", [
    dict(bands=12, lines=40, cols=64, Nx=5, Ny=2, nmap_method="KS2", method="EVD"),
    dict(bands=30, lines=33, cols=50, Nx=3, Ny=3, nmap_method="AD2", method="STBAS", bandwidth=5),
    dict(bands=9, lines=30, cols=41, Nx=4, Ny=2, nmap_method="KS2", method="MLE", first_line=4, n_lines=19,
         mini_stack_count=3),
    dict(bands=10, lines=28, cols=36, Nx=2, Ny=4, nmap_method="KS2", method="MLE", variant=1, min_neighbors=5,
         with_mask=True),
])
def test_fused_equals_two_calls(ctx, case, chunk, monkeypatch):
    if chunk:
        monkeypatch.setenv("FRINGE_CHUNK_ROWS", chunk)
    else:
        monkeypatch.delenv("FRINGE_CHUNK_ROWS", raising=False)
    c = dict(case)
    bands, lines, cols = c.pop("bands"), c.pop("lines"), c.pop("cols")
    Nx, Ny = c.pop("Nx"), c.pop("Ny")
    nmap_method = c.pop("nmap_method")
    with_mask = c.pop("with_mask", False)
    slc = synth.make_stack(bands, lines, cols, seed=11)
    mask = alpha = None
    if with_mask:
        rng = np.random.default_rng(5)
        mask = (rng.random((lines, cols)) > 0.1).astype(np.uint8)
        alpha = np.linspace(1.0, 1.3, bands)
    count, wts = ctx.nmap_block(slc, Nx, Ny, nmap_method, 0.05, mask=mask, alpha=alpha)
    out, tcorr, comp = ctx.evd_block(slc, wts, Nx, Ny, **c)
    fcount, fwts, fout, ftcorr, fcomp = ctx.nmap_evd_block(slc, Nx, Ny, nmap_method, 0.05, mask=mask, alpha=alpha, **c)
    assert _same(count, fcount) and _same(wts, fwts)
    assert _same(out, fout) and _same(tcorr, ftcorr) and _same(comp, fcomp)
    # without the mask copy-back the solve is unchanged
    n1, n2, gout, gtcorr, gcomp = ctx.nmap_evd_block(slc, Nx, Ny, nmap_method, 0.05, mask=mask, alpha=alpha,
                                                     want_mask=False, **c)
    assert n1 is None and n2 is None
    assert _same(out, gout) and _same(tcorr, gtcorr) and _same(comp, gcomp)


def test_fused_argument_errors(ctx):
    slc = synth.make_stack(6, 12, 16, seed=1)
    with pytest.raises(Exception):
        ctx.nmap_evd_block(slc, 2, 2, "KS2", 0.05, method="STBAS", bandwidth=99)
    with pytest.raises(Exception):
        ctx.nmap_evd_block(slc, 2, 2, "KS2", 0.05, first_line=10, n_lines=5)
