"""despeck on the GPU (fringe_despeck_block / _device) against the oracle: bit-identical in every
mode, with block edges, sub-ranges of lines, wide windows and zeroed pixels."""
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
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.parametrize("Nx,Ny", [(5, 2), (3, 3), (10, 10)])
@pytest.mark.parametrize("mode", ["amplitude", "ifg", "coherence", "amplitude+coherence"])
def test_despeck_bit_exact(ctx, oracle_lib, mode, Nx, Ny):
    slc = synth.make_stack(6, 40, 56, seed=3 + Nx, region=16)
    wts = oracle_lib.nmap_block(slc, Nx, Ny)[1]
    z2 = None if mode.startswith("amplitude") else slc[4]
    coh = mode.endswith("coherence")
    want = oracle_lib.despeck_block(slc[1], wts, Nx, Ny, z2=z2, coherence=coh)
    got = ctx.despeck_block(slc[1], wts, Nx, Ny, z2=z2, coherence=coh)
    assert _same(got, want)
    assert want.any() != (mode == "amplitude+coherence")


def test_despeck_line_range_and_device(ctx, oracle_lib):
    import torch
    slc = synth.make_stack(5, 33, 47, seed=11, region=16)
    wts = oracle_lib.nmap_block(slc, 4, 2)[1]
    want = oracle_lib.despeck_block(slc[0], wts, 4, 2, z2=slc[3], coherence=True, first_line=2, n_lines=20)
    got = ctx.despeck_block(slc[0], wts, 4, 2, z2=slc[3], coherence=True, first_line=2, n_lines=20)
    assert _same(got, want)
    dz1, dz2 = torch.from_numpy(slc[0]).cuda(), torch.from_numpy(slc[3]).cuda()
    dw = torch.from_numpy(wts.view(np.int32)).cuda()
    dgot = ctx.despeck_block_device(dz1, dw, 4, 2, z2=dz2, coherence=True, first_line=2, n_lines=20).cpu().numpy()
    assert _same(dgot, want)
    assert ctx.last_kernel_ms("despeck") > 0


def test_despeck_argument_errors(ctx):
    z = np.zeros((4, 4), np.complex64)
    w = np.zeros((4, 4, 1), np.uint32)
    with pytest.raises(Exception):
        ctx.despeck_block(z, w, 1, 1, first_line=3, n_lines=4)


def test_post_rows_against_golden_fixtures(ctx):
    """GPU results of the three SURVEY 8f rows against the committed fixtures, bit for bit."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "block_12x24x40.npz"))
    p = np.load(os.path.join(gold, "post_12x24x40.npz"))
    slc, wts = g["slc"], g["wts_ks"]
    da, mean = ctx.ampdispersion_block(slc, p["alpha"])
    assert _same(da, p["ampdisp_da"]) and _same(mean, p["ampdisp_mean"])
    assert _same(ctx.despeck_block(slc[2], wts, 5, 2), p["despeck_amp"])
    assert _same(ctx.despeck_block(slc[2], wts, 5, 2, z2=slc[9]), p["despeck_ifg"])
    assert _same(ctx.despeck_block(slc[2], wts, 5, 2, z2=slc[9], coherence=True), p["despeck_coh"])
    assert _same(ctx.cmul(slc[4], slc[7]), p["cmul"])
