"""ampdispersion on the GPU against the oracle: bit-identical through the C ABI (host and device
variants) and through the file-level driver with calibration constants in the VRT metadata."""
import os

import numpy as np
import pytest

from fringe_b200 import stackio, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from fringe_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


def test_abi_bit_exact(ctx, oracle_lib):
    import torch
    slc = synth.make_stack(11, 37, 53, seed=5, region=16, zero_fraction=0.15)
    slc[:, 3, 4] = 0
    slc[1:, 5, 6] = 0
    slc[:, 7, 8] = 2 + 0j
    alpha = np.linspace(1.0, 1.5, 11)
    for al in (None, alpha):
        da, mean = oracle_lib.ampdispersion_block(slc, al)
        gda, gmean = ctx.ampdispersion_block(slc, al)
        assert _same(gda, da) and _same(gmean, mean)
    dal = torch.from_numpy(alpha).cuda()
    dda, dmean = ctx.ampdispersion_block_device(torch.from_numpy(slc).cuda(), dal)
    assert _same(dda.cpu().numpy(), da) and _same(dmean.cpu().numpy(), mean)
    assert ctx.last_kernel_ms("ampdispersion") > 0


def test_cli_with_calibration_metadata(oracle_lib, tmp_path):
    from fringe_b200.cli import ampdispersion as cli
    slc = synth.make_stack(6, 150, 40, seed=9, region=16)
    dates = stackio.default_dates(6)
    consts = {d: {"amplitudeConstant": 2.0 + 0.25 * i} for i, d in enumerate(dates)}
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc, extra_md=consts)
    out, mean = str(tmp_path / "ps" / "da"), str(tmp_path / "ps" / "mean")
    cli.main(["-i", vrt, "-o", out, "-m", mean, "-b", "3", "-r", "1", "-l", "32"])     # several blocks
    da, m = stackio.read_envi(out), stackio.read_envi(mean)
    assert da.dtype == np.float32 and da.shape == (150, 40)
    assert stackio.read_envi_header(out)["n"] == "6"
    alpha = np.array([2.0 + 0.25 * i for i in range(6)])
    alpha = alpha / alpha[2]
    alpha[2] = 1.0
    rda, rm = oracle_lib.ampdispersion_block(slc, alpha)
    assert _same(da, rda) and _same(m, rm)
