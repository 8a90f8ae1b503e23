"""Datum adjustment on the GPU (fringe_cmul / fringe_cmul_device) against the oracle, bit for bit,
and the adjust_ministacks CLI on a small ministack tree."""
import os

import numpy as np
import pytest

from fringe_b200 import stackio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from fringe_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _pair(n, seed):
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    b = np.exp(1j * rng.uniform(-np.pi, np.pi, n)).astype(np.complex64)
    return a, b


@pytest.mark.parametrize("n", [0, 1, 2, 7, 4096, 100003])
def test_cmul_host_bit_exact(ctx, oracle_lib, n):
    a, b = _pair(n, n)
    got = ctx.cmul(a, b)
    want = oracle_lib.cmul(a, b) if n else np.empty(0, np.complex64)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_cmul_device_and_special_values(ctx, oracle_lib):
    import torch
    a, b = _pair(33333, 5)
    a[:6] = [0, np.inf, np.float32(3e38), np.nan, 1e-30, -0.0]
    b[:6] = [np.inf, 1, np.float32(3e38), 1, 1e-30, 1j]
    with np.errstate(all="ignore"):
        want = oracle_lib.cmul(a, b)
    got = ctx.cmul_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    same = got.view(np.uint32) == want.view(np.uint32)
    both_nan = np.isnan(got.view(np.float32)) & np.isnan(want.view(np.float32))
    assert np.all(same | both_nan)
    assert ctx.last_kernel_ms("cmul") > 0


def test_cli_on_ministack_tree(ctx, oracle_lib, tmp_path):
    from fringe_b200.cli import adjust_ministacks as adj
    dates = stackio.default_dates(5)
    lines, cols = 9, 13
    rng = np.random.default_rng(3)
    slcdir = tmp_path / "slcs"
    slcdir.mkdir()
    phasor = lambda: np.exp(1j * rng.uniform(-np.pi, np.pi, (lines, cols))).astype(np.complex64)
    mini, datum = {}, {}
    for i0 in range(0, 5, 2):
        grp = dates[i0:i0 + 2]
        d = tmp_path / "mini" / (grp[0] + "_" + grp[-1]) / "EVD"
        d.mkdir(parents=True)
        for dd in grp:
            mini[dd] = phasor()
            stackio.write_envi(str(d / (dd + ".slc")), mini[dd])
            (slcdir / (dd + ".vrt")).write_text("<VRTDataset/>")
        (tmp_path / "datum" / "EVD").mkdir(parents=True, exist_ok=True)
        datum[grp[-1]] = phasor()
        stackio.write_envi(str(tmp_path / "datum" / "EVD" / (grp[-1] + ".slc")), datum[grp[-1]])
    rc = adj.main(["-s", str(slcdir), "-m", str(tmp_path / "mini"), "-d", str(tmp_path / "datum"), "-M", "2",
                   "-o", str(tmp_path / "out")])
    assert rc == 0
    for i, dd in enumerate(dates):
        last = dates[min(i // 2 * 2 + 1, 4)]
        got = stackio.read_envi(str(tmp_path / "out" / (dd + ".slc")))
        want = oracle_lib.cmul(mini[dd], datum[last])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert stackio.raster_size(str(tmp_path / "out" / (dd + ".slc.vrt"))) == (cols, lines)
