"""Datum adjustment, CPU side: the oracle's product against an independent numpy
evaluation, and the date -> file mapping of python/adjustMiniStacks.py:66-101."""
import os

import numpy as np

from fringe_b200.cli import adjust_ministacks as adj


def test_oracle_cmul_matches_double_arithmetic(oracle_lib):
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(5001) + 1j * rng.standard_normal(5001)).astype(np.complex64)
    b = np.exp(1j * rng.uniform(-np.pi, np.pi, 5001)).astype(np.complex64)
    a[:4] = [0, 1, 1j, np.float32(3.0e38)]                    # zeros, units, overflow to inf in float
    b[:4] = [1, 0, 1j, np.float32(3.0e38)]
    got = oracle_lib.cmul(a, b)
    ar, ai, br, bi = (x.astype(np.float64) for x in (a.real, a.imag, b.real, b.imag))
    with np.errstate(over="ignore", invalid="ignore"):
        want = ((ar * br - ai * bi).astype(np.float32) + 1j * (ar * bi + ai * br).astype(np.float32)).astype(np.complex64)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_stack_dict_mapping(tmp_path):
    dates = ["20200101", "20200113", "20200125", "20200206", "20200218"]
    d = adj.getStackDict(dates, str(tmp_path / "mini"), str(tmp_path / "datum"), str(tmp_path / "out"), 2)
    assert list(d) == dates
    mini = os.path.join(str(tmp_path / "mini"), "20200101_20200113", "EVD", "20200113.slc")
    assert d["20200113"][0] == mini
    assert d["20200113"][1] == os.path.join(str(tmp_path / "datum"), "EVD", "20200113.slc")     # datum = last date of the ministack
    assert d["20200101"][1] == d["20200113"][1]
    assert d["20200218"][0].endswith(os.path.join("20200218_20200218", "EVD", "20200218.slc"))   # ragged last ministack
    assert d["20200125"][3] == os.path.join(str(tmp_path / "out"), "20200125.slc")


def test_get_dates_sorted(tmp_path):
    for n in ("20200301", "20200101", "20200201"):
        (tmp_path / (n + ".vrt")).write_text("x")
    (tmp_path / "notes.txt").write_text("x")
    assert adj.getDates(str(tmp_path)) == ["20200101", "20200201", "20200301"]
