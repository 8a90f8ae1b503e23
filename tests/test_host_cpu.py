"""CPU: host-side logic of the drop-in boundary that needs no device -- the binding class surface
(names / defaults of nmaplib.pyx, evdlib.cpp, phase_linklib.pyx), the VRT / ENVI raster layer, the
reference's error codes on the paths that fail before any compute, and the explicit
"no device, no CPU path" status."""
import os

import numpy as np
import pytest

from fringe_b200 import engine, stackio, synth
from fringe_b200.cli._common import use_bindings
from fringe_b200.partition import block_schedule

use_bindings()
import evdlib  # noqa: E402
import nmaplib  # noqa: E402
import phase_linklib  # noqa: E402

NO_GPU = engine.device_count() == 0


def test_binding_surface_and_defaults():
    a = nmaplib.Nmap()
    for name in ("inputDS", "countDS", "weightsDS", "maskDS", "method", "minimumProbability", "blocksize",
                 "memsize", "halfWindowX", "halfWindowY", "noGPU", "print", "run"):
        assert hasattr(a, name), name
    # nmap.hpp:56-65
    assert (a.blocksize, a.memsize, a.halfWindowX, a.halfWindowY, a.method, a.minimumProbability, a.noGPU) == \
           (64, 512, 5, 5, "KS2", 0.05, False)
    for cls in (evdlib.Evd, phase_linklib.Phaselink):
        e = cls()
        for name in ("inputDS", "outputFolder", "outputCompressedSlcFolder", "compSlc", "weightsDS",
                     "minimumNeighbors", "miniStackCount", "blocksize", "memsize", "halfWindowX", "halfWindowY",
                     "method", "bandWidth", "print", "run"):
            assert hasattr(e, name), name
        # evd.hpp:58-68
        assert (e.blocksize, e.memsize, e.halfWindowX, e.halfWindowY, e.minimumNeighbors, e.miniStackCount,
                e.method, e.bandWidth) == (64, 2048, 5, 5, 2, 1, "MLE", -1)


def test_envi_and_vrt_roundtrip(tmp_path):
    slc = synth.make_stack(4, 10, 12, seed=2, region=4)
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc, extra_md={"20200113": {"amplitudeConstant": "2.5"}})
    txt = open(vrt).read()
    assert txt.count("<VRTRasterBand") == 4 and 'domain="slc"' in txt and "amplitudeConstant" in txt
    assert stackio.raster_size(vrt) == (12, 10)
    d = stackio.default_dates(4)[1]
    back = stackio.read_envi(os.path.join(str(tmp_path), "SLC", d, d + ".slc"))
    assert np.array_equal(back, slc[1])
    w = np.arange(10 * 12 * 2, dtype=np.uint32).reshape(10, 12, 2)
    stackio.write_envi(str(tmp_path / "w"), w, {"HALFWINDOWX": 5, "HALFWINDOWY": 2})
    assert np.array_equal(stackio.read_envi(str(tmp_path / "w")), w)
    assert stackio.read_envi_header(str(tmp_path / "w"))["halfwindowx"] == "5"


def test_error_codes_before_compute(tmp_path):
    slc = synth.make_stack(4, 10, 12, seed=2, region=4)
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc)
    a = nmaplib.Nmap()
    a.inputDS, a.weightsDS, a.countDS = vrt, str(tmp_path / "w"), str(tmp_path / "c")
    a.method = "XYZ"
    with pytest.raises(RuntimeError, match="returned 1$"):       # nmap.cpp:41
        a.run()
    a.method = "KS2"
    a.inputDS = str(tmp_path / "nope.vrt")
    with pytest.raises(RuntimeError, match="102"):               # nmap.cpp:80
        a.run()
    a.inputDS = vrt
    stackio.write_envi(str(tmp_path / "badmask"), np.ones((10, 11), np.uint8))
    a.maskDS = str(tmp_path / "badmask")
    with pytest.raises(RuntimeError, match="104"):               # nmap.cpp:132-136
        a.run()
    e = evdlib.Evd()
    e.inputDS, e.weightsDS, e.outputFolder = vrt, str(tmp_path / "missing_wts"), str(tmp_path / "out")
    with pytest.raises(RuntimeError, match="105"):               # evd.cpp:95-101
        e.run()
    stackio.write_envi(str(tmp_path / "wts"), np.zeros((10, 12, 1), np.uint32), {"HALFWINDOWX": 2, "HALFWINDOWY": 2})
    e.weightsDS = str(tmp_path / "wts")
    e.halfWindowX, e.halfWindowY = 2, 2
    e.method, e.bandWidth = "STBAS", -1
    with pytest.raises(RuntimeError, match="101"):               # evd.cpp:74-91
        e.run()
    e.method = "EVD"
    e.halfWindowX = 5                                            # nulong mismatch -> 108 ... window mismatch -> 109
    with pytest.raises(RuntimeError, match="10[89]"):
        e.run()


@pytest.mark.skipif(not NO_GPU, reason="a CUDA device is present")
def test_driver_reports_no_device_instead_of_falling_back(tmp_path):
    slc = synth.make_stack(4, 10, 12, seed=2, region=4)
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc)
    a = nmaplib.Nmap()
    a.inputDS, a.weightsDS, a.countDS = vrt, str(tmp_path / "w"), str(tmp_path / "c")
    with pytest.raises(RuntimeError, match="204"):               # 200 + FRINGE_ERR_NO_DEVICE
        a.run()


def test_block_schedule_matches_reference_rules():
    # nmap.cpp:487-573: first block writes [0,H-Ny), rolls back Ny; middle [Ny,H-Ny); last [Ny,H)
    for rows, h, ny in [(150, 64, 2), (1500, 64, 5), (64, 64, 2), (10, 64, 5), (129, 64, 2)]:
        written = np.zeros(rows, int)
        for yoff, n, first, nwrite in block_schedule(rows, h, ny):
            assert 0 <= yoff and yoff + n <= rows and n <= h
            lo, hi = yoff + first, yoff + first + nwrite
            written[lo:hi] += 1
            # every written line has its full +-Ny window inside the block or at the image edge
            assert lo - ny >= yoff or lo == 0
            assert hi + ny <= yoff + n or hi == rows
        assert np.all(written == 1)


def test_despeck_binding_and_error_codes(tmp_path):
    """despecklib.Despeck (src/despeck/despecklib.pyx surface) and the checks of despeck.cpp:40-150 that
    run before any device work."""
    import despecklib
    d = despecklib.Despeck()
    assert (d.band1, d.band2, d.coherenceFlag, d.blocksize, d.memsize, d.halfWindowX, d.halfWindowY) == (1, -1, False, 64, 512, 5, 5)
    slc = synth.make_stack(4, 10, 12, seed=2, region=4)
    vrt = stackio.make_stack_on_disk(str(tmp_path), slc)
    d.inputDS, d.weightsDS, d.outputDS = str(tmp_path / "nope.vrt"), str(tmp_path / "w"), str(tmp_path / "out")
    with pytest.raises(RuntimeError, match="102"):               # despeck.cpp:43-50
        d.run()
    d.inputDS = vrt
    d.band1 = 9
    with pytest.raises(RuntimeError, match="102"):               # despeck.cpp:61-67
        d.run()
    d.band1, d.band2 = 1, 7
    with pytest.raises(RuntimeError, match="102"):               # despeck.cpp:70-76
        d.run()
    d.band2 = 2
    with pytest.raises(RuntimeError, match="105"):               # despeck.cpp:82-88
        d.run()
    stackio.write_envi(str(tmp_path / "w"), np.zeros((10, 12, 1), np.uint32), {"HALFWINDOWX": 2, "HALFWINDOWY": 2})
    with pytest.raises(RuntimeError, match="110"):               # 11 x 11 window: 108, 109 and 110 all trip, the last one wins
        d.run()
    d.halfWindowX, d.halfWindowY = 3, 2
    with pytest.raises(RuntimeError, match="109"):               # despeck.cpp:117-121 (after the band-count check 108)
        d.run()
    d.halfWindowX = 2
    if engine.device_count() == 0:
        with pytest.raises(RuntimeError, match="204"):           # no CPU path
            d.run()
        assert not os.path.exists(str(tmp_path / "out"))
