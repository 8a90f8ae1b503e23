"""The command lines keep the reference scripts' option letters, destinations and defaults
(src/nmap/nmap.py:6-36, src/evd/evd.py:6-34, src/sequential/sequential.py:17-52, src/despeck/despeck.py:6-33,
src/ampdispersion/ampdispersion.py:6-30, python/adjustMiniStacks.py:16-50)."""
import pytest

from fringe_b200.cli import adjust_ministacks, ampdispersion, despeck, evd, nmap, phase_link, sequential


def test_nmap_defaults_and_flags():
    a = nmap.cmdLineParser(["-i", "s.vrt", "-o", "w", "-c", "c"])
    assert (a.inputDS, a.outputDS, a.countDS, a.maskDS) == ("s.vrt", "w", "c", "")
    assert (a.linesPerBlock, a.memorySize, a.halfWindowX, a.halfWindowY, a.pValue, a.method, a.noGPU) == (64, 256, 5, 5, 0.05, "KS2", False)
    b = nmap.cmdLineParser(["--input", "s", "--output", "w", "--count", "c", "--mask", "m", "-l", "32", "-r", "9", "-x", "11",
                            "-y", "3", "-p", "0.1", "-s", "ad2", "--nogpu"])
    assert (b.maskDS, b.linesPerBlock, b.memorySize, b.halfWindowX, b.halfWindowY, b.pValue, b.method, b.noGPU) == ("m", 32, 9, 11, 3, 0.1, "ad2", True)
    with pytest.raises(SystemExit):
        nmap.cmdLineParser(["-i", "s.vrt"])


def test_evd_and_phase_link_defaults():
    a = evd.cmdLineParser(["-i", "s.vrt", "-w", "w", "-o", "out"])
    assert (a.linesPerBlock, a.memorySize, a.halfWindowX, a.halfWindowY, a.minNeighbors, a.method, a.bandWidth) == (64, 2048, 5, 5, 5, "MLE", -1)
    b = evd.cmdLineParser(["-i", "s", "-w", "w", "-o", "o", "-m", "STBAS", "-b", "4", "-n", "7"])
    assert (b.method, b.bandWidth, b.minNeighbors) == ("STBAS", 4, 7)
    assert phase_link.main.__module__ == "fringe_b200.cli.phase_link"


def test_sequential_defaults():
    a = sequential.cmdLineParser(["-i", "SLC", "-w", "w", "-o", "out"])
    assert (a.linesPerBlock, a.memorySize, a.halfWindowX, a.halfWindowY, a.minNeighbors, a.miniStackSize, a.forceprocessing, a.bbox) == (64, 2048, 29, 9, 5, 10, False, None)
    b = sequential.cmdLineParser(["-i", "SLC", "-w", "w", "-o", "out", "-b", "1", "2", "3", "4", "-s", "5", "-f"])
    assert b.bbox == ["1", "2", "3", "4"] and b.miniStackSize == 5 and b.forceprocessing


def test_despeck_ampdispersion_adjust_defaults():
    d = despeck.cmdLineParser(["-i", "s", "-o", "o", "-w", "w"])
    assert (d.linesPerBlock, d.memorySize, d.halfWindowX, d.halfWindowY, d.bands, d.cohFlag) == (64, 512, 5, 5, [], False)
    d = despeck.cmdLineParser(["-i", "s", "-o", "o", "-w", "w", "-b", "2", "5", "-c"])
    assert d.bands == [2, 5] and d.cohFlag
    a = ampdispersion.cmdLineParser(["-i", "s", "-o", "da"])
    assert (a.meanampDS, a.linesPerBlock, a.memorySize, a.refBand) == ("", 64, 256, 1)
    m = adjust_ministacks.cmdLineParser(["-s", "slcs", "-m", "mini", "-d", "datum", "-M", "10", "-o", "out"])
    assert (m.slcDir, m.miniStackDir, m.datumDir, m.miniStackSize, m.outDir, m.unwrapped) == ("slcs", "mini", "datum", 10, "out", False)


def test_despeck_band_rules():
    with pytest.raises(Exception, match="coherence"):
        despeck.main(["-i", "s", "-o", "o", "-w", "w", "-b", "1", "-c"])
    with pytest.raises(Exception, match="More than two"):
        despeck.main(["-i", "s", "-o", "o", "-w", "w", "-b", "1", "2", "3"])
