"""Builds the C++ host side: libfringe_host.so (block drivers + raster I/O over the C ABI) and the
Python extension modules nmaplib / evdlib / phase_linklib / despecklib / ampdispersionlib (pybind11), all in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sysconfig

import pybind11

from .build import CSRC, HERE, LIBDIR, HOST_CXX, _newer, _run

HOST = os.path.join(CSRC, "host")
BINDINGS = os.path.join(HERE, "bindings")
HOST_LIB = os.path.join(LIBDIR, "libfringe_host.so")


def build(force: bool = False) -> None:
    os.makedirs(BINDINGS, exist_ok=True)
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")]
    common = ["-O2", "-std=c++17", "-fPIC", "-pthread", "-Wall"]
    gdal_link = []
    gdal_config = shutil.which("gdal-config")
    if gdal_config and not os.environ.get("FRINGE_NO_GDAL"):          # optional GDAL backend of host/raster_io.hpp
        common += ["-DFRINGE_WITH_GDAL"] + subprocess.check_output([gdal_config, "--cflags"], text=True).split()
        gdal_link = subprocess.check_output([gdal_config, "--libs"], text=True).split()
    if force or _newer(HOST_LIB, [os.path.join(HOST, "drivers.cpp"), __file__] + hdrs):
        _run([HOST_CXX] + common + ["-shared", "-o", HOST_LIB, os.path.join(HOST, "drivers.cpp"),
                                    "-L" + LIBDIR, "-lfringe_b200", "-Wl,-rpath,$ORIGIN"] + gdal_link)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
    link = ["-L" + LIBDIR, "-lfringe_host", "-lfringe_b200", "-Wl,-rpath,$ORIGIN/../lib"]
    for mod, src, defs in (("nmaplib", "nmaplib.cpp", []), ("evdlib", "evdlib.cpp", []),
                           ("phase_linklib", "evdlib.cpp", ["-DFRINGE_PHASE_LINK"]),
                           ("despecklib", "despecklib.cpp", []),
                           ("ampdispersionlib", "ampdispersionlib.cpp", []),
                           ("calamplib", "calamplib.cpp", [])):
        out = os.path.join(BINDINGS, mod + ext)
        if force or _newer(out, [os.path.join(HOST, src), HOST_LIB, __file__] + hdrs):
            _run([HOST_CXX] + common + ["-fvisibility=hidden", "-shared"] + defs + inc +
                 ["-o", out, os.path.join(HOST, src)] + link)
