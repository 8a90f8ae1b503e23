#!/usr/bin/env python3
"""Golden vectors for the PS / DS integration from the REFERENCE's own python/integratePS.py.

The script is imported from /root/reference/python as it stands; its heavy imports (osgeo.gdal, isce, isceobj, Network)
are replaced by empty stand-in modules, and its functions integratePS2DS / get_fullres_ifgram / getCoherence are run on
array-backed objects that answer the few GDAL dataset calls they make (RasterXSize / RasterYSize, ReadAsArray,
GetRasterBand(b).ReadAsArray / WriteArray).  All arithmetic -- slcj * conj(slci), np.angle, np.exp(1j * ...), the masked
assignment, the 0.95 PS coherence -- is the reference's code and this container's numpy.
Run in the authoring container only:  python tests/golden/make_golden_integrate_ps.py
Writes tests/golden/integrate_ps_24x40.npz (inputs and outputs)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PY = "/root/reference/python"


class Band:
    def __init__(self, arr):
        self.arr = arr

    def ReadAsArray(self, x0=0, y0=0, xoff=None, yoff=None):
        return self.arr[y0:y0 + yoff, x0:x0 + xoff].copy()

    def WriteArray(self, data, x0=0, y0=0):
        self.arr[y0:y0 + data.shape[0], x0:x0 + data.shape[1]] = data


class Dataset:
    """(bands, lines, cols) array behind the handful of GDAL dataset calls integratePS.py makes."""

    def __init__(self, arr):
        self.arr = arr if arr.ndim == 3 else arr[None]
        self.RasterYSize, self.RasterXSize = self.arr.shape[1:]

    def GetRasterBand(self, b):
        return Band(self.arr[b - 1])

    def ReadAsArray(self, x0=0, y0=0, xoff=None, yoff=None):
        return self.arr[0][y0:y0 + yoff, x0:x0 + xoff].copy()


def load_reference_module():
    for name in ("osgeo", "osgeo.gdal", "isce", "isceobj", "Network"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["osgeo"].gdal = sys.modules["osgeo.gdal"]
    sys.modules["Network"].Network = object
    sys.path.insert(0, REF_PY)
    import integratePS
    return integratePS


def main():
    ips = load_reference_module()
    rng = np.random.default_rng(77)
    lines, cols, dates = 24, 40, 4
    slc = ((rng.standard_normal((dates, lines, cols)) + 1j * rng.standard_normal((dates, lines, cols))) * 3).astype(np.complex64)
    slc[1, 3, 5] = 0                                     # a zero sample: angle(0) = 0
    slc[2, 7, 9] = complex(-2.5, 0.0)                    # with slc[0] real positive below: product on the negative real axis
    slc[0, 7, 9] = complex(1.5, 0.0)
    ds = np.exp(1j * rng.uniform(-np.pi, np.pi, (dates, lines, cols))).astype(np.complex64)
    ps = (rng.random((lines, cols)) > 0.75).astype(np.uint8)
    ps[3, 5] = ps[7, 9] = 1
    tcorr = rng.random((lines, cols)).astype(np.float32)
    out = {}
    for j in range(1, dates):
        res = Dataset(np.zeros((lines, cols), np.complex64))
        ips.integratePS2DS(Dataset(ds[0]), Dataset(ds[j]), Dataset(slc), Dataset(tcorr), Dataset(ps), res,
                           nblocks=3, linesPerBlock=10, band_i=1, band_j=j + 1)
        out[f"ifg_0_{j}"] = res.arr[0]
    coh = Dataset(np.zeros((lines, cols), np.float32))
    ips.getCoherence(Dataset(tcorr), Dataset(ps), coh, cols, lines, 3, 10)
    np.savez_compressed(os.path.join(HERE, "integrate_ps_24x40.npz"), slc=slc, ds=ds, ps=ps, tcorr=tcorr, coherence=coh.arr[0], **out)
    print("wrote integrate_ps_24x40.npz:", {k: v.dtype for k, v in out.items()})


if __name__ == "__main__":
    main()
