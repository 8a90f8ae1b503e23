#!/usr/bin/env python3
"""adjustMiniStacks.py (wrapped mode) -- python/adjustMiniStacks.py:66-101,180-231 of the reference.

For every acquisition date the wrapped time-series phasor is
    <outDir>/<date>.slc = <miniStackDir>/<first>_<last>/EVD/<date>.slc  *  <datumDir>/EVD/<last>.slc
where <first>_<last> is the ministack the date belongs to.  The reference writes one VRT per date
whose GDAL "mul" pixel function evaluates the product lazily; here the product is computed on the GPU
(fringe_cmul) and materialised as an ENVI raster plus a raw VRT, so downstream tools read plain files.
Same options; --unwrapped (which only rearranges unwrapped-phase VRTs) is outside the hot path.
"""
import argparse
import glob
import os

import numpy as np

from .. import stackio


def cmdLineParser(argv=None):
    parser = argparse.ArgumentParser(description='Adjusts mini-stack wrapped phase series with the datum adjustment phase.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-s', '--slcDir', type=str, dest='slcDir', required=True, help='Input folder which contains vrt files of all slcs')
    parser.add_argument('-m', '--miniStackDir', type=str, dest='miniStackDir', required=True, help='Input mini stack directory')
    parser.add_argument('-d', '--datumDir', type=str, dest='datumDir', required=True, help='Input datum connection directory')
    parser.add_argument('-M', '--miniStackSize', type=int, dest='miniStackSize', required=True, help='size of each miniStack')
    parser.add_argument('-o', '--outDir', type=str, dest='outDir', required=True, help='output directory')
    parser.add_argument('--unwrapped', action='store_true', default=False,
                        help='not supported here: unwrapped adjustment only rewires VRTs of unwrapped phases')
    return parser.parse_args(argv)


def getDates(slcDir):
    """Dates = basenames of <slcDir>/*.vrt, sorted (adjustMiniStacks.py:52-63)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(slcDir, "*.vrt")))


def getStackDict(dateList, miniStackDir, datumDir, outDir, miniStackSize, subDir="EVD", fileExtension=".slc",
                 outputExtension=".slc"):
    """date -> [ministack raster, datum raster, temporal coherence, output] (adjustMiniStacks.py:66-101)."""
    miniStackDir, datumDir, outDir = (os.path.abspath(p) for p in (miniStackDir, datumDir, outDir))
    stackDict = {}
    for indStart in range(0, len(dateList), miniStackSize):
        dates = dateList[indStart:indStart + miniStackSize]
        miniStackPath = os.path.join(miniStackDir, dates[0] + "_" + dates[-1])
        datumPath = os.path.join(datumDir, subDir, dates[-1] + fileExtension)
        for dd in dates:
            stackDict[dd] = [os.path.join(miniStackPath, subDir, dd + fileExtension), datumPath,
                             os.path.join(miniStackPath, "EVD/tcorr.bin"), os.path.join(outDir, dd + outputExtension)]
    return stackDict


def adjust_wrapped(dateList, inps, ctx):
    stackDict = getStackDict(dateList, inps.miniStackDir, inps.datumDir, inps.outDir, inps.miniStackSize)
    width, length = stackio.raster_size(stackDict[dateList[0]][0])
    print("length, width: {0}, {1}".format(length, width))
    datum_cache = (None, None)
    for k in dateList:
        miniStackSlc, adjustSlc, _, output = stackDict[k]
        print("mini stack slc: ", miniStackSlc)
        print("datum compensation slc: ", adjustSlc)
        print("adjusted output phase: ", output)
        if datum_cache[0] != adjustSlc:
            datum_cache = (adjustSlc, stackio.read_envi(adjustSlc))
        out = ctx.cmul(stackio.read_envi(miniStackSlc), datum_cache[1])
        stackio.write_envi(output, out)
        stackio.write_raw_vrt(output + ".vrt", output, width, length)


def main(argv=None):
    inps = cmdLineParser(argv)
    if inps.unwrapped:
        raise SystemExit("--unwrapped is not supported: it only rearranges VRTs of unwrapped phases")
    os.makedirs(inps.outDir, exist_ok=True)
    dateList = getDates(inps.slcDir)
    if not dateList:
        raise SystemExit("no *.vrt files in " + inps.slcDir)
    from ..engine import Context
    with Context(0) as ctx:
        adjust_wrapped(dateList, inps, ctx)
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
