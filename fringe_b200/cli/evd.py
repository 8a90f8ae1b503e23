#!/usr/bin/env python3
"""evd.py / phase_link.py -- same command lines as src/evd/evd.py:6-34 and
src/phase_link/phase_link.py (identical options; the module loaded differs)."""
import argparse
import glob
import os

from ._common import use_bindings
from .. import stackio


def cmdLineParser(argv=None):
    parser = argparse.ArgumentParser(description='Perform MLE-based phase-linking on a stack of coregistered SLCs',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-i', '--input', type=str, dest='inputDS', required=True, help='Input GDAL SLC stack VRT')
    parser.add_argument('-w', '--wts', type=str, dest='wtsDS', required=True, help='Input weights dataset')
    parser.add_argument('-o', '--output', type=str, dest='outputFolder', required=True, help='Output folder with phase-linked SLCs')
    parser.add_argument('-l', '--linesperblock', type=int, dest='linesPerBlock', default=64, help='Quantum for block of lines')
    parser.add_argument('-r', '--ram', type=int, dest='memorySize', default=2048, help='Memory in Mb to use')
    parser.add_argument('-x', '--xhalf', type=int, dest='halfWindowX', default=5, help='Half window size (range)')
    parser.add_argument('-y', '--yhalf', type=int, dest='halfWindowY', default=5, help='Half window size (azimuth)')
    parser.add_argument('-n', '--minneigh', type=int, dest='minNeighbors', default=5, help='Minimum number of neighbors for computation')
    parser.add_argument('-m', '--method', type=str, dest='method', default='MLE', help='Decomposition method to use - MLE / EVD / STBAS')
    parser.add_argument('-b', '--bandwidth', type=int, dest='bandWidth', default=-1, help='Diagonal bandwidth for STBAS')
    return parser.parse_args(argv)


def runEvd(inps, phase_link=False):
    os.environ['OPENBLAS_NUM_THREADS'] = "1"      # kept from src/evd/evd.py:43 (harmless here)
    use_bindings()
    if phase_link:
        import phase_linklib
        aa = phase_linklib.Phaselink()
    else:
        import evdlib
        aa = evdlib.Evd()
    aa.inputDS = inps.inputDS
    aa.weightsDS = inps.wtsDS
    aa.outputFolder = inps.outputFolder
    aa.outputCompressedSlcFolder = aa.outputFolder
    aa.compSlc = "compslc.bin"
    aa.blocksize = inps.linesPerBlock
    aa.memsize = inps.memorySize
    aa.halfWindowX = inps.halfWindowX
    aa.halfWindowY = inps.halfWindowY
    aa.minimumNeighbors = inps.minNeighbors
    aa.method = inps.method
    aa.bandWidth = inps.bandWidth
    aa.run()


def create_vrts(slc_dir):
    """<date>.slc -> <date>.slc.vrt (the reference calls gdal.Translate(format='VRT'), evd.py:69-76)."""
    for f in glob.glob(os.path.join(slc_dir, "*.slc")):
        w, h = stackio.raster_size(f)
        stackio.write_raw_vrt(f + ".vrt", f, w, h)


def main(argv=None, phase_link=False):
    inps = cmdLineParser(argv)
    runEvd(inps, phase_link=phase_link)
    create_vrts(inps.outputFolder)


if __name__ == '__main__':
    main()
