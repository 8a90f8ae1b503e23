#!/usr/bin/env python3
"""ampdispersion.py -- same command line as src/ampdispersion/ampdispersion.py:6-30 (defaults included)."""
import argparse
import os

from ._common import use_bindings


def cmdLineParser(argv=None):
    parser = argparse.ArgumentParser(description='Compute amplitude dispersion and mean amplitude of a stack of coregistered SLCs',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-i', '--input', type=str, dest='inputDS', required=True, help='Input GDAL SLC stack VRT')
    parser.add_argument('-o', '--output', type=str, dest='outputDS', required=True, help='Output amplitude dispersion dataset')
    parser.add_argument('-m', '--mean', type=str, dest='meanampDS', default='', help='Output mean amplitude')
    parser.add_argument('-l', '--linesperblock', type=int, dest='linesPerBlock', default=64, help='Quantum for block of lines')
    parser.add_argument('-r', '--ram', type=int, dest='memorySize', default=256, help='Memory in Mb to use')
    parser.add_argument('-b', '--band', type=int, dest='refBand', default=1, help='Reference band to use for relative normalization')
    return parser.parse_args(argv)


def runAmpdispersion(inps):
    use_bindings()
    import ampdispersionlib
    aa = ampdispersionlib.Ampdispersion()
    aa.inputDS = inps.inputDS
    aa.outputDS = inps.outputDS
    aa.meanampDS = inps.meanampDS
    aa.blocksize = inps.linesPerBlock
    aa.memsize = inps.memorySize
    aa.refband = inps.refBand
    aa.run()


def main(argv=None):
    inps = cmdLineParser(argv)
    os.makedirs(os.path.abspath(os.path.dirname(inps.outputDS)), exist_ok=True)
    runAmpdispersion(inps)


if __name__ == '__main__':
    main()
