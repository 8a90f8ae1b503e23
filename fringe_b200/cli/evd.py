#!/usr/bin/env python3
"""Phase linking from the command line: options of src/evd/evd.py:6-34 (phase_link.py has the same set),
work done by evdlib.Evd or phase_linklib.Phaselink."""
import glob
import os

from ._common import BLOCK_LINES, REQUIRED, WINDOW_X, WINDOW_Y, build_parser, configure, ram, use_bindings
from .. import stackio

OPTIONS = [
    ('-i', '--input', 'inputDS', str, REQUIRED, 'stack VRT, one band per acquisition'),
    ('-w', '--wts', 'wtsDS', str, REQUIRED, 'neighbourhood bit mask written by nmap'),
    ('-o', '--output', 'outputFolder', str, REQUIRED, 'folder for the linked phasors (must not exist yet)'),
    BLOCK_LINES, ram(2048), WINDOW_X, WINDOW_Y,
    ('-n', '--minneigh', 'minNeighbors', int, 5, 'pixels with fewer neighbours are left empty (phase_link only)'),
    ('-m', '--method', 'method', str, 'MLE', 'estimator: MLE, EVD or STBAS'),
    ('-b', '--bandwidth', 'bandWidth', int, -1, 'number of off-diagonals kept by STBAS'),
]
WIRING = {'inputDS': 'inputDS', 'weightsDS': 'wtsDS', 'outputFolder': 'outputFolder',
          'outputCompressedSlcFolder': 'outputFolder', 'compSlc': lambda a: 'compslc.bin',
          'blocksize': 'linesPerBlock', 'memsize': 'memorySize', 'halfWindowX': 'halfWindowX',
          'halfWindowY': 'halfWindowY', 'minimumNeighbors': 'minNeighbors', 'method': 'method', 'bandWidth': 'bandWidth'}


def cmdLineParser(argv=None):
    return build_parser('Phase linking of a coregistered SLC stack over its statistically homogeneous neighbours',
                        OPTIONS).parse_args(argv)


def create_vrts(slc_dir):
    """A raw VRT next to every <date>.slc of the output folder (the reference shells out to GDAL for this)."""
    for path in glob.glob(os.path.join(slc_dir, "*.slc")):
        width, height = stackio.raster_size(path)
        stackio.write_raw_vrt(path + ".vrt", path, width, height)


def main(argv=None, phase_link=False):
    inps = cmdLineParser(argv)
    use_bindings()
    if phase_link:
        import phase_linklib
        job = phase_linklib.Phaselink()
    else:
        import evdlib
        job = evdlib.Evd()
    configure(job, inps, WIRING).run()
    create_vrts(inps.outputFolder)


if __name__ == '__main__':
    main()
