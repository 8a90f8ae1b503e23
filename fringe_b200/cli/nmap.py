#!/usr/bin/env python3
"""SHP selection from the command line: options of src/nmap/nmap.py:6-36, work done by nmaplib.Nmap."""
from ._common import BLOCK_LINES, REQUIRED, WINDOW_X, WINDOW_Y, build_parser, configure, ensure_parent, ram, use_bindings

OPTIONS = [
    ('-i', '--input', 'inputDS', str, REQUIRED, 'stack VRT, one band per acquisition'),
    ('-o', '--output', 'outputDS', str, REQUIRED, 'neighbourhood bit mask to write (UInt32 BIP)'),
    ('-c', '--count', 'countDS', str, REQUIRED, 'neighbour count raster to write'),
    ('-m', '--mask', 'maskDS', str, '', 'byte raster; pixels that are 0 there are skipped'),
    BLOCK_LINES, ram(256), WINDOW_X, WINDOW_Y,
    ('-p', '--prob', 'pValue', float, 0.05, 'two pixels are neighbours when the test p-value is at least this'),
    ('-s', '--stat', 'method', str, 'KS2', 'two-sample test: KS2 or AD2'),
    (None, '--nogpu', 'noGPU', 'flag', None, 'kept for compatibility; there is no CPU path'),
]
WIRING = {'inputDS': 'inputDS', 'weightsDS': 'outputDS', 'countDS': 'countDS', 'blocksize': 'linesPerBlock',
          'memsize': 'memorySize', 'halfWindowX': 'halfWindowX', 'halfWindowY': 'halfWindowY',
          'minimumProbability': 'pValue', 'method': lambda a: a.method.upper(), 'noGPU': 'noGPU'}


def cmdLineParser(argv=None):
    return build_parser('Neighbourhood mask and neighbour count of a coregistered SLC stack', OPTIONS).parse_args(argv)


def main(argv=None):
    inps = cmdLineParser(argv)
    ensure_parent(inps.outputDS)
    use_bindings()
    import nmaplib
    job = configure(nmaplib.Nmap(), inps, WIRING)
    if inps.maskDS:
        job.maskDS = inps.maskDS
    job.run()


if __name__ == '__main__':
    main()
