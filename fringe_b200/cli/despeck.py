#!/usr/bin/env python3
"""SHP-weighted multilooking from the command line: options of src/despeck/despeck.py:6-33, work done by
despecklib.Despeck."""
from ._common import BLOCK_LINES, REQUIRED, WINDOW_X, WINDOW_Y, build_parser, configure, ensure_parent, ram, use_bindings

OPTIONS = [
    ('-i', '--input', 'inputDS', str, REQUIRED, 'stack VRT, one band per acquisition'),
    ('-o', '--output', 'outputDS', str, REQUIRED, 'filtered raster to write'),
    ('-w', '--wts', 'wtsDS', str, REQUIRED, 'neighbourhood bit mask written by nmap'),
    BLOCK_LINES, ram(512), WINDOW_X, WINDOW_Y,
    ('-b', '--band', 'bands', 'ints', (), 'one band (its amplitude) or two (their interferogram), counted from 1'),
    ('-c', '--corr', 'cohFlag', 'flag', None, 'normalise the interferogram to a coherence'),
]
WIRING = {'inputDS': 'inputDS', 'weightsDS': 'wtsDS', 'outputDS': 'outputDS', 'blocksize': 'linesPerBlock',
          'memsize': 'memorySize', 'halfWindowX': 'halfWindowX', 'halfWindowY': 'halfWindowY'}


def cmdLineParser(argv=None):
    return build_parser('Average an amplitude or an interferogram over each pixel\'s homogeneous neighbours',
                        OPTIONS).parse_args(argv)


def main(argv=None):
    inps = cmdLineParser(argv)
    if len(inps.bands) > 2:
        raise Exception('Despeck can handle one or two bands. More than two bands provided')
    if len(inps.bands) == 1 and inps.cohFlag:
        raise Exception('User requested coherence when requesting despeckling of SLC magnitude')
    ensure_parent(inps.outputDS)
    use_bindings()
    import despecklib
    job = configure(despecklib.Despeck(), inps, WIRING)
    if inps.bands:
        job.band1 = inps.bands[0]
    if len(inps.bands) == 2:
        job.band2 = inps.bands[1]
        job.coherenceFlag = inps.cohFlag
    job.run()


if __name__ == '__main__':
    main()
