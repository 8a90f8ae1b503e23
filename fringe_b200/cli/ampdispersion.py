#!/usr/bin/env python3
"""Amplitude dispersion from the command line: options of src/ampdispersion/ampdispersion.py:6-30, work done
by ampdispersionlib.Ampdispersion."""
from ._common import BLOCK_LINES, REQUIRED, build_parser, configure, ensure_parent, ram, use_bindings

OPTIONS = [
    ('-i', '--input', 'inputDS', str, REQUIRED, 'stack VRT, one band per acquisition'),
    ('-o', '--output', 'outputDS', str, REQUIRED, 'amplitude dispersion raster to write'),
    ('-m', '--mean', 'meanampDS', str, '', 'mean amplitude raster to write'),
    BLOCK_LINES, ram(256),
    ('-b', '--band', 'refBand', int, 1, 'band whose calibration constant the others are divided by'),
]
WIRING = {'inputDS': 'inputDS', 'outputDS': 'outputDS', 'meanampDS': 'meanampDS', 'blocksize': 'linesPerBlock',
          'memsize': 'memorySize', 'refband': 'refBand'}


def cmdLineParser(argv=None):
    return build_parser('Mean calibrated amplitude and amplitude dispersion of every pixel of an SLC stack',
                        OPTIONS).parse_args(argv)


def main(argv=None):
    inps = cmdLineParser(argv)
    ensure_parent(inps.outputDS)
    use_bindings()
    import ampdispersionlib
    configure(ampdispersionlib.Ampdispersion(), inps, WIRING).run()


if __name__ == '__main__':
    main()
