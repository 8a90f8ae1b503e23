#!/usr/bin/env python3
"""Amplitude calibration constants from the command line: options of src/calamp/calamp.py:6-30, work done by
calamplib.Calamp (mean amplitude of the valid pixels of every band, written as amplitudeConstant into the slc metadata
of a copy of the stack VRT -- what nmap and ampdispersion read back)."""
from ._common import REQUIRED, build_parser, configure, use_bindings

OPTIONS = [
    ('-i', '--input', 'inputDS', str, REQUIRED, 'stack VRT, one band per acquisition'),
    ('-o', '--output', 'outputDS', str, REQUIRED, 'stack VRT to write, with the calibration constants in its metadata'),
    ('-m', '--mask', 'maskDS', str, '', 'mask raster: pixels with 0 (water, layover ...) stay out of the mean'),
    ('-d', '--default', 'defaultValue', float, 1.0, 'constant for a band without any valid pixel'),
    ('-l', '--linesperblock', 'linesPerBlock', int, 64, 'block height is a multiple of this many lines'),
    ('-r', '--ram', 'memorySize', int, 256, 'host memory budget for block buffers, MB'),
    ('-s', '--sqrt', 'sqrt', 'flag', None, 'apply a square root to the amplitudes (carried for compatibility)'),
]
WIRING = {'inputDS': 'inputDS', 'outputDS': 'outputDS', 'defaultValue': 'defaultValue', 'blocksize': 'linesPerBlock',
          'memsize': 'memorySize', 'applySqrt': 'sqrt'}


def cmdLineParser(argv=None):
    return build_parser('Amplitude calibration constants of a coregistered SLC stack', OPTIONS).parse_args(argv)


def main(argv=None):
    inps = cmdLineParser(argv)
    use_bindings()
    import calamplib
    job = configure(calamplib.Calamp(), inps, WIRING)
    if inps.maskDS:
        job.maskDS = inps.maskDS
    job.run()


if __name__ == '__main__':
    main()
