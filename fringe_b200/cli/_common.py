"""Shared plumbing of the command lines.

Every tool is described by a table of its options -- flag letters, long names, destination and default
are the reference script's (they are the interface); the texts are ours -- plus a mapping from parsed
option to attribute of the binding class.  `build_parser` and `configure` turn the two tables into an
argparse parser and a configured binding object."""
import argparse
import os
import sys

BINDINGS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings")


def use_bindings():
    """Make `import nmaplib / evdlib / phase_linklib / ...` resolve to the in-tree extension modules."""
    if BINDINGS not in sys.path:
        sys.path.insert(0, BINDINGS)


# options several tools share: (short, long, dest, type, default, text)
BLOCK_LINES = ('-l', '--linesperblock', 'linesPerBlock', int, 64, 'block height is a multiple of this many lines')
WINDOW_X = ('-x', '--xhalf', 'halfWindowX', int, 5, 'half width of the SHP window in range pixels')
WINDOW_Y = ('-y', '--yhalf', 'halfWindowY', int, 5, 'half height of the SHP window in azimuth lines')


def ram(default):
    return ('-r', '--ram', 'memorySize', int, default, 'host memory budget for block buffers, MB')


def build_parser(summary, table):
    """table rows: (short, long, dest, type, default, text); default REQUIRED marks a mandatory option,
    type 'flag' a boolean switch, type 'ints' a list of integers."""
    parser = argparse.ArgumentParser(description=summary, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for short, long_name, dest, kind, default, text in table:
        names = [n for n in (short, long_name) if n]
        if kind == 'flag':
            parser.add_argument(*names, dest=dest, action='store_true', default=False, help=text)
        elif kind == 'ints':
            parser.add_argument(*names, dest=dest, type=int, nargs='*', default=list(default), help=text)
        elif default is REQUIRED:
            parser.add_argument(*names, dest=dest, type=kind, required=True, help=text)
        else:
            parser.add_argument(*names, dest=dest, type=kind, default=default, help=text)
    return parser


class _Required:
    def __repr__(self):
        return 'REQUIRED'


REQUIRED = _Required()


def configure(obj, inps, wiring):
    """wiring: {attribute of the binding object: name of the parsed option, or a callable(inps)}."""
    for attr, src in wiring.items():
        setattr(obj, attr, src(inps) if callable(src) else getattr(inps, src))
    return obj


def ensure_parent(path):
    os.makedirs(os.path.abspath(os.path.dirname(path)), exist_ok=True)
