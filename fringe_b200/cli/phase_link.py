#!/usr/bin/env python3
"""phase_link.py -- src/phase_link/phase_link.py: same options as evd.py, phase_linklib.Phaselink."""
from .evd import main as _main


def main(argv=None):
    _main(argv, phase_link=True)


if __name__ == '__main__':
    main()
