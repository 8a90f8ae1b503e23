import os
import sys

BINDINGS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings")


def use_bindings():
    """Make `import nmaplib / evdlib / phase_linklib` resolve to the in-tree extension modules."""
    if BINDINGS not in sys.path:
        sys.path.insert(0, BINDINGS)
