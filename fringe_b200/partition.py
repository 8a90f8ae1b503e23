"""Row partition of an image across GPUs.

Every output pixel depends only on inputs within +-Ny lines (src/nmap/nmap.cpp:404-434,
src/evd/evd.cpp:530-547), so contiguous row tiles with Ny-line halos taken from the input are
independent: no collective is needed.  The same arithmetic is what the reference's block loop
does in time (firstlinetowrite / linestowrite, nmap.cpp:487-516); here it is done in space.
"""
from __future__ import annotations

from typing import NamedTuple


class RowTile(NamedTuple):
    r0: int          # first output line owned by this rank (image coordinates)
    r1: int          # one past the last owned line
    b0: int          # first line read (tile + halo, clipped to the image)
    b1: int          # one past the last line read
    first_line: int  # r0 relative to the block that is read  (the C ABI's first_line)
    n_lines: int     # r1 - r0                                (the C ABI's n_lines)


def row_tile(lines: int, rank: int, world: int, halo: int) -> RowTile:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank outside world")
    r0 = (lines * rank) // world
    r1 = (lines * (rank + 1)) // world
    b0, b1 = max(0, r0 - halo), min(lines, r1 + halo)
    return RowTile(r0, r1, b0, b1, r0 - b0, r1 - r0)


def block_schedule(rows: int, blockysize: int, halo: int):
    """The reference's streaming block schedule (src/nmap/nmap.cpp:302-573, src/evd/evd.cpp:399-872):
    yields (yoff, inysize, firstlinetowrite, linestowrite) for each block read."""
    yoff, blockcount = 0, 0
    while yoff < rows:
        blockcount += 1
        inysize = min(blockysize, rows - yoff)
        last = (yoff + blockysize) >= rows
        if blockcount == 1:
            first, nwrite, rollback = 0, inysize - halo, halo
            if last:
                nwrite, rollback = inysize, 0
        elif last:
            first, nwrite, rollback = halo, inysize - halo, 0
        else:
            first, nwrite, rollback = halo, inysize - 2 * halo, 0
        yield yoff, inysize, first, nwrite
        yoff = rows if last else yoff + nwrite - rollback
