#!/usr/bin/env python3
"""Sequential (ministack) estimator from the command line.

Behaviour and on-disk layout of src/sequential/sequential.py:144-254 with the stack bookkeeping of
src/sequential/Stack.py, minus GDAL (sizes come from the .vrt / .hdr files, the compressed-SLC VRTs are
written directly):

    <out>/fullStack/{slcs,stack}/                    VRTs of the input SLCs
    <out>/miniStacks/<first>_<last>/{slcs,stack,EVD}/ one folder per ministack; EVD/<date>.slc, tcorr.bin
    <out>/compressedSlc/<last>/<last>.slc(+.vrt)     compressed SLC of each ministack
    <out>/Datum_connection/{slcs,stack,EVD}/         EVD over all compressed SLCs

    <out>/adjusted/<date>.slc                         ministack phasor x datum phasor (python/adjustMiniStacks.py)

Ministack k links (k-1) compressed SLCs followed by its own acquisitions with miniStackCount = k and the
binding's default method (MLE); the datum connection runs with miniStackCount = 1.  All of it happens on the
device in one call per block of lines (fringe_sequential_block).  The reference also writes, inside every
ministack's EVD folder, phasor files for the compressed-SLC bands it prepended; nothing reads those and they
are not produced here.
"""
import glob
import os
import shutil

from ._common import BLOCK_LINES, REQUIRED, build_parser, ram
from .. import stackio

OPTIONS = [
    ('-i', '--inDir', 'inputDir', str, REQUIRED, 'folder with one sub-folder per acquisition date holding <date>.slc'),
    ('-w', '--weight_dataset', 'weightDS', str, REQUIRED, 'neighbourhood bit mask written by nmap'),
    ('-o', '--outDir', 'outputDir', str, REQUIRED, 'folder for everything this run produces'),
    BLOCK_LINES, ram(2048),
    ('-x', '--xhalf', 'halfWindowX', int, 29, 'half width of the SHP window in range pixels'),
    ('-y', '--yhalf', 'halfWindowY', int, 9, 'half height of the SHP window in azimuth lines'),
    ('-m', '--minneigh', 'minNeighbors', int, 5, 'pixels with fewer neighbours are left empty'),
    ('-s', '--mini_stack_size', 'miniStackSize', int, 10, 'acquisitions per ministack'),
    ('-f', '--force', 'forceprocessing', 'flag', None, 'redo ministacks whose folder already exists'),
]


def cmdLineParser(argv=None):
    parser = build_parser('Phase linking of a long SLC stack in ministacks with compressed-SLC hand-off', OPTIONS)
    parser.add_argument('-b', '--bbox', dest='bbox', nargs='+', type=str, default=None,
                        help='crop: first line, last line, first pixel, last pixel')
    return parser.parse_args(argv)


def find_slcs(folder):
    """Sorted <folder>/<date>/<date>.slc (or .slc.full) files; the date is the sub-folder's name."""
    found = glob.glob(os.path.join(folder, '*/*.slc')) or glob.glob(os.path.join(folder, '*/*.slc.full'))
    print('Number of SLCs discovered: ', len(found))
    return sorted(found)


def date_of(slc):
    return os.path.basename(os.path.dirname(slc))


def _size(slc):
    return stackio.raster_size(slc + '.vrt') if os.path.exists(slc + '.vrt') else stackio.raster_size(slc)


def write_stack(out_dir, slcs, crop, bbox):
    """Per-date raw VRTs under <out_dir>/slcs and the band-per-date <out_dir>/stack/stack.vrt.
    crop[i] says whether `bbox` (y0, y1, x0, x1) applies to slcs[i]: compressed SLCs are already cropped."""
    slc_dir, stack_dir = os.path.join(out_dir, 'slcs'), os.path.join(out_dir, 'stack')
    os.makedirs(slc_dir, exist_ok=True)
    os.makedirs(stack_dir, exist_ok=True)

    def window(i):
        width, height = _size(slcs[i])
        y0, y1, x0, x1 = bbox if (bbox and crop[i]) else (0, height, 0, width)
        return width, height, x0, y0, x1 - x0, y1 - y0

    for slc in slcs:
        width, height = _size(slc)
        stackio.write_raw_vrt(os.path.join(slc_dir, date_of(slc) + '.vrt'), slc, width, height)
    stack_vrt = os.path.join(stack_dir, 'stack.vrt')
    print('writing ', stack_vrt)
    with open(stack_vrt, 'w') as fid:
        *_, xsize, ysize = window(len(slcs) - 1)
        fid.write('<VRTDataset rasterXSize="{0}" rasterYSize="{1}">\n'.format(xsize, ysize))
        for i, slc in enumerate(slcs):
            width, height, x0, y0, xs, ys = window(i)
            date = date_of(slc)
            fid.write(stackio.STACK_BAND.format(width=width, height=height, xmin=x0, ymin=y0, xsize=xs, ysize=ys,
                                                date=date, acq=date, wvl=0.03, index=i + 1, extra='',
                                                path=os.path.abspath(os.path.join(slc_dir, date + '.vrt'))))
        fid.write('</VRTDataset>')
    return stack_vrt


def _open_output(path, lines, cols, dtype):
    """A zero-filled ENVI raster (+ .hdr, + raw .vrt for complex products) opened for in-place writing."""
    import numpy as np
    os.makedirs(os.path.dirname(path), exist_ok=True)
    stackio.write_envi_header(path, lines, cols, 1, dtype)
    arr = np.memmap(path, dtype=dtype, mode="w+", shape=(lines, cols))
    if np.dtype(dtype) == np.complex64:
        stackio.write_raw_vrt(path + ".vrt", path, cols, lines)
    return arr


def main(argv=None):
    """The chain runs on the device: every block of lines is uploaded once and goes through all ministacks, the datum
    connection and the wrapped-phase adjustment in one call (fringe_sequential_block); compressed SLCs never touch the
    disk between the stages.  Files and folders are the reference's (sequential.py:153-254); in addition
    <out>/adjusted/<date>.slc holds the product python/adjustMiniStacks.py builds lazily from them."""
    import threading

    import numpy as np

    from ..engine import Context, device_count, nulong
    from .._lib import lib

    inps = cmdLineParser(argv)
    root = inps.outputDir = os.path.abspath(inps.outputDir)
    bbox = tuple(int(v) for v in inps.bbox) if inps.bbox is not None else None
    if bbox:
        print('input bounding box in (y0, y1, x0, x1): {}'.format(bbox))
    slcs = find_slcs(inps.inputDir)
    if len(slcs) < 2 * inps.miniStackSize and len(slcs) <= inps.miniStackSize:
        raise SystemExit("fewer acquisitions than one ministack: run evd.py instead")
    dates = [date_of(p) for p in slcs]
    width, height = _size(slcs[0])
    y0, y1, x0, x1 = bbox if bbox else (0, height, 0, width)
    lines, cols = y1 - y0, x1 - x0
    s = inps.miniStackSize
    groups = [list(range(a, min(len(slcs), a + s))) for a in range(0, len(slcs), s)]
    nmini = len(groups)
    datum_dir = os.path.join(root, 'Datum_connection')
    if os.path.isdir(os.path.join(datum_dir, 'EVD')) and not inps.forceprocessing:
        print('{0} looks like it has already been processed. Skipping ... '.format(datum_dir))
        return 0
    if inps.forceprocessing:
        for sub in ('miniStacks', 'compressedSlc', 'Datum_connection', 'adjusted'):
            shutil.rmtree(os.path.join(root, sub), ignore_errors=True)
    comp_root = os.path.join(root, 'compressedSlc')
    os.makedirs(comp_root, exist_ok=True)
    write_stack(os.path.join(root, 'fullStack'), slcs, [True] * len(slcs), bbox)

    # the weights raster and its window must match the request (evd.cpp:94-166: codes 105-110)
    hdr = stackio.read_envi_header(inps.weightDS)
    nu = nulong(inps.halfWindowX, inps.halfWindowY)
    if int(hdr.get("halfwindowx", 0)) != inps.halfWindowX or int(hdr.get("halfwindowy", 0)) != inps.halfWindowY:
        raise RuntimeError("sequential: half window of the weights raster differs from the request (evd_process would return 109/110)")
    if int(hdr["samples"]) != cols or int(hdr["lines"]) != lines or int(hdr.get("bands", 1)) != nu:
        raise RuntimeError("sequential: weights raster does not match the stack (evd_process would return 106-108)")
    wts = np.memmap(inps.weightDS, dtype=np.uint32, mode="r", shape=(lines, cols, nu))
    inputs = [np.memmap(p, dtype=np.complex64, mode="r", shape=(height, width)) for p in slcs]

    # outputs, laid out as the reference does
    mini_out, mini_tc, comp_out, datum_out = [], [], [], []
    for k, grp in enumerate(groups):
        folder = os.path.join(root, 'miniStacks', dates[grp[0]] + '_' + dates[grp[-1]])
        earlier = [os.path.join(comp_root, dates[g[-1]], dates[g[-1]] + '.slc') for g in groups[:k]]
        evd_dir = os.path.join(folder, 'EVD')
        for d in grp:
            mini_out.append(_open_output(os.path.join(evd_dir, dates[d] + '.slc'), lines, cols, np.complex64))
        mini_tc.append(_open_output(os.path.join(evd_dir, 'tcorr.bin'), lines, cols, np.float32))
        last = dates[grp[-1]]
        comp_out.append(_open_output(os.path.join(comp_root, last, last + '.slc'), lines, cols, np.complex64))
        write_stack(folder, earlier + [slcs[d] for d in grp], [False] * len(earlier) + [True] * len(grp), bbox)
        datum_out.append(_open_output(os.path.join(datum_dir, 'EVD', last + '.slc'), lines, cols, np.complex64))
    datum_tc = _open_output(os.path.join(datum_dir, 'EVD', 'tcorr.bin'), lines, cols, np.float32)
    adjusted = [_open_output(os.path.join(root, 'adjusted', d + '.slc'), lines, cols, np.complex64) for d in dates]
    write_stack(datum_dir, [os.path.join(comp_root, dates[g[-1]], dates[g[-1]] + '.slc') for g in groups], [False] * nmini, None)

    # block schedule: the memory budget covers input + outputs of a block (about 3 stacks); every block carries the
    # halo the chain needs so that blocks are independent
    halo = lib.fringe_sequential_halo(len(slcs), s, inps.halfWindowY)
    per_line = cols * 8 * (3 * len(slcs) + 3 * nmini) + cols * 4 * (nu + nmini + 1)
    rows_per_block = max(inps.linesPerBlock, int(inps.memorySize * 1.0e6 / per_line) // inps.linesPerBlock * inps.linesPerBlock)
    blocks = [(a, min(lines, a + rows_per_block)) for a in range(0, lines, rows_per_block)]
    print('Number of ministacks: {0}, halo {1} lines, {2} block(s) of up to {3} lines'.format(nmini, halo, len(blocks), rows_per_block))
    ngpu = max(1, min(device_count(), len(blocks)))
    lock, errors = threading.Lock(), []
    todo = list(blocks)

    def worker(dev):
        try:
            with Context(dev) as ctx:
                while True:
                    with lock:
                        if not todo or errors:
                            return
                        a, b = todo.pop(0)
                    b0, b1 = max(0, a - halo), min(lines, b + halo)
                    stack = np.stack([m[y0 + b0:y0 + b1, x0:x1] for m in inputs])
                    res = ctx.sequential_block(stack, np.ascontiguousarray(wts[b0:b1]), inps.halfWindowX, inps.halfWindowY, s,
                                               first_line=a - b0, n_lines=b - a)
                    rows = slice(a - b0, b - b0)
                    for d in range(len(slcs)):
                        mini_out[d][a:b] = res["out_mini"][d][rows]
                        adjusted[d][a:b] = res["adjusted"][d][rows]
                    for k in range(nmini):
                        mini_tc[k][a:b] = res["tcorr_mini"][k][rows]
                        comp_out[k][a:b] = res["comp"][k][rows]
                        datum_out[k][a:b] = res["out_datum"][k][rows]
                    datum_tc[a:b] = res["tcorr_datum"][rows]
        except Exception as exc:                        # noqa: BLE001  (reported by the main thread)
            with lock:
                errors.append(exc)

    threads = [threading.Thread(target=worker, args=(d,)) for d in range(ngpu)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for arr in mini_out + mini_tc + comp_out + datum_out + adjusted + [datum_tc]:
        arr.flush()
    return 0


if __name__ == '__main__':
    main()
