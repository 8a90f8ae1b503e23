#!/usr/bin/env python3
"""Sequential (ministack) estimator from the command line.

Behaviour and on-disk layout of src/sequential/sequential.py:144-254 with the stack bookkeeping of
src/sequential/Stack.py, minus GDAL (sizes come from the .vrt / .hdr files, the compressed-SLC VRTs are
written directly):

    <out>/fullStack/{slcs,stack}/                    VRTs of the input SLCs
    <out>/miniStacks/<first>_<last>/{slcs,stack,EVD}/ one folder per ministack; EVD/<date>.slc, tcorr.bin
    <out>/compressedSlc/<last>/<last>.slc(+.vrt)     compressed SLC of each ministack
    <out>/Datum_connection/{slcs,stack,EVD}/         EVD over all compressed SLCs

Ministack k links (k-1) compressed SLCs followed by its own acquisitions with miniStackCount = k and the
binding's default method (MLE); the datum connection runs with miniStackCount = 1.
"""
import glob
import os
import shutil

from ._common import BLOCK_LINES, REQUIRED, build_parser, ram, use_bindings
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


def link(inps, stack_vrt, out_dir, first_real_band, comp_dir=None, comp_name=None):
    """One evd run; the estimator stays at the binding's default (MLE), as in the reference."""
    use_bindings()
    import evdlib
    job = evdlib.Evd()
    job.inputDS, job.weightsDS, job.outputFolder = stack_vrt, inps.weightDS, out_dir
    job.miniStackCount = first_real_band
    job.blocksize, job.memsize = inps.linesPerBlock, inps.memorySize
    job.halfWindowX, job.halfWindowY = inps.halfWindowX, inps.halfWindowY
    job.minimumNeighbors = inps.minNeighbors
    job.outputCompressedSlcFolder = comp_dir or out_dir
    job.compSlc = comp_name or 'compslc.bin'
    job.run()


def main(argv=None):
    inps = cmdLineParser(argv)
    root = inps.outputDir = os.path.abspath(inps.outputDir)
    bbox = tuple(int(v) for v in inps.bbox) if inps.bbox is not None else None
    if bbox:
        print('input bounding box in (y0, y1, x0, x1): {}'.format(bbox))
    comp_root = os.path.join(root, 'compressedSlc')
    os.makedirs(comp_root, exist_ok=True)

    slcs = find_slcs(inps.inputDir)
    write_stack(os.path.join(root, 'fullStack'), slcs, [True] * len(slcs), bbox)

    for k, start in enumerate(range(0, len(slcs), inps.miniStackSize), start=1):
        own = slcs[start:start + inps.miniStackSize]
        folder = os.path.join(root, 'miniStacks', date_of(own[0]) + '_' + date_of(own[-1]))
        if os.path.isdir(folder) and not inps.forceprocessing:
            print('{0} looks like it has already been processed. Skipping ... '.format(folder))
            continue
        print('Processing {0}'.format(folder))
        # compressed SLCs of the earlier ministacks first (sorted; the reference's glob order is arbitrary
        # and the result does not depend on it), uncropped because they already are
        earlier = sorted(glob.glob(os.path.join(comp_root, '*/*.slc')))
        members = earlier + own
        stack_vrt = write_stack(folder, members, [False] * len(earlier) + [True] * len(own), bbox)
        evd_dir = os.path.join(folder, 'EVD')
        if inps.forceprocessing and os.path.isdir(evd_dir):
            shutil.rmtree(evd_dir)
        last = date_of(own[-1])
        comp_dir = os.path.join(comp_root, last)
        os.makedirs(comp_dir, exist_ok=True)
        link(inps, stack_vrt, evd_dir, k, comp_dir, last + '.slc')
        comp = os.path.join(comp_dir, last + '.slc')
        width, height = stackio.raster_size(comp)
        stackio.write_raw_vrt(comp + '.vrt', comp, width, height)

    datum = os.path.join(root, 'Datum_connection')
    comps = find_slcs(comp_root)
    stack_vrt = write_stack(datum, comps, [False] * len(comps), None)
    link(inps, stack_vrt, os.path.join(datum, 'EVD'), 1)


if __name__ == '__main__':
    main()
