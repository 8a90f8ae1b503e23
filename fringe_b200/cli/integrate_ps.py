#!/usr/bin/env python3
"""PS / DS integration from the command line: options and outputs of python/integratePS.py (single-reference network,
:229-262): for every pair first date - date_j the wrapped interferogram <out>/<date_0>_<date_j>.int is the product of
the adjusted DS phasors except at PS pixels, where it is the unit-modulus full-resolution interferogram; the coherence
raster <out>/tcorr_ds_ps.bin gets 0.95 at PS pixels; with -u a run_unwrap_ps_ds.sh script lists the unwrapping commands.
The per-pixel work runs on the GPU (fringe_integrate_ps / fringe_ps_coherence); rasters are read without GDAL."""
import argparse
import os

import numpy as np

from .. import stackio


def cmdLineParser(argv=None):
    parser = argparse.ArgumentParser(description='integrate PS pixels into existing unwrapped DS filed',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-s', '--slc_stack', type=str, dest='slcStack', required=True, help='slc stack dataset ')
    parser.add_argument('-d', '--ds_stack_dir', type=str, dest='dsStackDir', required=True,
                        help='The directory that contains adjusted stack based on DS analysis')
    parser.add_argument('-t', '--tcorr_file', type=str, dest='tcorrFile', required=True,
                        help='A temporal coheremce file which represents the coherence of the stack')
    parser.add_argument('-p', '--psPixels_file', type=str, dest='psPixelsFile', required=True, help='The map of PS pixles')
    parser.add_argument('-o', '--output_dir', type=str, dest='outDir', required=True,
                        help='The output directory to store the wrapped phase series of DS and PS pixels')
    parser.add_argument('-c', '--coreg_slc_dir', type=str, dest='coregSlcDir', required=False,
                        help='The directory that contains the coregistered stack of SLCs')
    parser.add_argument('-u', '--unw_method', '--unwrap_method', type=str, dest='unwrapMethod', choices=('snaphu', 'phass'),
                        help='phase unwrapping method; writes run_unwrap_ps_ds.sh with the unwrap commands')
    parser.add_argument('-x', '--xml_file', type=str, dest='xmlFile', required=False,
                        help='path of reference xml file for unwrapping with snaphu')
    return parser.parse_args(argv)


def main(argv=None):
    inps = cmdLineParser(argv)
    from ..engine import Context
    bands = stackio.read_stack_vrt(inps.slcStack)           # [(date, reader)] in band order
    dates = sorted(d for d, _ in bands)
    readers = dict(bands)
    tcorr = stackio.read_envi(inps.tcorrFile).astype(np.float32)
    ps = stackio.read_envi(inps.psPixelsFile)
    ps = (ps == 1).astype(np.uint8)
    os.makedirs(inps.outDir, exist_ok=True)
    lines, cols = tcorr.shape
    print("number of SLC: ", len(dates)); print("number of rows: ", lines); print("number of columns: ", cols)
    pairs = [(dates[0], d) for d in dates[1:]]              # Network.single_master
    cor_name = os.path.join(inps.outDir, "tcorr_ds_ps.bin")
    with Context(0) as ctx:
        stackio.write_envi(cor_name, ctx.ps_coherence(tcorr, ps, 0.95))
        slc_i = readers[dates[0]]()
        ds_i = stackio.read_envi(os.path.join(inps.dsStackDir, dates[0] + ".slc"))
        for date_i, date_j in pairs:
            print(date_i + "-" + date_j)
            ds_j = stackio.read_envi(os.path.join(inps.dsStackDir, date_j + ".slc"))
            out = ctx.integrate_ps(ds_i, ds_j, slc_i, readers[date_j](), ps)
            stackio.write_envi(os.path.join(inps.outDir, "{0}_{1}.int".format(date_i, date_j)), out)
    if inps.unwrapMethod is not None:
        unw_dir = os.path.join(inps.outDir, "unwrap")
        os.makedirs(unw_dir, exist_ok=True)
        with open("run_unwrap_ps_ds.sh", "w") as runf:
            runf.write("set -e\n")
            for date_i, date_j in pairs:
                int_name = os.path.join(inps.outDir, "{0}_{1}.int".format(date_i, date_j))
                unw_name = os.path.join(unw_dir, "{0}_{1}.unw".format(date_i, date_j))
                cmd = "unwrap_fringe.py -m " + inps.unwrapMethod + " -i " + int_name + " -c " + cor_name + " -o " + unw_name
                if inps.xmlFile is not None:
                    cmd += " -x " + inps.xmlFile
                runf.write(cmd + "\n")
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
