#!/usr/bin/env python3
"""sequential.py -- the ministack estimator of src/sequential/sequential.py:144-254 with the Stack /
MiniStack helpers of src/sequential/Stack.py, minus GDAL: sizes come from the .vrt / .hdr files and
the compressed-SLC VRTs are written directly instead of shelling out to gdal_translate.

Layout produced (identical): <out>/fullStack/, <out>/miniStacks/<start>_<end>/EVD/{<Date>.slc,tcorr.bin},
<out>/compressedSlc/<lastdate>/<lastdate>.slc(+.vrt), <out>/Datum_connection/EVD/.
"""
import argparse
import glob
import os

from ._common import use_bindings
from .. import stackio


def cmdLineParser(argv=None):
    parser = argparse.ArgumentParser(description='Perform MLE-based phase-linking on a stack of coregistered SLCs',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-i', '--inDir', type=str, dest='inputDir', required=True, help='Input folder which contains folders for each SLC')
    parser.add_argument('-w', '--weight_dataset', type=str, dest='weightDS', required=True, help='Input weights dataset')
    parser.add_argument('-o', '--outDir', type=str, dest='outputDir', required=True, help='Output folder')
    parser.add_argument('-l', '--linesperblock', type=int, dest='linesPerBlock', default=64, help='Quantum for block of lines')
    parser.add_argument('-r', '--ram', type=int, dest='memorySize', default=2048, help='Memory in Mb to use')
    parser.add_argument('-x', '--xhalf', type=int, dest='halfWindowX', default=29, help='Half window size (range)')
    parser.add_argument('-y', '--yhalf', type=int, dest='halfWindowY', default=9, help='Half window size (azimuth)')
    parser.add_argument('-m', '--minneigh', type=int, dest='minNeighbors', default=5, help='Minimum number of neighbors for computation')
    parser.add_argument('-b', '--bbox', dest='bbox', nargs='+', type=str, default=None,
                        help='bounding box : minLine maxLine minPixel maxPixel')
    parser.add_argument('-s', '--mini_stack_size', type=int, dest='miniStackSize', default=10, help='mini stack size')
    parser.add_argument('-f', '--force', dest='forceprocessing', action='store_true', default=False, help='Force reprocessing')
    return parser.parse_args(argv)


class Stack(object):
    """src/sequential/Stack.py:36-175"""

    def __init__(self, slcDir=None):
        self.slcDir = slcDir
        self.bbox = None

    def configure(self, outDir):
        self.outSlcVrtDir = os.path.join(outDir, "slcs")
        self.outStackVrtDir = os.path.join(outDir, "stack")
        os.makedirs(self.outSlcVrtDir, exist_ok=True)
        os.makedirs(self.outStackVrtDir, exist_ok=True)

    def gatherSLCs(self):
        self.slcList = glob.glob(os.path.join(self.slcDir, '*/*.slc'))
        if len(self.slcList) == 0:
            self.slcList = glob.glob(os.path.join(self.slcDir, '*/*.slc.full'))
        print('Number of SLCs discovered: ', len(self.slcList))
        self.slcList.sort()
        self.size = len(self.slcList)
        self.applyBbox = [True] * self.size

    def getSize(self):
        self.size = len(self.slcList)

    def getDates(self):
        self.dateList = [os.path.basename(os.path.dirname(slc)) for slc in self.slcList]

    @staticmethod
    def _size(slc):
        return stackio.raster_size(slc + '.vrt') if os.path.exists(slc + '.vrt') else stackio.raster_size(slc)

    def get_x_y_offsets(self, ind):
        width, height = self._size(self.slcList[ind])
        ymin, ymax, xmin, xmax = 0, height, 0, width
        if self.bbox and self.applyBbox[ind]:
            ymin, ymax, xmin, xmax = self.bbox
        return width, height, xmin, ymin, xmax - xmin, ymax - ymin

    def writeStackVRT(self):
        dates = []
        for slc in self.slcList:
            width, height = self._size(slc)
            outname = os.path.basename(os.path.dirname(slc))
            stackio.write_raw_vrt(os.path.join(self.outSlcVrtDir, outname + '.vrt'), slc, width, height)
            dates.append(outname)
        self.stackVRT = os.path.join(self.outStackVrtDir, 'stack.vrt')
        print("writing ", self.stackVRT)
        with open(self.stackVRT, 'w') as fid:
            _, _, _, _, xsize, ysize = self.get_x_y_offsets(-1)
            fid.write('<VRTDataset rasterXSize="{0}" rasterYSize="{1}">\n'.format(xsize, ysize))
            for ind, date in enumerate(dates):
                width, height, xmin, ymin, xs, ys = self.get_x_y_offsets(ind)
                fid.write(stackio.STACK_BAND.format(width=width, height=height, xmin=xmin, ymin=ymin, xsize=xs,
                                                    ysize=ys, date=date, acq=date, wvl=0.03, index=ind + 1, extra="",
                                                    path=os.path.abspath(os.path.join(self.outSlcVrtDir, date + '.vrt'))))
            fid.write('</VRTDataset>')


class MiniStack(Stack):
    """src/sequential/Stack.py:177-190; the compressed-SLC list is sorted here (the reference's glob is
    not, Stack.py:184 -- the result is order-invariant)."""

    def updateMiniStack(self, compressedSlcDir):
        compSlcList = sorted(glob.glob(os.path.join(compressedSlcDir, '*/*.slc')))
        self.slcList = compSlcList + self.slcList
        self.applyBbox = [False] * len(compSlcList) + self.applyBbox


def runEvd(inps, inputDataset, weightDS, outDir, miniStackCount, compressedSlcDir=None, compressedSlcName=None):
    use_bindings()
    import evdlib
    aa = evdlib.Evd()                      # method stays at the struct default "MLE" (evd.hpp:66)
    aa.inputDS = inputDataset
    aa.weightsDS = weightDS
    aa.outputFolder = outDir
    aa.miniStackCount = miniStackCount
    aa.blocksize = inps.linesPerBlock
    aa.memsize = inps.memorySize
    aa.halfWindowX = inps.halfWindowX
    aa.halfWindowY = inps.halfWindowY
    aa.minimumNeighbors = inps.minNeighbors
    aa.outputCompressedSlcFolder = compressedSlcDir if compressedSlcDir else aa.outputFolder
    aa.compSlc = compressedSlcName if compressedSlcName else "compslc.bin"
    aa.run()


def main(argv=None):
    inps = cmdLineParser(argv)
    weightDS = inps.weightDS
    inps.outputDir = os.path.abspath(inps.outputDir)
    outDir = os.path.join(inps.outputDir, "fullStack")
    compressedSlcDir = os.path.join(inps.outputDir, "compressedSlc")
    os.makedirs(compressedSlcDir, exist_ok=True)

    stack = Stack(inps.inputDir)
    if inps.bbox is not None:
        inps.bbox = tuple(int(i) for i in inps.bbox)
        print('input bounding box in (y0, y1, x0, x1): {}'.format(inps.bbox))
    stack.bbox = inps.bbox
    stack.gatherSLCs()
    stack.getDates()
    stack.configure(outDir)
    stack.writeStackVRT()

    miniStackCount = 0
    indStart = 0
    while indStart < stack.size:
        miniStackCount += 1
        indEnd = min(indStart + inps.miniStackSize, stack.size)
        startDate, endDate = stack.dateList[indStart], stack.dateList[indEnd - 1]
        outDir = os.path.join(inps.outputDir, "miniStacks/" + startDate + "_" + endDate)
        if os.path.isdir(outDir) and (not inps.forceprocessing):
            print('{0} looks like it has already been processed. Skipping ... '.format(outDir))
        else:
            print('Processing {0}'.format(outDir))
            miniStack = MiniStack()
            miniStack.slcList = stack.slcList[indStart:indEnd]
            miniStack.getSize()
            miniStack.applyBbox = [True] * miniStack.size
            miniStack.updateMiniStack(compressedSlcDir)
            miniStack.bbox = inps.bbox
            miniStack.getDates()
            miniStack.configure(outDir)
            miniStack.writeStackVRT()
            evdDir = os.path.join(outDir, "EVD")
            if inps.forceprocessing and os.path.isdir(evdDir):
                import shutil
                shutil.rmtree(evdDir)
            compressedSlcName = miniStack.dateList[-1] + ".slc"
            compSlcDir = os.path.join(compressedSlcDir, miniStack.dateList[-1])
            os.makedirs(compSlcDir, exist_ok=True)
            runEvd(inps, miniStack.stackVRT, weightDS, evdDir, miniStackCount, compSlcDir, compressedSlcName)
            comp = os.path.join(compSlcDir, compressedSlcName)
            w, h = stackio.raster_size(comp)
            stackio.write_raw_vrt(comp + ".vrt", comp, w, h)
        indStart += inps.miniStackSize

    outDir = os.path.join(inps.outputDir, "Datum_connection")
    compSlcStack = Stack(compressedSlcDir)
    compSlcStack.gatherSLCs()
    compSlcStack.bbox = None
    compSlcStack.applyBbox = [False] * compSlcStack.size
    compSlcStack.getDates()
    compSlcStack.configure(outDir)
    compSlcStack.writeStackVRT()
    runEvd(inps, compSlcStack.stackVRT, weightDS, outDir + "/EVD", 1)


if __name__ == '__main__':
    main()
