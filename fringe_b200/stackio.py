"""GDAL-free readers / writers for the on-disk formats on this path:
flat CFloat32 SLC files, per-date raw VRTs, the stack VRT with per-band `slc` metadata, and ENVI
rasters.  Templates follow python/tops2vrt.py:144-152,201-225 and src/sequential/Stack.py:8-34."""
from __future__ import annotations

import os
import re

import numpy as np

RAW_VRT = '''<VRTDataset rasterXSize="{width}" rasterYSize="{height}">
    <VRTRasterBand dataType="CFloat32" band="1" subClass="VRTRawRasterBand">
        <SourceFilename>{path}</SourceFilename>
        <ImageOffset>0</ImageOffset>
        <PixelOffset>8</PixelOffset>
        <LineOffset>{linewidth}</LineOffset>
        <ByteOrder>LSB</ByteOrder>
    </VRTRasterBand>
</VRTDataset>'''

STACK_BAND = '''    <VRTRasterBand dataType="CFloat32" band="{index}">
        <SimpleSource>
            <SourceFilename>{path}</SourceFilename>
            <SourceBand>1</SourceBand>
            <SourceProperties RasterXSize="{width}" RasterYSize="{height}" DataType="CFloat32"/>
            <SrcRect xOff="{xmin}" yOff="{ymin}" xSize="{xsize}" ySize="{ysize}"/>
            <DstRect xOff="0" yOff="0" xSize="{xsize}" ySize="{ysize}"/>
        </SimpleSource>
        <Metadata domain="slc">
            <MDI key="Date">{date}</MDI>
            <MDI key="Wavelength">{wvl}</MDI>
            <MDI key="AcquisitionTime">{acq}</MDI>{extra}
        </Metadata>
    </VRTRasterBand>
'''

ENVI_TYPES = {np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3, np.dtype(np.float32): 4,
              np.dtype(np.complex64): 6, np.dtype(np.uint32): 13}
ENVI_DTYPES = {v: k for k, v in ENVI_TYPES.items()}


def write_envi(path: str, arr: np.ndarray, metadata: dict | None = None) -> None:
    """arr: (lines, cols) or (lines, cols, bands) -> BIP ENVI file + '<path>.hdr' (SUFFIX=ADD)."""
    arr = np.ascontiguousarray(arr)
    lines, cols = arr.shape[:2]
    bands = arr.shape[2] if arr.ndim == 3 else 1
    arr.tofile(path)
    with open(path + ".hdr", "w") as f:
        f.write(f"ENVI\nsamples = {cols}\nlines   = {lines}\nbands   = {bands}\nheader offset = 0\n"
                f"file type = ENVI Standard\ndata type = {ENVI_TYPES[arr.dtype]}\ninterleave = bip\nbyte order = 0\n")
        for k, v in (metadata or {}).items():
            f.write(f"{k} = {v}\n")


def write_envi_header(path: str, lines: int, cols: int, bands: int, dtype, metadata: dict | None = None) -> None:
    with open(path + ".hdr", "w") as f:
        f.write(f"ENVI\nsamples = {cols}\nlines   = {lines}\nbands   = {bands}\nheader offset = 0\n"
                f"file type = ENVI Standard\ndata type = {ENVI_TYPES[np.dtype(dtype)]}\ninterleave = bip\nbyte order = 0\n")
        for k, v in (metadata or {}).items():
            f.write(f"{k} = {v}\n")


def read_envi_header(path: str) -> dict:
    hdr = path + ".hdr" if os.path.exists(path + ".hdr") else os.path.splitext(path)[0] + ".hdr"
    out = {}
    for line in open(hdr):
        if "=" in line:
            k, v = line.split("=", 1)
            out[k.strip().lower()] = v.strip()
    return out


def read_envi(path: str) -> np.ndarray:
    h = read_envi_header(path)
    lines, cols, bands = int(h["lines"]), int(h["samples"]), int(h.get("bands", 1))
    dt = ENVI_DTYPES[int(h["data type"])]
    a = np.fromfile(path, dtype=dt, count=lines * cols * bands)
    return a.reshape(lines, cols, bands) if bands > 1 else a.reshape(lines, cols)


def raster_size(path: str) -> tuple[int, int]:
    """(width, height) of a .vrt or an ENVI raster."""
    if path.lower().endswith(".vrt"):
        txt = open(path).read()
        m = re.search(r'rasterXSize="(\d+)"\s+rasterYSize="(\d+)"', txt)
        return int(m.group(1)), int(m.group(2))
    h = read_envi_header(path)
    return int(h["samples"]), int(h["lines"])


def write_raw_vrt(vrt_path: str, data_path: str, width: int, height: int) -> None:
    with open(vrt_path, "w") as f:
        f.write(RAW_VRT.format(width=width, height=height, path=os.path.abspath(data_path), linewidth=8 * width))


def write_stack_vrt(stack_vrt: str, sources: list[tuple[str, str]], size: tuple[int, int], bbox=None,
                    extra_md: dict | None = None) -> None:
    """sources: [(date 'YYYYMMDD', path of per-date .vrt)]; size = (width, height) of the sources;
    bbox = (ymin, ymax, xmin, xmax) crop or None; extra_md = {date: {key: value}} extra slc-domain items
    (e.g. amplitudeConstant, read at nmap.cpp:209)."""
    width, height = size
    ymin, ymax, xmin, xmax = bbox if bbox else (0, height, 0, width)
    xsize, ysize = xmax - xmin, ymax - ymin
    with open(stack_vrt, "w") as f:
        f.write(f'<VRTDataset rasterXSize="{xsize}" rasterYSize="{ysize}">\n')
        for i, (date, path) in enumerate(sources):
            extra = "".join(f'\n            <MDI key="{k}">{v}</MDI>' for k, v in (extra_md or {}).get(date, {}).items())
            f.write(STACK_BAND.format(index=i + 1, path=os.path.abspath(path), width=width, height=height,
                                      xmin=xmin, ymin=ymin, xsize=xsize, ysize=ysize, date=date, wvl=0.03,
                                      acq=date, extra=extra))
        f.write("</VRTDataset>")


def read_stack_vrt(stack_vrt: str):
    """[(date, reader)] of a stack VRT written by write_stack_vrt / tops2vrt.py: reader() returns that band's
    (lines, cols) complex64 array (SrcRect crop applied)."""
    txt = open(stack_vrt).read()
    m = re.search(r'rasterXSize="(\d+)"\s+rasterYSize="(\d+)"', txt)
    xsize, ysize = int(m.group(1)), int(m.group(2))
    out = []
    for band in re.findall(r"<VRTRasterBand.*?</VRTRasterBand>", txt, flags=re.S):
        src = re.search(r"<SourceFilename[^>]*>(.*?)</SourceFilename>", band).group(1).strip()
        if not os.path.isabs(src):
            src = os.path.join(os.path.dirname(os.path.abspath(stack_vrt)), src)
        rect = re.search(r'<SrcRect xOff="(\d+)" yOff="(\d+)"', band)
        x0, y0 = (int(rect.group(1)), int(rect.group(2))) if rect else (0, 0)
        date = re.search(r'<MDI key="Date">(.*?)</MDI>', band).group(1).strip()

        def reader(src=src, x0=x0, y0=y0):
            raw = open(src).read()
            path = re.search(r"<SourceFilename[^>]*>(.*?)</SourceFilename>", raw).group(1).strip() if src.endswith(".vrt") else src
            w = int(re.search(r'rasterXSize="(\d+)"', raw).group(1)) if src.endswith(".vrt") else raster_size(src)[0]
            h = int(re.search(r'rasterYSize="(\d+)"', raw).group(1)) if src.endswith(".vrt") else raster_size(src)[1]
            if not os.path.isabs(path):
                path = os.path.join(os.path.dirname(src), path)
            arr = np.memmap(path, dtype=np.complex64, mode="r", shape=(h, w))
            return np.ascontiguousarray(arr[y0:y0 + ysize, x0:x0 + xsize])
        out.append((date, reader))
    return out


def default_dates(n: int, start: str = "20200101", step_days: int = 12) -> list[str]:
    import datetime
    d0 = datetime.datetime.strptime(start, "%Y%m%d")
    return [(d0 + datetime.timedelta(days=step_days * i)).strftime("%Y%m%d") for i in range(n)]


def make_stack_on_disk(root: str, slc: np.ndarray, dates: list[str] | None = None, extra_md=None) -> str:
    """Lay a (bands, lines, cols) complex64 stack out the way the reference's workflow does:
    <root>/SLC/<date>/<date>.slc (+ .hdr, + .vrt) and <root>/stack/stack.vrt.  Returns the stack VRT."""
    bands, lines, cols = slc.shape
    dates = dates or default_dates(bands)
    sources = []
    for b, date in enumerate(dates):
        d = os.path.join(root, "SLC", date)
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, date + ".slc")
        write_envi(path, slc[b].astype(np.complex64))
        write_raw_vrt(path + ".vrt", path, cols, lines)
        sources.append((date, path + ".vrt"))
    os.makedirs(os.path.join(root, "stack"), exist_ok=True)
    stack_vrt = os.path.join(root, "stack", "stack.vrt")
    write_stack_vrt(stack_vrt, sources, (cols, lines), extra_md=extra_md)
    return stack_vrt
