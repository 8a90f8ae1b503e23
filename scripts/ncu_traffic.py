#!/usr/bin/env python3
"""DRAM traffic of one kernel launch from an ncu --set full capture, written as the small JSON bench.py reads for
`roofline.traffic` (bytes per pixel of that launch, scaled there by the pixels of the timed launch).
usage: ncu_traffic.py report.ncu-rep <kernel substring> <pixels of the captured launch> <config name> > profiles/rN_<kernel>_traffic.json"""
import csv
import json
import subprocess
import sys


def main(path, kernel, pixels, config):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, zip(units, r)))
        if kernel not in d["Kernel Name"][1]:
            continue
        def val(k):
            unit, v = d[k]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            return float(v) * scale
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        print(json.dumps({"kernel": d["Kernel Name"][1], "config": config, "pixels": int(pixels), "dram_bytes_read": rd,
                          "dram_bytes_write": wr, "bytes_per_pixel": (rd + wr) / float(pixels), "source": path.split("/")[-1]}))
        return
    raise SystemExit("kernel not found")


if __name__ == "__main__":
    main(*sys.argv[1:5])
