"""Dynamic instruction counts and stall samples per source line: joins `ncu --page source --csv` (per SASS
instruction) with `nvdisasm -gi` of the same cubin (line info).
usage: python scripts/ncu_lines.py source.csv listing.dis 'k_mle<(int)20>' 'k_mleILi20E' src.cu [npix]"""
import collections
import csv
import re
import sys

csvp, disp, kname, mangled, srcp = sys.argv[1:6]
npix = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
line_of = {}
cur = None
infn = False
chain = False
for line in open(disp):
    if '.text.' in line and ('.section' in line or line.startswith('.text.')):
        infn = mangled in line
    if not infn:
        continue
    m = re.search(r'//## File "[^"]*?([^/"]+)", line (\d+)', line)      # innermost frame of an inlined chain
    if m:
        if 'inlined at' in line or not chain:           # innermost frame first; the frames of its callers follow
            cur = (m.group(1), int(m.group(2)))
        chain = True
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+', line)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
        chain = False
rows = list(csv.reader(open(csvp)))
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and kname in r[1]:
        start = i
        break
hdr = rows[start + 1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
ex = collections.Counter(); samp = collections.Counter()
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    ln = line_of.get(addr - base, ("?", 0))
    ex[ln] += int(r[ie]); samp[ln] += int(r[isamp])
src = open(srcp).read().split('\n')
tot, tots = sum(ex.values()), sum(samp.values())
print(f"total executed {tot} ({tot / npix:.0f} per pixel), samples {tots}, mapped offsets {len(line_of)}")
for ln, c in sorted(ex.items(), key=lambda x: -x[1])[:45]:
    text = src[ln[1] - 1].strip()[:70] if ln[0] == srcp.split('/')[-1] and ln[1] else ln[0]
    print(f"{ln[1]:5d} {c / npix:9.1f}/px {100 * c / tot:5.1f}% inst {100 * samp[ln] / max(tots, 1):5.1f}% samples | {text}")
