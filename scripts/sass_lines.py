"""Attribute SASS instructions (nvdisasm -gi output) to source lines: python scripts/sass_lines.py dis src [basename]"""
import collections
import re
import sys
dis, srcp = sys.argv[1], sys.argv[2]
base = sys.argv[3] if len(sys.argv) > 3 else srcp.split('/')[-1]
cur = None; cnt = collections.Counter(); tot = collections.Counter(); kinds = collections.defaultdict(collections.Counter)
for line in open(dis):
    m = re.search(r'//## File ".*%s", line (\d+)' % re.escape(base), line)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(2).split('.')[0]; tot[cur] += 1; kinds[cur][op] += 1
        if op in ('STL', 'LDL'):
            cnt[cur] += 1
src = open(srcp).read().split('\n')
print("total", sum(tot.values()))
print("spill sites:")
for ln, c in sorted(cnt.items(), key=lambda x: -x[1])[:25]:
    print(ln, c, '|', src[ln - 1].strip()[:90])
print("top lines:")
for ln, c in sorted(tot.items(), key=lambda x: -x[1])[:14]:
    print(ln, c, dict(kinds[ln].most_common(7)), '|', src[ln - 1].strip()[:60])
