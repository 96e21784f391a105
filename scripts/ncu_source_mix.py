"""Instruction mix / shared-memory wavefronts per opcode and per CUDA line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
cur = None
seen = {}
curline = None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[2] == '-':
        curline = (cur.split('/')[-1], r[0])
        continue
    d = dict(zip(hdr, r))
    try:
        a = int(r[2], 16)
    except ValueError:
        continue
    if a in seen:
        continue
    f = lambda k: float(d.get(k) or 0)
    seen[a] = (r[3].strip(), f("Instructions Executed"), f("# Samples"), curline, f("L1 Wavefronts Shared"))
tot = sum(v[1] for v in seen.values())
print("total warp instr %.4g per SM %.4g" % (tot, tot / 148))
c = collections.Counter()
cs = collections.Counter()
for s, n, sm, cl, wf in seen.values():
    op = re.sub(r'^@!?U?P\w+\s+', '', s).split()[0].split('.')[0]
    c[op] += n
    cs[op] += sm
for op, n in c.most_common(45):
    print("%-12s %.4g %.1f%% samples %d" % (op, n, 100 * n / tot, cs[op]))
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    for a in sorted(seen):
        if lo <= (a & 0xfffff) <= hi:
            s, n, sm, cl, wf = seen[a]
            print(hex(a)[-5:], cl[1], "%-70s" % s[:70], "%.3g" % n, int(sm))
