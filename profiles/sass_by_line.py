#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (per-SASS-instruction counters) with
`nvdisasm -g -c` line info, and aggregate executed warp instructions and stall samples per
source line.  Usage: sass_by_line.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-substr>"""
import collections
import csv
import re
import sys

ncu_csv, sass, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}

lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = ("?", 0)
locs = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        locs.append(cur)
n = min(len(locs), len(data))
print("sass instrs: nvdisasm %d, ncu %d" % (len(locs), len(data)))
ex = collections.Counter()
sm = collections.Counter()
thr = collections.Counter()
for k in range(n):
    r = data[k]
    ex[locs[k]] += int(r[ix["Instructions Executed"]])
    thr[locs[k]] += int(r[ix["Thread Instructions Executed"]])
    sm[locs[k]] += int(r[ix["# Samples"]])
tot, tots = sum(ex.values()), sum(sm.values())
print("%-28s %12s %7s %7s %6s" % ("file:line", "warp_inst", "share", "stall%", "thr/w"))
for loc, v in ex.most_common(topn):
    print("%-28s %12d %7.3f %7.3f %6.1f" % ("%s:%d" % loc, v, v / tot, sm[loc] / max(tots, 1),
                                            thr[loc] / max(v, 1)))
byfile = collections.Counter()
for loc, v in ex.items():
    byfile[loc[0]] += v
print({k: round(v / tot, 3) for k, v in byfile.most_common()})
