#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: headline counters, stall reasons, opcode mix and
executed warp instructions per source line (needs -lineinfo and --import-source on).

usage: analyze.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top_n] [annot_out]
Runs here (no GPU): ncu -i ... --page raw/source --csv, cuobjdump -xelf, nvdisasm -g -c.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
annot_out = sys.argv[5] if len(sys.argv) > 5 else None


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, check=False, **kw).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
row = next(r for r in raw[2:] if kern.split("INS_")[0].replace("_ZN3mcb", "")[2:] in r[hdr.index("Kernel Name")]
           or True)
d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]
print("kernel:", d["Kernel Name"][0][:100])
for k in KEYS:
    if k in d:
        print("  %-62s %s %s" % (k, d[k][0], d[k][1]))
st = {k: float(v[0]) for k, v in d.items()
      if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v[0] not in ("", "n/a")}
tot = sum(st.values()) or 1
print("stall reasons (share of samples):")
for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
    print("  %-40s %5.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))

src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
shdr = next(r for r in src if "Address" in r and "Source" in r)
data = [r for r in src if len(r) == len(shdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(shdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = next(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin"))
lines = run(["nvdisasm", "-g", "-c", cubin]).split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur, locs, text = ("?", 0), [], []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        locs.append(cur)
        text.append(l.strip())
n = min(len(locs), len(data))
print("sass instructions: nvdisasm %d, ncu rows %d" % (len(locs), len(data)))
ex, sm, thr, ops = (collections.Counter() for _ in range(4))
out = []
for k in range(n):
    r = data[k]
    e = int(r[ix["Instructions Executed"]])
    ex[locs[k]] += e
    thr[locs[k]] += int(r[ix["Thread Instructions Executed"]])
    sm[locs[k]] += int(r[ix["# Samples"]])
    ins = re.sub(r"^/\*[0-9a-f]+\*/\s+", "", text[k])
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    ops[ins.split()[0].split(".")[0]] += e
    out.append("%5d %-22s %9d %5.1f %5d  %s" % (
        k, "%s:%d" % locs[k], e, int(r[ix["Thread Instructions Executed"]]) / max(e, 1),
        int(r[ix["# Samples"]]), text[k][:100]))
if annot_out:
    open(annot_out, "w").write("\n".join(out))
tot, tots = sum(ex.values()) or 1, sum(sm.values()) or 1
print("executed warp instructions (this launch): %d" % tot)
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(16)))
print("%-28s %12s %7s %7s %6s" % ("file:line", "warp_inst", "share", "stall%", "thr/w"))
for loc, v in ex.most_common(topn):
    print("%-28s %12d %7.3f %7.3f %6.1f" % ("%s:%d" % loc, v, v / tot, sm[loc] / tots,
                                            thr[loc] / max(v, 1)))
byfile = collections.Counter()
for loc, v in ex.items():
    byfile[loc[0]] += v
print({k: round(v / tot, 3) for k, v in byfile.most_common()})
