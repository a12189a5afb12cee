#!/usr/bin/env python3
"""Aggregate per-source-line instruction / stall-sample shares of one kernel over all captured launches.
    python tools/ncu_lines.py REPORT KERNEL_REGEX [rows]"""
import collections
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
nrows = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = "", None
agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit():
        d = dict(zip(hdr, r))

        def num(k):
            try:
                return float(d.get(k) or 0)
            except ValueError:
                return 0.0
        key = (cur_file, int(r[0]))
        agg[key][0] += num("# Samples")
        agg[key][1] += num("Instructions Executed")
        agg[key][2] = r[1]
ti = sum(v[1] for v in agg.values()) or 1
ts = sum(v[0] for v in agg.values()) or 1
print(pat, "warp-instructions (all captured launches)", int(ti), "stall samples", int(ts))
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nrows]:
    print(f"{100 * v[1] / ti:5.1f}% inst {100 * v[0] / ts:5.1f}% smp  {f}:{ln:<4d} {v[2].strip()[:100]}")
byfile = collections.defaultdict(float)
for (f, ln), v in agg.items():
    byfile[f] += v[1]
print({k: round(100 * v / ti, 1) for k, v in byfile.items()})
