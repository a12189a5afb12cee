#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
    python tools/ncu_source.py REPORT KERNEL_REGEX [rows]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
nrows = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, items, first_fn = "", None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if first_fn is None:
            first_fn = r[1]
        elif r[1] != first_fn:
            break                     # only the first captured instance
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        # the header has two "Source" columns (cuda line text, sass); zip keeps the last -> re-read the first
        def num(k):
            try:
                return float(d.get(k) or 0)
            except ValueError:
                return 0.0
        items.append((cur_file, int(r[0]), r[1], num("# Samples"), num("Instructions Executed")))
tot_s = sum(i[3] for i in items) or 1
tot_i = sum(i[4] for i in items) or 1
print((first_fn or "")[:110])
print(f"warp-instructions {tot_i:.0f}   stall samples {tot_s:.0f}")
for f, ln, src, smp, ins in sorted(items, key=lambda i: -i[3])[:nrows]:
    print(f"{100 * smp / tot_s:5.1f}% smp {100 * ins / tot_i:5.1f}% inst  {f}:{ln:<4d} {src.strip()[:100]}")
