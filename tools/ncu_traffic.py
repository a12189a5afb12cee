#!/usr/bin/env python3
"""Turn the ncu launch list of one bench.py run into the per-group numbers bench.py quotes.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        -k regex:'^(void )?(rt2?_|ia_|at_|df_|pj_|ls_|sf_|cm_|pt_|fs_|ss_|f2_|co_)' --csv --log-file gpurun_out/launches.csv \\
        python bench.py --steps 1 --warmup 3 --graph off --no-recon-probe --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/launches.csv profiles/r02_traffic.json profiles/r02_ncu_launches_our_kernels_in_step.txt

Groups are bench.KERNEL_GROUPS.  The run issues several identical eager steps (warm-up, timed, profiling, e2e); the number
of steps is the launch count of a once-per-step kernel (rt_finalize_kernel), and every figure is per step.  ncu times are
cold-cache and serialised: what must agree with the bench line is each group's SHARE, not the absolute time."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, out_json, out_txt = sys.argv[1], sys.argv[2], sys.argv[3]
    import bench
    pats = [(g, re.compile(p)) for g, p in bench.KERNEL_GROUPS]
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    def to_bytes(d, k):
        u = d.get("unit:" + k, "byte").lower()
        return d.get(k, 0.0) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    def to_us(d):
        u = d.get("unit:gpu__time_duration.sum", "ns").lower()
        return d.get("gpu__time_duration.sum", 0.0) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0}.get(u, 1e-3)
    steps = sum(1 for d in launches.values() if "rt_finalize_kernel" in d["name"]) or 1
    groups = collections.OrderedDict()
    kernels = collections.OrderedDict()
    for d in launches.values():
        grp = next((g for g, p in pats if p.search(d["name"])), "other")
        e = groups.setdefault(grp, {"launches": 0, "us": 0.0, "bytes": 0.0})
        e["launches"] += 1
        e["us"] += to_us(d)
        e["bytes"] += to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum")
        short = re.sub(r"\(.*$", "", d["name"]).replace("void ", "")
        k = kernels.setdefault((grp, short, d["grid"], d["block"]), {"launches": 0, "us": 0.0, "bytes": 0.0})
        k["launches"] += 1
        k["us"] += to_us(d)
        k["bytes"] += to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum")
    total_us = sum(e["us"] for e in groups.values()) or 1.0
    res = {"source": os.path.basename(src), "steps_seen": steps, "groups": {}}
    for g, e in sorted(groups.items(), key=lambda kv: -kv[1]["us"]):
        res["groups"][g] = {"launches_per_step": round(e["launches"] / steps, 2), "us_per_step": round(e["us"] / steps, 2),
                            "share_of_our_kernels": round(e["us"] / total_us, 4),
                            "dram_bytes_per_step": int(e["bytes"] / steps),
                            "dram_bytes_per_launch": int(e["bytes"] / max(e["launches"], 1))}
    with open(out_json, "w") as f:
        json.dump(res, f, indent=1)
    with open(out_txt, "w") as f:
        f.write(f"# ncu launch list of OUR kernels inside bench.py steps ({steps} eager steps seen; per-step figures; cold-cache, serialised)\n")
        f.write(f"# {'group':16s} {'launches':>8s} {'us/step':>10s} {'share':>7s} {'DRAM MB/step':>13s}\n")
        for g, v in res["groups"].items():
            f.write(f"  {g:16s} {v['launches_per_step']:8.1f} {v['us_per_step']:10.1f} {100 * v['share_of_our_kernels']:6.1f}% "
                    f"{v['dram_bytes_per_step'] / 1e6:13.2f}\n")
        f.write("\n# per kernel and launch shape\n")
        for (grp, name, grid, block), k in sorted(kernels.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"  {grp:14s} {name[:58]:58s} grid={grid:18s} block={block:12s} n/step={k['launches'] / steps:5.1f} "
                    f"us={k['us'] / k['launches']:8.2f} DRAM_MB={k['bytes'] / k['launches'] / 1e6:8.2f}\n")
    print(json.dumps(res)[:600])


if __name__ == "__main__":
    main()
