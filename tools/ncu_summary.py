#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): one line of key metrics + top stall reasons per distinct kernel.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source KERNEL_SUBSTR]"""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "inst"), ("dram__bytes_read.sum", "dramR_MB"), ("dram__bytes_write.sum", "dramW_MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
            ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs")]
    for d in data:
        name = d[idx["Kernel Name"]]
        print(f"{name[:70]}  grid={d[idx['Grid Size']]} block={d[idx['Block Size']]}")
        print("    " + "  ".join(f"{lab}={d[idx[c]]}" for c, lab in cols if c in idx))
        items = []
        for h in hdr:
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    items.append((float(d[idx[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in items) or 1.0
        print("    stalls: " + ", ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(items, reverse=True)[:7]))


if __name__ == "__main__":
    main()
