#!/usr/bin/env python
"""Summarise an ncu report as a markdown table: for every kernel name the row of its longest launch.

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv
"""
import csv
import re
import sys

COLS = [("us", "gpu__time_duration.sum", 1e-3), ("grid", "launch__grid_size", 1), ("regs", "launch__registers_per_thread", 1),
        ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("rdMB", "dram__bytes_read.sum", None),
        ("wrMB", "dram__bytes_write.sum", None), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("l1hit%", "l1tex__t_sector_hit_rate.pct", 1),
        ("l2hit%", "lts__t_sector_hit_rate.pct", 1), ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1)]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    idx = {n: i for i, n in enumerate(names)}
    kcol = idx["Kernel Name"]
    best = {}
    count = {}
    for r in data:
        if len(r) <= kcol:
            continue
        k = re.sub(r"\(.*", "", r[kcol]).replace("<unnamed>::", "").replace("void ", "")
        t = num(r[idx["gpu__time_duration.sum"]]) * TO_US.get(units[idx["gpu__time_duration.sum"]], 1.0)
        count[k] = count.get(k, 0) + 1
        if k not in best or t > best[k][0]:
            best[k] = (t, r)
    print("| kernel | launches | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for k, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        cells = []
        for label, metric, scale in COLS:
            if metric not in idx:
                cells.append("-")
                continue
            v, u = num(r[idx[metric]]), units[idx[metric]]
            if label == "us":
                v = t
            elif scale is None:
                v *= TO_MB.get(u, 1.0)
            cells.append("%.4g" % v)
        print("| %s | %d | %s |" % (k, count[k], " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1])
