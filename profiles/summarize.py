#!/usr/bin/env python
"""The ncu launch list of the bench command (`--metrics gpu__time_duration.sum --clock-control none`) ->
profiles/<tag>_launches.csv: per kernel the launches, total time, share of the step and average.

    python profiles/summarize.py <round tag> <launches.csv>

(`ncu --set full` captures are summarised by summarize_kernel.py; make_traffic.py derives traffic.json from them.)
"""
import csv
import os
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))


def launches(tag, path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "")[:80]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(HERE, f"{tag}_launches.csv")
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,share_pct,avg_ms\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{ns / 1e6:.3f},{100 * ns / tot:.2f},{ns / 1e6 / n:.4f}\n")
    return out


if __name__ == "__main__":
    print(launches(sys.argv[1], sys.argv[2]))
