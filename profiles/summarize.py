#!/usr/bin/env python
"""Turn raw gpurun_out/ ncu artefacts into the committed summaries under profiles/.

    python profiles/summarize.py <round tag> <launches.csv> <full.ncu-rep> [reads in the full capture]

Writes profiles/<tag>_launches.csv (per-kernel launch list summary of the bench command),
profiles/<tag>_se_kernel.json / .md (ncu --set full metrics of bsx_map_se_kernel) and refreshes
profiles/traffic.json (DRAM bytes per read, used by bench.py's roofline.traffic).
"""
import csv
import json
import os
import subprocess
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


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def full(tag, rep, reads):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals) if h in WANT}
    def num(k):
        return float(m[k][0].replace(",", "")) if k in m and m[k][0] not in ("", "n/a") else None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    rd = num("dram__bytes_read.sum") * scale[m["dram__bytes_read.sum"][1]]
    wr = num("dram__bytes_write.sum") * scale[m["dram__bytes_write.sum"][1]]
    summ = {"kernel": "bsx_map_se_kernel", "reads_in_capture": reads, "metrics": {k: {"value": v, "unit": u} for k, (v, u) in m.items()},
            "dram_bytes_per_read": (rd + wr) / reads, "l1_sectors_per_read": num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / reads,
            "l2_sectors_per_read": num("lts__t_sectors_srcunit_tex_op_read.sum") / reads,
            "warp_instructions_per_read": num("smsp__inst_executed.sum") / reads}
    json.dump(summ, open(os.path.join(HERE, f"{tag}_se_kernel.json"), "w"), indent=1)
    json.dump({"dram_bytes_per_read": summ["dram_bytes_per_read"], "source": f"profiles/{tag}_se_kernel.json",
               "reads_in_capture": reads}, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    with open(os.path.join(HERE, f"{tag}_se_kernel.md"), "w") as f:
        f.write(f"# ncu --set full: bsx_map_se_kernel ({tag}, {reads} reads, cfg2 genome)\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in WANT:
            if k in m:
                f.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
        f.write(f"\nDerived per read: DRAM bytes {summ['dram_bytes_per_read']:.0f}, L1 sectors {summ['l1_sectors_per_read']:.0f}, "
                f"L2 sectors {summ['l2_sectors_per_read']:.0f}, warp instructions {summ['warp_instructions_per_read']:.0f}\n")
    return summ


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    reads = int(sys.argv[4]) if len(sys.argv) > 4 else 4_000_000
    if os.path.exists(lcsv):
        print(launches(tag, lcsv))
    if os.path.exists(rep):
        print(json.dumps(full(tag, rep, reads), indent=1)[:600])
