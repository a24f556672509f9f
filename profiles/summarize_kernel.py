#!/usr/bin/env python
"""One `ncu --set full` report -> profiles/<tag>.md (+ .json): the counters the design is argued from.
    python profiles/summarize_kernel.py <report.ncu-rep> <tag> <units in the captured launch> "<title>" """
import csv, json, os, subprocess, sys
rep, tag, units, title = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
HERE = os.path.dirname(os.path.abspath(__file__))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
val = {}
for i, n in enumerate(h):
    try:
        x = float(v[i].replace(",", ""))
    except ValueError:
        val[n] = v[i]; continue
    val[n] = x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Tbyte": 1e12, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6}.get(u[i], 1)
    if n == "gpu__time_duration.sum":   # keep the table in the report's unit, the rate in seconds
        dur_s = x * {"s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u[i], 1e-9)
        dur_unit = u[i]
W = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
     "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
     "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
kname = val.get("Kernel Name", "?")
out = {k: val.get(k) for k in W if k in val}
ld_s, ld_r = val.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0), val.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", 1)
per = {"dram_bytes_per_unit": (val.get("dram__bytes_read.sum", 0) + val.get("dram__bytes_write.sum", 0)) / units,
       "dram_read_bytes_per_unit": val.get("dram__bytes_read.sum", 0) / units, "dram_write_bytes_per_unit": val.get("dram__bytes_write.sum", 0) / units,
       "warp_instructions_per_unit": val.get("smsp__inst_executed.sum", 0) / units, "l2_read_sectors_per_unit": val.get("lts__t_sectors_srcunit_tex_op_read.sum", 0) / units,
       "global_load_sectors_per_request": ld_s / ld_r if ld_r else None,
       "units_per_second_under_ncu": units / dur_s if val.get("gpu__time_duration.sum") else None}
json.dump({"kernel": kname, "units_in_capture": units, "metrics": out, "derived": per}, open(os.path.join(HERE, tag + ".json"), "w"), indent=1)
with open(os.path.join(HERE, tag + ".md"), "w") as f:
    f.write(f"# ncu --set full: {title}\n\nkernel `{kname}`, {units:.0f} units in the captured launch (`--clock-control none`; cold-cache, serialised: shares, not absolutes)\n\n| metric | value |\n|---|---|\n")
    for k in W:
        if k in val:
            x = val[k]
            label = f"{k} ({dur_unit})" if k == "gpu__time_duration.sum" else k
            f.write(f"| {label} | {x:,.3f} |\n" if isinstance(x, float) else f"| {label} | {x} |\n")
    f.write("\n| derived | value |\n|---|---|\n")
    for k, x in per.items():
        f.write(f"| {k} | {x:,.2f} |\n" if x is not None else f"| {k} | n/a |\n")
print(json.dumps(per))
