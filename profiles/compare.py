#!/usr/bin/env python
"""Side-by-side per-read memory / issue metrics of several `ncu --set full` reports.
    python profiles/compare.py READS rep1.ncu-rep rep2.ncu-rep ..."""
import csv, subprocess, sys
reads = float(sys.argv[1]); reps = sys.argv[2:]
M = [("gpu__time_duration.sum", "ms", 1e-6 * 0 + 1), ("smsp__inst_executed.sum", "inst/read", None),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%", 1), ("dram__bytes_read.sum", "B/read", None),
     ("dram__bytes_write.sum", "B/read", None), ("lts__t_sector_hit_rate.pct", "%", 1), ("l1tex__t_sector_hit_rate.pct", "%", 1),
     ("lts__t_sectors_srcunit_tex_op_read.sum", "sect/read", None), ("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "sect/read", None),
     ("lts__t_sectors_srcunit_tex_op_write.sum", "sect/read", None), ("lts__t_sectors_srcunit_ltcfabric.sum", "sect/read", None),
     ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sect/read", None), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "sect/read", None),
     ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "sect/read", None), ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "sect/read", None),
     ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum", "sect/read", None), ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sect/read", None),
     ("l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum", "sect/read", None),
     ("l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum", "sect/read", None), ("l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum", "sect/read", None),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "", 1), ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "", 1),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "", 1), ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "", 1),
     ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "", 1), ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "", 1),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "", 1), ("launch__registers_per_thread", "", 1),
     ("launch__shared_mem_per_block_dynamic", "", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", 1)]
cols = []
for rep in reps:
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    d = {}
    for i, n in enumerate(h):
        try:
            x = float(v[i].replace(",", ""))
        except ValueError:
            continue
        unit = u[i]
        x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Tbyte": 1e12}.get(unit, 1)
        d[n] = x
    cols.append(d)
print(f"{'metric':90s}" + "".join(f"{r.split('/')[-1][:18]:>20s}" for r in reps))
for n, unit, scale in M:
    vals = [(c.get(n, float('nan')) / (reads if scale is None else 1)) for c in cols]
    print(f"{n + ' [' + unit + ']':90s}" + "".join(f"{x:20.2f}" for x in vals))
