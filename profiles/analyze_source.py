#!/usr/bin/env python
"""Per-function / per-line instruction and stall-sample breakdown of an ncu report (source page).
    python profiles/analyze_source.py <report.ncu-rep> <reads in capture> [top lines]"""
import csv, re, subprocess, sys, os
rep, reads = sys.argv[1], float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "bsmap_b200/csrc/bsx_map_impl.cuh")).read().splitlines()
funcs, cur = [], "?"
for l in src:
    m = re.match(r'^(?:__device__|__global__).*?\b(\w+)\s*\(', l)
    if m and not l.strip().startswith('//'): cur = m.group(1)
    m2 = re.match(r'^(\w+)\(const __grid_constant__', l)
    if m2: cur = m2.group(1)
    funcs.append(cur)
agg, lines, tot, curf = {}, [], 0, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': curf = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0].isdigit():
        try: s = int(r[4]); inst = int(r[7])
        except ValueError: continue
        ln = int(r[0])
        key = funcs[ln - 1] if curf == 'bsx_map_impl.cuh' and ln <= len(funcs) else curf
        a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += inst; tot += inst
        lines.append((inst, s, curf, ln, r[1].strip()[:90]))
ts = sum(a[0] for a in agg.values())
print(f"warp instructions per read: {tot / reads:.0f}")
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:28s} inst/read {i / reads:8.0f} ({100 * i / tot:4.1f}%)  samples {100 * s / ts:4.1f}%")
print()
lines.sort(reverse=True)
for inst, s, f, ln, t in lines[:top]:
    print(f"{inst / reads:6.0f} inst/read {100 * s / ts:4.1f}% smp {f}:{ln} {t}")
