#!/usr/bin/env python
"""Whole-process timing of the `methratio` command line (GPU pile-up) on BSMAP SAM output, next to the numpy
restatement of the reference's methratio.py (oracle/methratio_oracle.py, one host core, bounded sample).

    python tests/methratio_bench.py [--reads 4000000] [--len 100] [--genome-mb 200]
"""
import argparse, hashlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this script lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from bsmap_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=4_000_000)
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--genome-mb", type=float, default=200.0)
ap.add_argument("--oracle-lines", type=int, default=100_000)
a = ap.parse_args()
dev = "cuda" if torch.cuda.is_available() else "cpu"
td = tempfile.mkdtemp(prefix="bsx_meth_")
n_chr = 5
g = synth.make_genome(1, [int(a.genome_mb * 1e6 / n_chr)] * n_chr, device=dev)
sim = synth.simulate_reads(g, a.reads, a.len, seed=11, subs="cfg2")
fa, fq, sam = os.path.join(td, "ref.fa"), os.path.join(td, "reads.fq"), os.path.join(td, "aln.sam")
synth.write_fasta(fa, [x.cpu() for x in g])
synth.write_fastq(fq, sim["seq"].cpu(), synth.read_names({k: v.cpu() for k, v in sim.items() if k != "seq"}))
t0 = time.perf_counter()
subprocess.run([os.path.join(ROOT, "bsmap_b200", "bsmap"), "-a", fq, "-d", fa, "-o", sam, "-s", "16", "-v", "5", "-S", "7"], check=True, capture_output=True)
res = {"reads": a.reads, "read_len": a.len, "genome_mb": a.genome_mb, "bsmap_seconds": time.perf_counter() - t0, "sam_bytes": os.path.getsize(sam)}
runs = []
for k in range(3):
    out = os.path.join(td, "meth%d.txt" % k)
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "bsmap_b200", "methratio"), "-o", out, "-d", fa, "-q", sam], capture_output=True, text=True,
                       env=dict(os.environ, BSX_CLI_TIMING="1"))
    runs.append(time.perf_counter() - t0)
    assert r.returncode == 0, r.stderr
    res.setdefault("stages", []).append([l for l in r.stderr.splitlines() if "bsx timing" in l][-1])
    os.unlink(out) if k < 2 else None
res["methratio_seconds_runs"] = [round(x, 3) for x in runs]
res["methratio_alignments_per_s"] = a.reads / min(runs)
res["summary"] = r.stdout.strip()
res["table_bytes"] = os.path.getsize(out)
# fused: reads -> methratio table in one process, no alignment text in between (bsmap --methratio without -o)
fused = []
for k in range(3):
    out = os.path.join(td, "fused%d.txt" % k)
    t0 = time.perf_counter()
    r2 = subprocess.run([os.path.join(ROOT, "bsmap_b200", "bsmap"), "-a", fq, "-d", fa, "-s", "16", "-v", "5", "-S", "7", "--methratio", out],
                        capture_output=True, text=True, env=dict(os.environ, BSX_CLI_TIMING="1"))
    fused.append(time.perf_counter() - t0)
    assert r2.returncode == 0, r2.stdout[-300:] + r2.stderr[-300:]
res["fused_fastq_to_table_seconds_runs"] = [round(x, 3) for x in fused]
res["fused_identical_to_two_step"] = hashlib.md5(open(out, "rb").read()).hexdigest() == hashlib.md5(open(os.path.join(td, "meth2.txt"), "rb").read()).hexdigest()
res["fused_stages"] = [l for l in r2.stderr.splitlines() if "bsx timing" in l]
# bounded CPU sample: the first N alignment lines through the numpy restatement of methratio.py
import methratio_oracle as MO
small = os.path.join(td, "small.sam")
with open(sam) as f, open(small, "w") as o:
    for i, line in enumerate(f):
        if i >= a.oracle_lines: break
        o.write(line)
names = ["chr%d" % (i + 1) for i in range(n_chr)]
seqs = [bytes(x.cpu().numpy()) for x in g]
t0 = time.perf_counter()
MO.parse_alignments(small, set(names))
txt, st = MO.methratio(names, seqs, [small])
dt = time.perf_counter() - t0
res["cpu_port"] = {"kind": "port (numpy restatement of methratio.py)", "cores": 1, "lines": a.oracle_lines, "seconds": dt,
                   "alignments_per_s": st[0] / dt, "note": "includes the per-position report over the whole genome"}
print(json.dumps(res))
