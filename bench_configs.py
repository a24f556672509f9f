#!/usr/bin/env python
"""bench_configs.py -- throughput of the other BASELINE.json configs (1, 3, 4, 5) on one B200.

bench.py is the graded bench (config 2).  This companion measures the remaining configs through the
same C ABI, kernel-only (inputs resident, CUDA events) and end to end (host buffers), and prints one
JSON line per config.  Parity for every config is covered by tests/; this script is about rates.

    python bench_configs.py [--configs cfg1,cfg3,cfg4,cfg5] [--genome-mb 3100] [--steps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed_kernel(torch, fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    return s, e0, e1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg3,cfg4,cfg5")
    ap.add_argument("--genome-mb", type=float, default=3100.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=4_000_000)
    a = ap.parse_args()
    import torch
    import bsmap_b200 as B
    from bsmap_b200 import synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream

    def big_genome(seed, repeats=0):
        n_chr = 25
        ln = int(a.genome_mb * 1e6 / n_chr)
        g = synth.make_genome(seed, [ln] * n_chr, device=dev)
        if repeats:
            g = synth.plant_repeats(g, seed, unit_len=300, copies=repeats, divergence=0.03)
        return g, [f"chr{i + 1}" for i in range(n_chr)], [ln] * n_chr

    def build_index(p, g, names, lens):
        host = [torch.empty(l, dtype=torch.uint8, pin_memory=True) for l in lens]
        for h, x in zip(host, g):
            h.copy_(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix = B.Index.from_pointers(p, names, [h.data_ptr() for h in host], lens, device=0)
        return ix, time.perf_counter() - t0

    def pin(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    def run_se(name, p, ix, seq_dev, L, stride, extra):
        n = seq_dev.shape[0]
        seq_h = pin(seq_dev); len_h = pin(torch.full((n,), L, dtype=torch.int16, device=dev))
        rec_h = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        mp = B.Mapper(ix, p, max_batch=n, stride=stride)
        mp.upload(n, seq_h.data_ptr(), len_h.data_ptr(), stream=st)
        for _ in range(2):
            mp.run_se(n, stream=st)
        torch.cuda.synchronize(); mp.stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstream)
        for _ in range(a.steps):
            mp.run_se(n, stream=st)
        e1.record(tstream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        s = mp.stats(reset=True)
        recs, _ = mp.download_se(n, stream=st)
        small = B.Mapper(ix, p, max_batch=min(n, 1 << 20), stride=stride)
        small.map_se_ptr(n, seq_h.data_ptr(), len_h.data_ptr(), rec_h.data_ptr())
        t0 = time.perf_counter()
        for _ in range(a.steps):
            small.map_se_ptr(n, seq_h.data_ptr(), len_h.data_ptr(), rec_h.data_ptr())
        e2e = (time.perf_counter() - t0) / a.steps
        out = dict(config=name, reads=n, kernel_reads_per_s=n / (ms * 1e-3), e2e_reads_per_s=n / e2e, ms_per_step=ms,
                   mapped_fraction=float((recs["nhits"] > 0).mean()), candidates_per_read=s["candidates"] / a.steps / n,
                   headers_per_read=s["probes"] / a.steps / n, hbm_gathers_per_read=s["gathers"] / a.steps / n,
                   full_extensions_per_read=s["full_extensions"] / a.steps / n, **extra)
        print(json.dumps(out), flush=True)
        mp.close(); small.close()

    for cfg in a.configs.split(","):
        if cfg == "cfg1":
            # 100k x 50 nt vs 5 Mb, -s 16 -v 2 -I 4 (the reference's own CPU-runnable case); 2 M reads for a stable rate
            g = synth.make_genome(1, [1_000_000] * 5, device=dev)
            p = B.make_params(s=16, v=2, I=4, S=7)
            ix, tb = build_index(p, g, [f"chr{i + 1}" for i in range(5)], [1_000_000] * 5)
            n = 2_000_000
            sim = synth.simulate_reads(g, n, 50, seed=11, subs="cfg1")
            seq = torch.zeros((n, 64), dtype=torch.uint8, device=dev); seq[:, :50] = sim["seq"]
            run_se("cfg1: 5 Mb genome, 50 nt SE, -s16 -v2 -I4", p, ix, seq, 50, 64, dict(index_seconds=tb))
            ix.close()
        elif cfg == "cfg5":
            g, names, lens = big_genome(5, repeats=2000)
            p = B.make_params(s=12, v=15, I=4, w=1000, S=7)
            ix, tb = build_index(p, g, names, lens)
            n = 200_000
            sim = synth.simulate_reads(g, n, 144, seed=55, subs="cfg5")
            seq = torch.zeros((n, 144), dtype=torch.uint8, device=dev); seq[:, :144] = sim["seq"]
            del g
            run_se("cfg5: 144 nt SE, -s12 -v15 -w1000, planted 2000-copy repeat", p, ix, seq, 144, 144, dict(index_seconds=tb, genome_mb=a.genome_mb))
            ix.close()
        elif cfg == "cfg3":
            g, names, lens = big_genome(2)
            p = B.make_params(s=16, v=2, I=4, m=28, x=500, S=7, pairend=1)
            ix, tb = build_index(p, g, names, lens)
            n = a.pairs
            s1 = torch.zeros((n, 112), dtype=torch.uint8, device=dev); s2 = torch.zeros((n, 112), dtype=torch.uint8, device=dev)
            CH = 1 << 20
            for s0 in range(0, n, CH):
                m = min(CH, n - s0)
                sim = synth.simulate_pairs(g, m, 100, seed=33, frag_min=150, frag_max=450, subs="cfg2", first_index=s0)
                s1[s0:s0 + m, :100] = sim["seq1"]; s2[s0:s0 + m, :100] = sim["seq2"]
            del g
            a_h, b_h = pin(s1), pin(s2)
            l_h = pin(torch.full((n,), 100, dtype=torch.int16, device=dev))
            torch.cuda.synchronize()
            mp = B.Mapper(ix, p, max_batch=n, stride=112)
            mp.upload(n, a_h.data_ptr(), l_h.data_ptr(), b_h.data_ptr(), l_h.data_ptr(), stream=st)
            for _ in range(2):
                mp.run_pe(n, stream=st)
            torch.cuda.synchronize(); mp.stats(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(tstream)
            for _ in range(a.steps):
                mp.run_pe(n, stream=st)
            e1.record(tstream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            s = mp.stats(reset=True)
            small = B.Mapper(ix, p, max_batch=1 << 19, stride=112)
            pr_h = torch.empty((n, 28), dtype=torch.uint8, pin_memory=True)
            ra_h = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True); rb_h = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
            args = (n, a_h.data_ptr(), l_h.data_ptr(), b_h.data_ptr(), l_h.data_ptr(), pr_h.data_ptr(), ra_h.data_ptr(), rb_h.data_ptr())
            small.map_pe_ptr(*args)
            t0 = time.perf_counter()
            for _ in range(a.steps):
                small.map_pe_ptr(*args)
            e2e = (time.perf_counter() - t0) / a.steps
            pr = np.frombuffer(pr_h.numpy().tobytes(), dtype=B.PAIR_REC)
            print(json.dumps(dict(config="cfg3: 2x100 nt PE, -m 28 -x 500, PairAlign on device", pairs=n, kernel_pairs_per_s=n / (ms * 1e-3),
                                  e2e_pairs_per_s=n / e2e, ms_per_step=ms, paired_fraction=float(pr["paired"].mean()),
                                  candidates_per_pair=s["candidates"] / a.steps / n, index_seconds=tb)), flush=True)
            mp.close(); small.close(); ix.close()
        elif cfg == "cfg4":
            # RRBS: 75-nt reads at digestion sites with adapter read-through, -D C-CGG -A; 50 Mb genome (host-side simulator)
            import cases as CS
            g = synth.make_genome(4, [10_000_000] * 5)
            gb = [x.numpy().tobytes() for x in g]
            n0 = 100_000
            reads, names_r = synth.simulate_rrbs(gb, n0, 75, 44, adapter=CS.ADAPTER.encode())
            p = B.make_params(D="C-CGG", v=2, S=5, A=[CS.ADAPTER])
            t0 = time.perf_counter()
            ix = B.Index(p, [f"chr{i + 1}" for i in range(5)], gb, device=0)
            tb = time.perf_counter() - t0
            buf, _ = B.pack_reads(reads, stride=80)
            rep = 20
            seq = torch.from_numpy(np.tile(buf, (rep, 1))).to(dev)
            run_se("cfg4: RRBS -D C-CGG, 75 nt reads with adapter trimming -A, 50 Mb genome", p, ix, seq, 75, 80, dict(index_seconds=tb))
            ix.close()


if __name__ == "__main__":
    main()
