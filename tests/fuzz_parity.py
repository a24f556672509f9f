#!/usr/bin/env python
"""Randomised differential test: GPU path (C ABI) against the CPU oracle over random parameter sets, read lengths,
adapters, N-rich and trimmed reads, single- and paired-end, WGBS and RRBS.  Records, per-level counts and the
candidate counter must agree exactly.

    python tests/fuzz_parity.py [--rounds 40] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this script lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O      # noqa: E402  (test infrastructure: the checker)
import bsmap_b200 as B      # noqa: E402
from bsmap_b200 import synth  # noqa: E402

ADAPTER = "AGATCGGAAGAGC"


def mutate_reads(rng, seqs, L, adapter_frac, n_frac):
    out = []
    for s in seqs:
        s = bytearray(s)
        if rng.random() < adapter_frac:                      # adapter read-through at a random cut
            cut = int(rng.integers(max(8, L // 3), L))
            tail = (ADAPTER.encode() * 12)[:L - cut]
            s[cut:] = tail
        if rng.random() < n_frac:                            # Ns, sometimes more than -f allows
            for _ in range(int(rng.integers(1, 9))):
                s[int(rng.integers(0, len(s)))] = ord("N")
        if rng.random() < 0.05:
            s = s[:int(rng.integers(5, len(s)))]             # ragged lengths, some shorter than the seed
        if rng.random() < 0.05:
            s = bytearray(bytes(s).lower())
        out.append(bytes(s))
    return out


def one_round(rng, k):
    rrbs = rng.random() < 0.2
    paired = rng.random() < 0.35
    L = int(rng.choice([36, 50, 75, 100, 125, 144]))
    kw = dict(v=int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15])), w=int(rng.choice([1, 2, 3, 20, 100, 1000])), r=int(rng.integers(0, 2)),
              S=int(rng.integers(1, 1000)), n=int(rng.random() < 0.25), f=int(rng.choice([0, 2, 5])),
              L=int(rng.choice([144, 144, 60, 90])))
    if rrbs:
        kw["D"] = "C-CGG"
    else:
        kw["s"] = int(rng.integers(8, 17)); kw["I"] = int(rng.choice([1, 2, 3, 4, 4, 4, 5, 8, 16]))
    if rng.random() < 0.4:
        kw["A"] = [ADAPTER]
    if paired:
        kw.update(pairend=1, m=int(rng.choice([28, 60])), x=int(rng.choice([300, 500])))
    n_chr = int(rng.integers(1, 5))
    lens = [int(rng.integers(60_000, 400_000)) for _ in range(n_chr)]
    g = synth.make_genome(int(rng.integers(1, 10_000)), lens)
    if rng.random() < 0.5:
        g = synth.plant_repeats(g, 3, unit_len=int(rng.choice([120, 250])), copies=int(rng.choice([20, 200])), divergence=0.02)
    gb = [x.numpy().tobytes() for x in g]
    if rng.random() < 0.3:                                    # an N run in the reference
        b = bytearray(gb[0]); ln = int(rng.integers(10, 3000)); b[1000:1000 + ln] = b"N" * ln; gb[0] = bytes(b)
    names = [f"chr{i + 1}" for i in range(n_chr)]
    n = int(rng.choice([200, 1500, 4000]))
    op, p = O.make_params(**kw), B.make_params(**kw)
    oref = O.OracleRef(op, names, gb)
    ix = B.Index(p, names, gb)
    mp = B.Mapper(ix, p, max_batch=int(rng.choice([64, 1000, 8192])), stride=int(rng.choice([144, 152, 160])))
    desc = dict(round=k, read_len=L, paired=paired, reads=n, genome=lens, opts=dict(kw))
    bad = None
    if not paired:
        if rrbs:
            reads, _ = synth.simulate_rrbs(gb, n, L, int(rng.integers(1, 1000)), adapter=ADAPTER.encode())
        else:
            sim = synth.simulate_reads(g, n, L, seed=int(rng.integers(1, 1000)), subs=str(rng.choice(["cfg1", "cfg2", "cfg5"])))
            reads = [bytes(r) for r in sim["seq"].numpy()]
        reads = mutate_reads(rng, reads, L, 0.2 if "A" in kw else 0.02, 0.1)
        clip = [s[:min(kw["L"], 144)] for s in reads]
        obuf, olens = O.pack_reads(clip)
        orec, ocnt, ostats = oref.map_se(obuf, olens)
        buf, lens_ = B.pack_reads(clip, stride=mp.stride)
        recs, counts = mp.map_se(buf, lens_)
        for f in orec.dtype.names:
            if not np.array_equal(recs[f], orec[f]):
                i = int(np.nonzero(recs[f] != orec[f])[0][0]); bad = f"SE field {f} read {i}: gpu {recs[i]} oracle {orec[i]} read {reads[i]!r}"; break
        if bad is None and not np.array_equal(counts, ocnt):
            bad = "SE per-level counts differ"
        if bad is None and mp.stats()["candidates"] != int(ostats[0]):
            bad = f"candidate counter {mp.stats()['candidates']} != oracle {int(ostats[0])}"
        if bad is None:                                       # the packed entry point on the same batch, where it is exact
            pk, low = B.pack_reads_2bit(buf, lens_)
            if not (low and "A" in kw):
                precs, pcnt = mp.map_se_packed(pk, lens_)
                if not (np.array_equal(precs, recs) and np.array_equal(pcnt, counts)):
                    bad = "packed read input gives different records"
    else:
        sim = synth.simulate_pairs(g, n, L, seed=int(rng.integers(1, 1000)), frag_min=int(rng.choice([40, 150])), frag_max=450, subs="cfg2")
        ra = mutate_reads(rng, [bytes(r) for r in sim["seq1"].numpy()], L, 0.1 if "A" in kw else 0.0, 0.05)
        rb = mutate_reads(rng, [bytes(r) for r in sim["seq2"].numpy()], L, 0.1 if "A" in kw else 0.0, 0.05)
        ca = [s[:min(kw["L"], 144)] for s in ra]; cb = [s[:min(kw["L"], 144)] for s in rb]
        oa, ola = O.pack_reads(ca); ob, olb = O.pack_reads(cb)
        opr, ora, orb, oca, ocb, ostats = oref.map_pe(oa, ola, ob, olb)
        ba, la = B.pack_reads(ca, stride=mp.stride); bb, lb = B.pack_reads(cb, stride=mp.stride)
        pr, xa, xb, cnta, cntb = mp.map_pe(ba, la, bb, lb)
        for nm, got, exp in (("pair", pr, opr), ("a", xa, ora), ("b", xb, orb)):
            for f in exp.dtype.names:
                if not np.array_equal(got[f], exp[f]):
                    i = int(np.nonzero(got[f] != exp[f])[0][0]); bad = f"PE {nm}.{f} pair {i}: gpu {got[i]} oracle {exp[i]}"; break
            if bad:
                break
        if bad is None and not (np.array_equal(cnta, oca) and np.array_equal(cntb, ocb)):
            bad = "PE per-level counts differ"
    mp.close(); ix.close(); oref.close()
    return desc, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    fails = 0
    for k in range(a.rounds):
        desc, bad = one_round(rng, k)
        print(("FAIL " if bad else "ok   ") + str(desc) + ("  -> " + bad if bad else ""), flush=True)
        fails += bad is not None
    print(f"{a.rounds - fails}/{a.rounds} rounds identical")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
