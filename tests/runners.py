"""Run a parity case through the oracle restatement or through the unmodified reference binary."""
from __future__ import annotations

import gzip
import os
import tempfile

import numpy as np

import cases as CS
import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_quals(case):
    d = case.data()
    if case.fasta_reads:   # reads.cpp:108: FASTA reads get qual = 'I' * len
        return [b"I" * len(s) for s in d["seqs"]]
    return d["quals"]


def clip(case, seqs):
    L = case.opts.get("L", 144)
    return [s[:min(L, 144)] for s in seqs]


def oracle_run(case):
    """-> dict(main=bytes, unpair=bytes, recs..., stats)"""
    d = case.data()
    p = O.make_params(**case.param_kwargs())
    ref = O.OracleRef(p, d["gnames"], d["gseqs"])
    out = {}
    head = ref.header() if p.out_sam else b""
    if not case.paired:
        buf, lens = O.pack_reads(clip(case, d["seqs"]))
        recs, counts, stats = ref.map_se(buf, lens)
        txt, na = ref.format_se(d["names"], d["seqs"], case_quals(case), recs, counts)
        out.update(main=head + txt, unpair=b"", recs=recs, counts=counts, stats=stats, n_aligned=na)
    else:
        ba, la = O.pack_reads(clip(case, d["seqs"]))
        bb, lb = O.pack_reads(clip(case, d["seqs_b"]))
        pr, ra, rb, ca, cb, stats = ref.map_pe(ba, la, bb, lb)
        txt, un, st = ref.format_pe(d["names"], d["seqs"], d["quals"], d["names_b"], d["seqs_b"], d["quals_b"],
                                    pr, ra, rb, ca, cb)
        out.update(main=head + txt, unpair=un, pr=pr, ra=ra, rb=rb, ca=ca, cb=cb, stats=stats, n_aligned=st)
    ref.close()
    return out


def reference_run(case, keep_dir=None):
    """run oracle/_ref/bsmap -p 1 on the case -> (main bytes, unpair bytes, stdout)"""
    with tempfile.TemporaryDirectory() as td:
        td = keep_dir or td
        fa, a, b = CS.write_inputs(case, td)
        o = os.path.join(td, "out." + case.out_ext)
        o2 = os.path.join(td, "out_unpair.bsp") if (case.paired and case.out_ext != "sam") else None
        args = case.cli(a, b, fa, o, o2) + ["-p", "1"]
        stdout = O.run_reference(args, cwd=td)
        main = open(o, "rb").read()
        un = open(o2, "rb").read() if o2 else b""
    return main, un, stdout


def golden_paths(case):
    return (os.path.join(GOLDEN, f"{case.name}.{case.out_ext}.gz"),
            os.path.join(GOLDEN, f"{case.name}.unpair.bsp.gz"))


def golden_load(case):
    m, u = golden_paths(case)
    main = gzip.open(m, "rb").read()
    un = gzip.open(u, "rb").read() if os.path.exists(u) else b""
    return main, un


def first_diff(a: bytes, b: bytes, ctx=2):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i}:\n  got: {x[:300]!r}\n  exp: {y[:300]!r}"
    return f"length differs: {len(la)} vs {len(lb)} lines"
