"""Small deterministic parity cases shared by the golden-fixture generator and the tests.

Each case yields a genome, reads (and mates), the bsmap command-line options, and the matching
parameter dictionary.  Inputs are regenerated from seeds (bsmap_b200/synth.py); the expected
outputs of the UNMODIFIED reference binary are committed under tests/golden/ by
tests/golden/make_golden.py.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bsmap_b200 import synth  # noqa: E402

ADAPTER = "AGATCGGAAGAGCGGTTCAGCAGGAATGCCGAGA"


class Case:
    def __init__(self, name, **kw):
        self.name = name
        self.opts = kw.pop("opts")             # dict: s, I, v, w, r, m, x, n, S, u, R, D, A, L
        self.out_ext = kw.pop("out_ext", "sam")
        self.paired = kw.pop("paired", False)
        self.maker = kw.pop("maker")
        self.fasta_reads = kw.pop("fasta_reads", False)
        self._data = None

    def data(self):
        if self._data is None:
            self._data = self.maker()
        return self._data

    def cli(self, a, b, d, o, o2=None):
        args = ["-a", a]
        if self.paired:
            args += ["-b", b]
        args += ["-d", d, "-o", o]
        if o2:
            args += ["-2", o2]
        for k in ("s", "v", "w", "I", "r", "m", "x", "n", "S", "L"):
            if k in self.opts:
                args += [f"-{k}", str(self.opts[k])]
        if "D" in self.opts:
            args += ["-D", self.opts["D"]]
        for ad in self.opts.get("A", ()):
            args += ["-A", ad]
        if self.opts.get("u"):
            args += ["-u"]
        if self.opts.get("R"):
            args += ["-R"]
        return args

    def param_kwargs(self):
        o = self.opts
        kw = dict(s=o.get("s", 16), I=o.get("I", 4), v=o.get("v", 2), w=o.get("w", 1000), r=o.get("r", 1),
                  m=o.get("m", 28), x=o.get("x", 500), n=o.get("n", 0), S=o.get("S", 0), L=o.get("L", 144),
                  u=o.get("u", 0), R=o.get("R", 0), D=o.get("D"), A=tuple(o.get("A", ())),
                  pairend=1 if self.paired else 0, out_sam=1 if self.out_ext == "sam" else 0)
        return kw


def _genome(seed, lens, n_runs=0, lower=0, repeats=0, repeat_unit=300):
    g = synth.make_genome(seed, lens)
    if repeats:
        g = synth.plant_repeats(g, seed, unit_len=repeat_unit, copies=repeats, divergence=0.03)
    g = [x.numpy().copy() for x in g]
    rng = np.random.default_rng(seed + 1000)
    for _ in range(n_runs):
        c = int(rng.integers(len(g)))
        p = int(rng.integers(0, len(g[c]) - 200))
        ln = int(rng.choice([1, 3, 10, 25, 40, 120]))
        g[c][p:p + ln] = ord("N")
    for _ in range(lower):
        c = int(rng.integers(len(g)))
        p = int(rng.integers(0, len(g[c]) - 500))
        g[c][p:p + 300] |= 0x20
    names = [f"chr{i + 1}" for i in range(len(g))]
    return names, [x.tobytes() for x in g]


def _se_reads(gbytes, n, L, seed, subs):
    import torch
    g = [torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()) for b in gbytes]
    sim = synth.simulate_reads(g, n, L, seed=seed, subs=subs)
    names = synth.read_names(sim)
    seqs = [bytes(r) for r in sim["seq"].numpy()]
    return names, seqs


def _quals(seqs, seed):
    rng = np.random.default_rng(seed)
    return [bytes(rng.integers(35, 74, size=len(s), dtype=np.uint8)) for s in seqs]


def mk_se(seed, lens, n, L, subs, **g):
    def f():
        names, gb = _genome(seed, lens, **g)
        rn, rs = _se_reads(gb, n, L, seed + 7, subs)
        return dict(gnames=names, gseqs=gb, names=rn, seqs=rs, quals=[b"I" * len(s) for s in rs])
    return f


def mk_se_mixed(seed, lens, n, s_=16, I_=4):
    """variable lengths, N's, lower case, adapters, junk reads, random qualities"""
    def q4(l):   # App. B Q4: lengths whose seed_start_offset is stale/uninitialised in the reference
        return (l - I_ + 1) % s_ == 0

    def f():
        names, gb = _genome(seed, lens, n_runs=30, lower=10)
        rng = np.random.default_rng(seed + 5)
        rn, rs = _se_reads(gb, n, 120, seed + 7, "cfg2")
        ok_lens = [l for l in range(20, 121) if not q4(l)]
        out = []
        for i, s in enumerate(rs):
            kind = int(rng.integers(10))
            l = int(rng.choice(ok_lens))
            s = bytearray(s[:l])
            if kind == 0:     # adapter read-through
                cut = int(rng.integers(20, max(21, l - 8)))
                while q4(cut):
                    cut += 1
                s = bytearray((bytes(s[:cut]) + ADAPTER.encode() + b"ACGT" * 30)[:l])
            elif kind == 1:   # a few N's
                for _ in range(int(rng.integers(1, 5))):
                    s[int(rng.integers(l))] = ord("N")
            elif kind == 2:   # too many N's -> QC
                for p in rng.choice(l, size=8, replace=False):
                    s[int(p)] = ord("N")
            elif kind == 3:   # lower case
                s = bytearray(bytes(s).lower())
            elif kind == 4:   # junk (unmappable)
                s = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=l).tobytes())
            elif kind == 5:   # short read below the seed size
                s = s[:int(rng.integers(5, 16))]
            out.append(bytes(s))
        return dict(gnames=names, gseqs=gb, names=rn, seqs=out, quals=_quals(out, seed + 9))
    return f


def mk_pe(seed, lens, n, L, fmin, fmax, subs="cfg1", nrich=True, **g):
    def f():
        import torch
        names, gb = _genome(seed, lens, **g)
        gt = [torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()) for b in gb]
        sim = synth.simulate_pairs(gt, n, L, seed=seed + 3, frag_min=fmin, frag_max=fmax, subs=subs)
        base = synth.read_names(dict(chrom=sim["chrom"], pos=sim["pos"], strand=sim["strand"]))
        s1 = [bytes(r) for r in sim["seq1"].numpy()]
        s2 = [bytes(r) for r in sim["seq2"].numpy()]
        rng = np.random.default_rng(seed + 11)
        # a few broken pairs: junk mate, N-rich mate
        for i in range(0, n, 17):
            s2[i] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L).tobytes()
        for i in range(5, n if nrich else 0, 41):
            b = bytearray(s1[i])
            for p in rng.choice(L, size=9, replace=False):
                b[int(p)] = ord("N")
            s1[i] = bytes(b)
        return dict(gnames=names, gseqs=gb, names=[x + "/1" for x in base], seqs=s1, quals=_quals(s1, seed + 12),
                    names_b=[x + "/2" for x in base], seqs_b=s2, quals_b=_quals(s2, seed + 13))
    return f


def mk_rrbs(seed, lens, n, L, paired=False):
    def f():
        names, gb = _genome(seed, lens)
        r = synth.simulate_rrbs(gb, n, L, seed + 2, adapter=ADAPTER.encode(), paired=paired)
        if paired:
            s1, s2, rn = r
            return dict(gnames=names, gseqs=gb, names=[x + "/1" for x in rn], seqs=s1, quals=_quals(s1, seed + 3),
                        names_b=[x + "/2" for x in rn], seqs_b=s2, quals_b=_quals(s2, seed + 4))
        s1, rn = r
        return dict(gnames=names, gseqs=gb, names=rn, seqs=s1, quals=_quals(s1, seed + 3))
    return f


def mk_selection_pin():
    """a 100-nt unit at Watson offsets 20000 / 90001 / 150002, six identical
    fully converted reads -> hits discovered in ascending position order, read idx takes hit
    myrand(idx) % 3."""
    def f():
        names, gb = _genome(77, [200_000])
        g = bytearray(gb[0])
        unit = bytes(g[5000:5100])
        for p in (20000, 90001, 150002):
            g[p:p + 100] = unit
        g[5000:5100] = bytes(g[6000:6100])
        read = unit.replace(b"C", b"T")
        return dict(gnames=names, gseqs=[bytes(g)], names=[f"pin{i}" for i in range(6)], seqs=[read] * 6,
                    quals=[b"I" * 100] * 6)
    return f


def mk_palindrome_pin():
    """a reverse-palindromic read hits the same Watson location through both
    strands; the dedupe key ignores strand, so it is reported once (UM)."""
    def f():
        names, gb = _genome(78, [100_000])
        g = bytearray(gb[0])
        half = bytes(x for x in g[3000:3050] if True)
        half = half.replace(b"C", b"A").replace(b"G", b"T")   # A/T only: immune to conversion
        pal = half + synth.revcomp_bytes(half)
        g[40000:40100] = pal
        return dict(gnames=names, gseqs=[bytes(g)], names=["pal0", "pal1"], seqs=[pal, pal], quals=[b"I" * 100] * 2)
    return f


CASES = [
    Case("se_cfg1", opts=dict(s=16, v=2, I=4, S=7), maker=mk_se(1, [200_000] * 5, 3000, 50, "cfg1")),
    Case("se_cfg2_r0_uR", opts=dict(s=16, v=5, I=4, r=0, S=3, u=1, R=1),
         maker=mk_se(2, [300_000] * 3, 2500, 100, "cfg2", repeats=40)),
    Case("se_cfg2_bsp", opts=dict(s=16, v=5, I=4, S=7), out_ext="bsp",
         maker=mk_se(3, [300_000] * 3, 2500, 100, "cfg2", repeats=40)),
    Case("se_cfg2_bsp_r0u", opts=dict(s=16, v=5, I=4, S=7, r=0, u=1), out_ext="bsp",
         maker=mk_se(3, [300_000] * 3, 1500, 100, "cfg2", repeats=40)),
    Case("se_mixed_A", opts=dict(s=16, v=4, I=4, S=11, u=1, A=[ADAPTER]), maker=mk_se_mixed(4, [250_000] * 4, 3000)),
    Case("se_mixed_fa", opts=dict(s=14, v=3, I=2, S=5, u=1, R=1), fasta_reads=True,
         maker=mk_se_mixed(5, [250_000] * 2, 1500, 14, 2)),
    Case("se_cfg5", opts=dict(s=12, v=15, I=4, w=1000, S=3), maker=mk_se(6, [400_000] * 2, 600, 144, "cfg5", repeats=1200)),
    Case("se_cfg5_w20_r0", opts=dict(s=12, v=15, I=4, w=20, r=0, S=3, u=1),
         maker=mk_se(6, [400_000] * 2, 600, 144, "cfg5", repeats=1200)),
    Case("se_n1", opts=dict(s=16, v=3, I=4, n=1, S=9), maker=mk_se(7, [300_000] * 2, 2000, 100, "cfg2")),
    Case("se_w3_I1", opts=dict(s=12, v=6, I=1, w=3, S=2), maker=mk_se(8, [300_000] * 2, 1500, 90, "cfg2", repeats=60, repeat_unit=200)),
    Case("se_L60", opts=dict(s=16, v=3, I=4, L=60, S=4), maker=mk_se(9, [300_000] * 2, 1500, 100, "cfg2")),
    Case("se_I16_s10", opts=dict(s=10, v=4, I=16, S=4), maker=mk_se(10, [150_000] * 2, 1000, 100, "cfg2")),
    Case("se_pin_order", opts=dict(s=16, v=5, I=4, S=7), maker=mk_selection_pin()),
    Case("se_pin_palindrome", opts=dict(s=16, v=5, I=4, S=7, n=1), out_ext="bsp", maker=mk_palindrome_pin()),
    Case("pe_sam", opts=dict(s=16, v=2, I=4, m=28, x=500, S=7, u=1), paired=True,
         maker=mk_pe(20, [300_000] * 3, 2000, 100, 150, 450, repeats=30)),
    Case("pe_sam_v5_R", opts=dict(s=16, v=5, I=4, m=28, x=500, S=5, R=1), paired=True,
         maker=mk_pe(21, [300_000] * 3, 1500, 100, 150, 450, subs="cfg2", repeats=30)),
    Case("pe_bsp_r0", opts=dict(s=16, v=3, I=4, m=28, x=500, S=7, r=0, u=1), paired=True, out_ext="bsp",
         # no QC (N-rich) mates here: BSP + -u prints a QC mate through an uninitialised Hit (pairs.cpp:254,
         # align.cpp:706) -- stack garbage decides whether its sequence is reverse-complemented
         maker=mk_pe(22, [300_000] * 3, 1500, 100, 150, 450, nrich=False, repeats=30)),
    Case("pe_readthrough", opts=dict(s=16, v=2, I=4, m=28, x=500, S=7), paired=True,
         maker=mk_pe(23, [300_000] * 2, 1500, 100, 60, 480)),
    Case("pe_n1", opts=dict(s=16, v=2, I=4, n=1, S=7), paired=True, maker=mk_pe(24, [300_000] * 2, 1000, 100, 150, 450)),
    Case("rrbs_se_A", opts=dict(D="C-CGG", v=2, S=5, A=[ADAPTER]), maker=mk_rrbs(30, [400_000] * 2, 2000, 75)),
    Case("rrbs_se_u_bsp", opts=dict(D="C-CGG", v=3, S=5, u=1, A=[ADAPTER]), out_ext="bsp", maker=mk_rrbs(31, [400_000] * 2, 1500, 75)),
    Case("rrbs_pe", opts=dict(D="C-CGG", v=2, S=5, A=[ADAPTER], u=1), paired=True, maker=mk_rrbs(32, [400_000] * 2, 1500, 75, paired=True)),
]
BY_NAME = {c.name: c for c in CASES}


def write_inputs(case: Case, tmpdir: str):
    d = case.data()
    fa = os.path.join(tmpdir, "ref.fa")
    synth.write_fasta(fa, [np.frombuffer(b, dtype=np.uint8) for b in d["gseqs"]], names=d["gnames"])
    ext = "fa" if case.fasta_reads else "fq"
    a = os.path.join(tmpdir, "a." + ext)
    b = os.path.join(tmpdir, "b." + ext) if case.paired else None
    if case.fasta_reads:
        synth.write_fasta_reads(a, d["seqs"], d["names"])
    else:
        synth.write_fastq(a, d["seqs"], d["names"], d["quals"])
        if b:
            synth.write_fastq(b, d["seqs_b"], d["names_b"], d["quals_b"])
    return fa, a, b


def input_digest(case: Case) -> str:
    d = case.data()
    h = hashlib.sha256()
    for k in ("gnames", "names", "names_b"):
        for x in d.get(k, ()):
            h.update(x.encode())
    for k in ("gseqs", "seqs", "quals", "seqs_b", "quals_b"):
        for x in d.get(k, ()):
            h.update(x)
    return h.hexdigest()
