"""Headline-scale (3.1 Gb) parity workloads shared by the golden generator (CPU, oracle's OWN index build + the
unmodified reference binary) and the GPU tests.  TEST INFRASTRUCTURE.

Everything is derived from counter-based generators (bsmap_b200/synth.py), so the CPU container and the GPU box
produce bit-identical genomes and reads; what is committed under tests/golden/scale/ are sha256 digests only.

The genome and the SE read set are bench.py's config 2 (BASELINE.json configs[1]); the PE workload is
bench_configs' config 3; the `wide` workload drives bsx_map_se_wide_kernel (-v 12, 144 nt) on the same index.
All three sit on one s=16 / I=4 index whose coordinates run past 2^31 (25 x 124 Mb).
"""
from __future__ import annotations

import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

GENOME_SEED = 2
CHROMS = 25
CHROM_LEN = 124_000_000
NAMES = [f"chr{i + 1}" for i in range(CHROMS)]
LENS = [CHROM_LEN] * CHROMS
CHUNK = 1 << 26                      # entries per digest chunk of pos / ctx / ctx2

INDEX_OPTS = dict(s=16, I=4, v=12, S=7)      # built with -v >= 8 so that the wide context (ctx2) exists
SE = dict(name="se_cfg2", n=1_000_000, L=100, stride=104, seed=2024, subs="cfg2", opts=dict(s=16, v=5, I=4, S=7))
PE = dict(name="pe_cfg3", n=200_000, L=100, stride=104, seed=33, subs="cfg2", opts=dict(s=16, v=2, I=4, m=28, x=500, S=7, pairend=1))
WIDE = dict(name="se_wide", n=50_000, L=144, stride=144, seed=55, subs="cfg5", opts=dict(s=16, v=12, I=4, w=1000, S=7))

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale")
DIGESTS = os.path.join(GOLDEN_DIR, "cfg2_index_and_records.json")
REFRUN = os.path.join(GOLDEN_DIR, "cfg2_reference_binary.json")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data).hexdigest()


def sha_chunks(fetch, n, chunk=CHUNK, threads=8):
    """sha256 of consecutive `chunk`-entry slices; fetch(lo, hi) -> numpy array (hashlib releases the GIL)"""
    cuts = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda c: sha(fetch(*c)), cuts))


def make_genome(device="cpu"):
    from bsmap_b200 import synth
    return synth.make_genome(GENOME_SEED, LENS, device=device)


def se_reads(genome, w, lo=0, hi=None):
    """uint8[n, stride] ASCII read slots (torch tensor on the genome's device) of an SE workload, reads lo..hi"""
    import torch
    from bsmap_b200 import synth
    hi = w["n"] if hi is None else hi
    out = torch.zeros((hi - lo, w["stride"]), dtype=torch.uint8, device=genome[0].device)
    CH = 1 << 19
    for s0 in range(lo, hi, CH):
        m = min(CH, hi - s0)
        sim = synth.simulate_reads(genome, m, w["L"], seed=w["seed"], subs=w["subs"], first_index=s0)
        out[s0 - lo:s0 - lo + m, :w["L"]] = sim["seq"]
    return out


def se_read_names(genome, w, lo=0, hi=None):
    from bsmap_b200 import synth
    hi = w["n"] if hi is None else hi
    names = []
    CH = 1 << 19
    for s0 in range(lo, hi, CH):
        m = min(CH, hi - s0)
        sim = synth.simulate_reads(genome, m, w["L"], seed=w["seed"], subs=w["subs"], first_index=s0)
        names += synth.read_names(sim, first_index=s0)
    return names


def pe_reads(genome, w):
    import torch
    from bsmap_b200 import synth
    n = w["n"]
    a = torch.zeros((n, w["stride"]), dtype=torch.uint8, device=genome[0].device)
    b = torch.zeros_like(a)
    CH = 1 << 18
    for s0 in range(0, n, CH):
        m = min(CH, n - s0)
        sim = synth.simulate_pairs(genome, m, w["L"], seed=w["seed"], frag_min=150, frag_max=450, subs=w["subs"], first_index=s0)
        a[s0:s0 + m, :w["L"]] = sim["seq1"]; b[s0:s0 + m, :w["L"]] = sim["seq2"]
    return a, b


def inline_context(refcat, crefcat, tab, pos, lo, hi, seed_size, outward=0):
    """numpy restatement of the index's inline context for list entries [lo, hi): per entry the 16 reference bases
    that precede the seed and the 16 that follow it on the entry's strand (outward=16: the next 16 on either side).
    An entry is a reverse-strand entry iff it lies in the second half of its key's list (tab[2k+1] <= i < tab[2k+2])."""
    j0 = int(np.searchsorted(tab, lo, side="right")) - 1        # slot j holds entries tab[j] <= i < tab[j+1]; tab is non-decreasing
    j1 = int(np.searchsorted(tab, hi - 1, side="right")) - 1
    cnt = np.diff(np.clip(tab[j0:j1 + 2].astype(np.int64), lo, hi))
    strand = np.repeat((np.arange(j0, j1 + 1) & 1).astype(bool), cnt)
    c = pos[lo:hi].astype(np.int64)

    def window(start):
        j = start >> 4
        sh = ((start & 15) * 2).astype(np.uint64)
        out = np.empty(hi - lo, dtype=np.uint32)
        for m, sel in ((refcat, ~strand), (crefcat, strand)):
            jj = j[sel]
            w = (m[jj].astype(np.uint64) << np.uint64(32)) | m[jj + 1].astype(np.uint64)
            out[sel] = ((w >> (np.uint64(32) - sh[sel])) & np.uint64(0xffffffff)).astype(np.uint32)
        return out

    before = window(c - 16 - outward)
    after = window(c + seed_size + outward)
    return np.stack([before, after], axis=1)     # == uint2 {before, after}


def sorted_sam_digest(text: bytes):
    """sha256 over the sorted alignment lines (header lines dropped): thread-count independent"""
    lines = [ln for ln in text.split(b"\n") if ln and not ln.startswith(b"@")]
    lines.sort()
    h = hashlib.sha256()
    for ln in lines:
        h.update(ln); h.update(b"\n")
    return h.hexdigest(), len(lines)
