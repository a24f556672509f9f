"""Deterministic synthetic genomes and simulated bisulfite reads.

Everything is counter-based (splitmix64 of a seed and a position), written with torch int64
ops only, so the same code gives bit-identical data on the CPU (tests, golden fixtures) and on a
GPU (bench.py generates the 3.1 Gb genome and 20 M reads in HBM in a few seconds).  torch is used
purely as an array library here; nothing in this file is on the mapping path.

Read simulation follows the probe recipe in SURVEY.md 8(d): uniform position, Watson/Crick 50/50
(Crick = reverse-complement the window first), every C -> T with p = conv, a few substitutions,
constant quality 'I', truth encoded in the read name.
"""
from __future__ import annotations

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD = _s64(0x9E3779B97F4A7C15)
_MUL1 = _s64(0xBF58476D1CE4E5B9)
_MUL2 = _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """logical shift right of an int64 tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _MUL1
    z = (z ^ _lsr(z, 27)) * _MUL2
    return z ^ _lsr(z, 31)


def _u(x: torch.Tensor, bits: int) -> torch.Tensor:
    """top `bits` bits of a hash as a non-negative int64"""
    return _lsr(x, 64 - bits)


_ACGT = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)
_COMP = torch.zeros(256, dtype=torch.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def genome_chunk(seed: int, chrom: int, start: int, n: int, device="cpu") -> torch.Tensor:
    """ASCII bases [start, start+n) of chromosome `chrom` of genome `seed` (uint8 tensor)."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    h = splitmix64(i ^ _s64((seed * 0x632BE59BD9B4E019) ^ (chrom << 40)))
    return _ACGT.to(device)[_u(h, 2)]


def make_genome(seed: int, chrom_lens, device="cpu", chunk=1 << 26):
    """list of uint8 tensors (ASCII ACGT), one per chromosome"""
    out = []
    for c, ln in enumerate(chrom_lens):
        g = torch.empty(ln, dtype=torch.uint8, device=device)
        for s in range(0, ln, chunk):
            n = min(chunk, ln - s)
            g[s:s + n] = genome_chunk(seed, c, s, n, device)
        out.append(g)
    return out


def plant_repeats(genome, seed: int, unit_len=300, copies=400, divergence=0.03):
    """Plant `copies` diverged copies of one random unit into chromosome 0.. (cfg5 stress)."""
    dev = genome[0].device
    unit = genome_chunk(seed ^ 0x5EED, 977, 0, unit_len, dev)
    total = sum(int(g.numel()) for g in genome)
    k = torch.arange(copies, dtype=torch.int64, device=dev)
    h = splitmix64(k * 7919 + _s64(seed * 0x1234567))
    lens = torch.tensor([int(g.numel()) for g in genome], dtype=torch.int64, device=dev)
    starts = torch.cumsum(lens, 0) - lens
    gpos = _u(h, 40) % total
    ch = torch.searchsorted(starts, gpos, right=True) - 1
    pos = gpos - starts[ch]
    pos = torch.minimum(pos, lens[ch] - unit_len - 1)
    for c in range(copies):
        j = torch.arange(unit_len, dtype=torch.int64, device=dev)
        hh = splitmix64(j + c * 100003 + _s64(seed * 0xABCDEF))
        mut = (_u(hh, 20).double() / (1 << 20)) < divergence
        nb = _ACGT.to(dev)[_u(splitmix64(hh), 2)]
        u = torch.where(mut, nb, unit)
        genome[int(ch[c])][int(pos[c]):int(pos[c]) + unit_len] = u
    return genome


_SUBS = {
    "cfg1": [0, 0, 0, 1, 1, 2],
    "cfg2": [0, 0, 0, 1, 1, 2, 3, 4],
    "cfg5": list(range(14)),
    "none": [0],
}


def simulate_reads(genome, n_reads: int, read_len: int, seed: int, subs="cfg2", conv=0.97,
                   first_index: int = 0, max_subs_slots: int = 16):
    """Simulate single-end directional bisulfite reads.

    Returns dict(seq=uint8[n, L] ASCII, chrom=int64[n], pos=int64[n] (0-based Watson start),
    strand=int64[n] (0 Watson / 1 Crick)).  Read r depends only on (seed, first_index + r).
    """
    dev = genome[0].device
    L = read_len
    lens = torch.tensor([int(g.numel()) for g in genome], dtype=torch.int64, device=dev)
    room = lens - L + 1
    cum = torch.cumsum(room, 0)
    total = int(cum[-1])
    r = torch.arange(first_index, first_index + n_reads, dtype=torch.int64, device=dev)
    base = splitmix64(r * 0x100000 + _s64(seed * 0x2545F4914F6CDD1D))
    h_pos = splitmix64(base + 1)
    h_str = splitmix64(base + 2)
    h_ns = splitmix64(base + 3)
    gpos = _u(h_pos, 48) % total
    ch = torch.searchsorted(cum, gpos, right=True)
    pos = gpos - (cum[ch] - room[ch])
    strand = _u(h_str, 1)
    seq = torch.empty((n_reads, L), dtype=torch.uint8, device=dev)
    j = torch.arange(L, dtype=torch.int64, device=dev)
    for c in range(len(genome)):
        m = (ch == c).nonzero().squeeze(1)
        if m.numel() == 0:
            continue
        idx = pos[m, None] + j[None, :]
        seq[m] = genome[c][idx]
    # Crick reads: reverse complement of the window
    rc = _COMP.to(dev)[seq.long()].flip(1)
    seq = torch.where(strand[:, None] == 1, rc, seq)
    # bisulfite conversion C->T with prob conv (per base)
    hb = splitmix64(base[:, None] * 0x3 + j[None, :] + 0x1000)
    keep = (_u(hb, 20).double() / (1 << 20)) >= conv
    isC = seq == 67
    seq = torch.where(isC & ~keep, torch.full_like(seq, 84), seq)
    # substitutions
    tab = torch.tensor(_SUBS[subs], dtype=torch.int64, device=dev)
    nsub = tab[_u(h_ns, 32) % len(_SUBS[subs])]
    for s in range(min(max_subs_slots, max(_SUBS[subs]))):
        hs = splitmix64(base + 0x40 + s)
        p = _u(hs, 32) % L
        nb = _ACGT.to(dev)[_u(splitmix64(hs), 2)]
        do = nsub > s
        rows = do.nonzero().squeeze(1)
        seq[rows, p[rows]] = nb[rows]
    return dict(seq=seq, chrom=ch, pos=pos, strand=strand)


def simulate_pairs(genome, n_pairs: int, read_len: int, seed: int, frag_min=150, frag_max=450,
                   subs="cfg2", conv=0.97, first_index: int = 0):
    """Paired-end directional bisulfite reads: mate 1 = first L nt of the converted fragment,
    mate 2 = first L nt of the reverse complement of the converted fragment."""
    dev = genome[0].device
    L = read_len
    lens = torch.tensor([int(g.numel()) for g in genome], dtype=torch.int64, device=dev)
    r = torch.arange(first_index, first_index + n_pairs, dtype=torch.int64, device=dev)
    base = splitmix64(r * 0x100000 + _s64(seed * 0x2545F4914F6CDD1D) + 7)
    flen = frag_min + _u(splitmix64(base + 5), 32) % (frag_max - frag_min + 1)
    room = lens - frag_max + 1
    cum = torch.cumsum(room, 0)
    total = int(cum[-1])
    gpos = _u(splitmix64(base + 1), 48) % total
    ch = torch.searchsorted(cum, gpos, right=True)
    pos = gpos - (cum[ch] - room[ch])
    strand = _u(splitmix64(base + 2), 1)
    F = frag_max
    j = torch.arange(F, dtype=torch.int64, device=dev)
    frag = torch.empty((n_pairs, F), dtype=torch.uint8, device=dev)
    for c in range(len(genome)):
        m = (ch == c).nonzero().squeeze(1)
        if m.numel() == 0:
            continue
        idx = torch.minimum(pos[m, None] + j[None, :], lens[c] - 1)
        frag[m] = genome[c][idx]
    valid = j[None, :] < flen[:, None]
    # Crick: reverse complement within the fragment length
    ridx = (flen[:, None] - 1 - j[None, :]).clamp(min=0)
    rcfrag = _COMP.to(dev)[torch.gather(frag, 1, ridx).long()]
    frag = torch.where(strand[:, None] == 1, rcfrag, frag)
    hb = splitmix64(base[:, None] * 0x3 + j[None, :] + 0x1000)
    keep = (_u(hb, 20).double() / (1 << 20)) >= conv
    frag = torch.where((frag == 67) & ~keep, torch.full_like(frag, 84), frag)
    m1 = frag[:, :L].clone()
    # mate 2: first L of reverse complement of converted fragment
    jj = torch.arange(L, dtype=torch.int64, device=dev)
    idx2 = (flen[:, None] - 1 - jj[None, :]).clamp(min=0)
    m2 = _COMP.to(dev)[torch.gather(frag, 1, idx2).long()]
    if frag_min < L:  # read-through beyond the fragment: fill with a fixed adapter-ish tail
        tail = _ACGT.to(dev)[(jj % 4)]
        short = jj[None, :] >= flen[:, None]
        m1 = torch.where(short, tail[None, :].expand_as(m1), m1)
        m2 = torch.where(short, tail[None, :].expand_as(m2), m2)
    tab = torch.tensor(_SUBS[subs], dtype=torch.int64, device=dev)
    for mate, arr in ((0, m1), (1, m2)):
        nsub = tab[_u(splitmix64(base + 3 + 16 * mate), 32) % len(_SUBS[subs])]
        for s in range(max(_SUBS[subs])):
            hs = splitmix64(base + 0x40 + s + 64 * mate)
            p = _u(hs, 32) % L
            nb = _ACGT.to(dev)[_u(splitmix64(hs), 2)]
            rows = (nsub > s).nonzero().squeeze(1)
            arr[rows, p[rows]] = nb[rows]
    return dict(seq1=m1, seq2=m2, chrom=ch, pos=pos, strand=strand, flen=flen)


def rrbs_fragments(genome, site=b"CCGG", digest_pos=1, frag_min=40, frag_max=400):
    """Digestion fragments of every chromosome between adjacent sites (torch, on the genome's device):
    int64 tensors (chrom, start, end) of the fragments whose length lies in [frag_min, frag_max]."""
    dev = genome[0].device
    st = torch.tensor(list(site), dtype=torch.uint8, device=dev)
    cs, a_, e_ = [], [], []
    for c, g in enumerate(genome):
        n = int(g.numel())
        hit = torch.ones(n - len(site) + 1, dtype=torch.bool, device=dev)
        for t in range(len(site)):
            hit &= g[t:n - len(site) + 1 + t] == st[t]
        sites = hit.nonzero().squeeze(1) + digest_pos
        a, e = sites[:-1], sites[1:] + len(site) - 2 * digest_pos
        ok = ((e - a) >= frag_min) & ((e - a) <= frag_max)
        a_.append(a[ok]); e_.append(e[ok]); cs.append(torch.full((int(ok.sum()),), c, dtype=torch.int64, device=dev))
    return torch.cat(cs), torch.cat(a_), torch.cat(e_)


def simulate_rrbs_reads(genome, frags, n_reads: int, read_len: int, seed: int, adapter: bytes, conv=0.97, nsub_max=2, first_index: int = 0):
    """RRBS single-end reads on the device: a random digestion fragment, Watson or Crick (reverse complement), bisulfite
    converted, read through into `adapter` and then A's when the fragment is shorter than the read; up to nsub_max
    substitutions.  Read r depends only on (seed, first_index + r).  -> uint8[n, L]"""
    dev = genome[0].device
    L = read_len
    fc, fa, fe = frags
    lens = torch.tensor([int(g.numel()) for g in genome], dtype=torch.int64, device=dev)
    off = torch.cumsum(lens, 0) - lens
    flat = torch.cat(list(genome))
    r = torch.arange(first_index, first_index + n_reads, dtype=torch.int64, device=dev)
    base = splitmix64(r * 0x100000 + _s64(seed * 0x2545F4914F6CDD1D) + 11)
    k = _u(splitmix64(base + 1), 48) % int(fc.numel())
    strand = _u(splitmix64(base + 2), 1)
    a, e = off[fc[k]] + fa[k], off[fc[k]] + fe[k]
    flen = e - a
    j = torch.arange(L, dtype=torch.int64, device=dev)
    inside = j[None, :] < flen[:, None]
    idx = torch.where(strand[:, None] == 0, a[:, None] + j[None, :], e[:, None] - 1 - j[None, :])
    idx = torch.where(inside, idx, torch.zeros_like(idx))
    seq = flat[idx]
    seq = torch.where(strand[:, None] == 1, _COMP.to(dev)[seq.long()], seq)
    hb = splitmix64(base[:, None] * 0x3 + j[None, :] + 0x1000)
    keep = (_u(hb, 20).double() / (1 << 20)) >= conv
    seq = torch.where((seq == 67) & ~keep, torch.full_like(seq, 84), seq)
    ad = torch.full((L + len(adapter) + 1,), 65, dtype=torch.uint8, device=dev)
    ad[:len(adapter)] = torch.tensor(list(adapter), dtype=torch.uint8, device=dev)
    tail = ad[(j[None, :] - flen[:, None]).clamp(min=0, max=L + len(adapter))]
    seq = torch.where(inside, seq, tail)
    nsub = _u(splitmix64(base + 3), 32) % (nsub_max + 1)
    for s_ in range(nsub_max):
        hs = splitmix64(base + 0x40 + s_)
        pp = _u(hs, 32) % L
        nb = _ACGT.to(dev)[_u(splitmix64(hs), 2)]
        rows = (nsub > s_).nonzero().squeeze(1)
        seq[rows, pp[rows]] = nb[rows]
    return seq


# ---------------------------------------------------------------------------------------------
# file writers (numpy; used to feed the reference binary and the bsmap CLI)
# ---------------------------------------------------------------------------------------------

def write_fasta(path: str, genome, names=None, width: int = 60):
    with open(path, "wb") as f:
        for c, g in enumerate(genome):
            a = g.cpu().numpy() if isinstance(g, torch.Tensor) else np.asarray(g, dtype=np.uint8)
            name = names[c] if names else f"chr{c + 1}"
            f.write(b">" + name.encode() + b"\n")
            n = a.size
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = a[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(a[full:].tobytes() + b"\n")


def read_names(sim, first_index=0, suffix=""):
    ch = sim["chrom"].cpu().numpy()
    pos = sim["pos"].cpu().numpy()
    st = sim["strand"].cpu().numpy()
    return [f"r{first_index + i}_{ch[i] + 1}_{pos[i] + 1}_{'+-'[st[i]]}{suffix}" for i in range(len(ch))]


def write_fastq(path: str, seqs, names, quals=None):
    """seqs: uint8[n, L] array or list of bytes; constant quality 'I' when quals is None"""
    with open(path, "wb") as f:
        if isinstance(seqs, torch.Tensor):
            seqs = seqs.cpu().numpy()
        out = []
        for i, nm in enumerate(names):
            s = seqs[i].tobytes() if not isinstance(seqs[i], (bytes, bytearray)) else bytes(seqs[i])
            q = quals[i] if quals is not None else b"I" * len(s)
            out.append(b"@" + nm.encode() + b"\n" + s + b"\n+\n" + q + b"\n")
            if len(out) >= 65536:
                f.write(b"".join(out))
                out = []
        f.write(b"".join(out))


def write_fasta_reads(path: str, seqs, names):
    with open(path, "wb") as f:
        if isinstance(seqs, torch.Tensor):
            seqs = seqs.cpu().numpy()
        out = []
        for i, nm in enumerate(names):
            s = seqs[i].tobytes() if not isinstance(seqs[i], (bytes, bytearray)) else bytes(seqs[i])
            out.append(b">" + nm.encode() + b"\n" + s + b"\n")
        f.write(b"".join(out))


# ---------------------------------------------------------------------------------------------
# small-scale extras for parity cases (numpy, CPU only)
# ---------------------------------------------------------------------------------------------

def revcomp_bytes(b: bytes) -> bytes:
    return bytes(_COMP.numpy()[np.frombuffer(b, dtype=np.uint8)][::-1])


def bisulfite_bytes(b: bytes, rng: np.random.Generator, conv=0.97) -> bytes:
    a = np.frombuffer(b, dtype=np.uint8).copy()
    c = (a == 67) & (rng.random(a.size) < conv)
    a[c] = 84
    return a.tobytes()


def simulate_rrbs(genome_bytes, n_reads, read_len, seed, site=b"CCGG", digest_pos=1,
                  frag_min=40, frag_max=400, adapter=b"AGATCGGAAGAGCGGTTCAGCAGGAATGCCGAGA", nsub_max=2,
                  paired=False):
    """RRBS reads: fragments between adjacent digestion sites; read = first read_len nt of the
    (converted) fragment or of its reverse complement, read-through continues into `adapter`."""
    rng = np.random.default_rng(seed)
    frags = []
    for c, g in enumerate(genome_bytes):
        up = g.upper()
        sites, q = [], up.find(site)
        while q >= 0:
            sites.append(q + digest_pos)
            q = up.find(site, q + 1)
        for a, b in zip(sites[:-1], sites[1:]):
            e = b + len(site) - 2 * digest_pos
            if frag_min <= e - a <= frag_max:
                frags.append((c, a, e))
    out, out2, names = [], [], []
    for i in range(n_reads):
        c, a, e = frags[int(rng.integers(len(frags)))]
        frag = genome_bytes[c][a:e].upper()
        strand = int(rng.integers(2))
        if strand:
            frag = revcomp_bytes(frag)
        frag = bisulfite_bytes(frag, rng)
        r1 = (frag + adapter + b"A" * read_len)[:read_len]
        r2 = (revcomp_bytes(frag) + revcomp_bytes(adapter)[:0] + adapter + b"A" * read_len)[:read_len]
        def mut(r):
            r = bytearray(r)
            for _ in range(int(rng.integers(nsub_max + 1))):
                r[int(rng.integers(len(r)))] = b"ACGT"[int(rng.integers(4))]
            return bytes(r)
        out.append(mut(r1)); out2.append(mut(r2))
        names.append(f"rr{i}_{c + 1}_{a + 1}_{e - a}_{'+-'[strand]}")
    return (out, out2, names) if paired else (out, names)
