"""oracle/methratio_oracle.py -- CPU restatement of the reference's methratio.py (TEST INFRASTRUCTURE ONLY).

Only tests/ may import this module; the product path (bsmap_b200/csrc/bsx_meth.cu) never does.
Parity is PINNED: tests/test_methratio_cpu.py compares it with tests/golden/methratio/*.gz, the outputs of the
reference script itself (run through tests/golden/make_methratio_golden.py) on the golden alignment files.

Follows /root/reference/methratio.py line by line:
  get_alignment  (methratio.py:30-65)   filters, fill-in trimming, mate-overlap removal (SAM only)
  pileup         (methratio.py:95-118)  per reference C (Watson hits) / G (Crick hits): T/A -> depth, C/G -> meth+depth
  combine CpG    (methratio.py:122-131)
  report         (methratio.py:133-154) ratio and Wilson interval, %.3f
  -r             (methratio.py:52-55)   of the alignments that pass the filters, the first one in file order per
                                        (chromosome, fragment end, direction) counts; tested before trimming
"""
from __future__ import annotations

import numpy as np

FLAG_LETTERS = "pPuUrR12sfd"


def parse_alignments(path: str, chroms, unique=False, pair=False):
    """-> list of (seq, strand2, cr, pos, insert, mate_pos or None, sam_format)"""
    sam = path[-4:].upper() == ".SAM"
    out = []
    for line in open(path):
        if sam and line.startswith("@"):
            continue
        col = line.rstrip("\n").split("\t")
        if sam:
            f = int(col[1])
            if f & 0x4:
                continue
            if unique and f & 0x100:
                continue
            if pair and not f & 0x2:
                continue
            cr, pos, seq, insert = col[2], int(col[3]) - 1, col[9], int(col[8])
            if cr not in chroms:
                continue
            strand = ""
            for aux in col[11:]:
                if aux[:5] == "ZS:Z:":
                    strand = aux[5:7]
                    break
            if strand == "":
                raise ValueError(line)
            out.append((seq, strand, cr, pos, insert, int(col[7]) - 1, True))
        else:
            flag = col[3][:2]
            if flag in ("NM", "QC"):
                continue
            if unique and flag != "UM":
                continue
            if pair and col[7] == "0":
                continue
            seq, strand, cr, pos, insert = col[1], col[6], col[4], int(col[5]) - 1, int(col[7])
            if cr not in chroms:
                continue
            out.append((seq, strand, cr, pos, insert, None, False))
    return out


def trim(seq, strand, pos, insert, mate_pos, sam, trim_fillin):
    if trim_fillin > 0:
        if strand == "+-":
            seq = seq[:-trim_fillin]
        elif strand == "--":
            seq, pos = seq[trim_fillin:], pos + trim_fillin
        elif insert != 0 and len(seq) > abs(insert) - trim_fillin:
            trim_nt = len(seq) - (abs(insert) - trim_fillin)
            if strand == "++":
                seq = seq[:-trim_nt]
            elif strand == "-+":
                seq, pos = seq[trim_nt:], pos + trim_nt
    if sam and insert > 0:
        seq = seq[:mate_pos - pos]
    return seq, pos


def methratio(names, seqs, files, chroms=None, unique=False, pair=False, meth0=False, trim_fillin=2, combine_cpg=False, min_depth=1, rm_dup=False):
    """names/seqs: the reference FASTA records (bytes or str).  -> (table text, (nmap, nc, nd))"""
    ref = {}
    for n, s in zip(names, seqs):
        s = s.decode() if isinstance(s, bytes) else s
        if not chroms or n in chroms:
            ref[n] = s.upper()
    chroms = set(ref)
    meth = {c: np.zeros(len(s), dtype=np.int64) for c, s in ref.items()}
    depth = {c: np.zeros(len(s), dtype=np.int64) for c, s in ref.items()}
    refarr = {c: np.frombuffer(s.encode(), dtype=np.uint8) for c, s in ref.items()}
    coverage = {c: np.zeros(len(s), dtype=np.uint8) for c, s in ref.items()} if rm_dup else None
    nmap = 0
    for path in files:
        for seq, strand, cr, pos, insert, mate_pos, sam in parse_alignments(path, chroms, unique, pair):
            if rm_dup:                                       # methratio.py:52-55
                frag_end, direction = (pos + len(seq), 2) if strand in ("+-", "-+") else (pos, 1)
                if coverage[cr][frag_end] & direction:
                    continue
                coverage[cr][frag_end] |= direction
            seq, pos = trim(seq, strand, pos, insert, mate_pos, sam, trim_fillin)
            if pos + len(seq) > len(ref[cr]):
                continue
            nmap += 1
            if not seq:
                continue
            match, convert = (ord("C"), ord("T")) if strand[0] == "+" else (ord("G"), ord("A"))
            r = refarr[cr][pos:pos + len(seq)]
            q = np.frombuffer(seq.encode(), dtype=np.uint8)
            at = r == match
            np.add.at(depth[cr], pos + np.nonzero(at & ((q == convert) | (q == match)))[0], 1)
            np.add.at(meth[cr], pos + np.nonzero(at & (q == match))[0], 1)
    if combine_cpg:
        for cr in depth:
            r = refarr[cr]
            cg = np.nonzero((r[:-1] == ord("C")) & (r[1:] == ord("G")))[0]
            for a in (depth[cr], meth[cr]):
                a[cg] += a[cg + 1]
                a[cg + 1] = 0
    z95, z95sq = 1.96, 1.96 * 1.96
    lines = ["chr\tpos\tstrand\tcontext\tratio\ttotal_C\tmethy_C\tCI_lower\tCI_upper\n"]
    nc = nd = 0
    ss = {"C": "+", "G": "-"}
    for cr in sorted(depth):
        d_all, m_all, refcr = depth[cr], meth[cr], ref[cr]
        for i in np.nonzero(d_all >= min_depth)[0]:
            i = int(i)
            d, m = int(d_all[i]), int(m_all[i])
            nc += 1
            nd += d
            if m == 0 and not meth0:
                continue
            ratio = float(m) / d
            ctx = refcr[i - 2:i + 3]
            pmid = ratio + z95sq / (2 * d)
            sd = z95 * ((ratio * (1 - ratio) / d + z95sq / (4 * d * d)) ** 0.5)
            nm = 1 + z95sq / d
            lines.append("%s\t%d\t%c\t%s\t%.3f\t%d\t%d\t%.3f\t%.3f\n" % (cr, i + 1, ss[refcr[i]], ctx, ratio, d, m, (pmid - sd) / nm, (pmid + sd) / nm))
    return "".join(lines), (nmap, nc, nd)
