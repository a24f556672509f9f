"""BAM read-input cases: (reference / product) command lines over BAM files built from the parity cases."""
from __future__ import annotations

import os

import bamio
import cases as CS

NAMES = ["bam_se", "bam_se_quirks_B11_E400", "bam_pe_interleaved", "bam_pe_odd_tail"]


def _mutate(i, seq, qual):
    """lower case, IUPAC codes, a foreign character, missing qualities, an over-long read: what the 4-bit round trip changes"""
    if i % 7 == 0:
        seq = seq.lower()
    if i % 11 == 0:
        seq = seq[:5] + "R" + seq[6:9] + "." + seq[10:]
    if i % 13 == 0:
        qual = None
    if i % 17 == 0:
        seq, qual = seq + seq[:60], (None if qual is None else qual + qual[:60])
    return seq, qual


def build(name, td):
    """-> (argv for bsmap without the program name, output path)"""
    if name.startswith("bam_se"):
        case = CS.BY_NAME["se_cfg2_r0_uR"]
        d = case.data()
        fa, _, _ = CS.write_inputs(case, td)
        recs = []
        for i, (n, s, q) in enumerate(zip(d["names"], d["seqs"], d["quals"])):
            s, q = s.decode(), q.decode()
            if "quirks" in name:
                s, q = _mutate(i, s, q)
            recs.append((n, 4, s, q))
        bam = os.path.join(td, "a.bam")
        bamio.write_bam(bam, recs)
        out = os.path.join(td, "out.sam")
        argv = case.cli(bam, None, fa, out)
        if "quirks" in name:
            argv += ["-B", "11", "-E", "400"]      # -B does not skip BAM records in the reference, it only offsets the index
        return argv, out
    case = CS.BY_NAME["pe_sam"]
    d = case.data()
    fa, _, _ = CS.write_inputs(case, td)
    recs = []
    for n, s, q, nb, sb, qb in zip(d["names"], d["seqs"], d["quals"], d["names_b"], d["seqs_b"], d["quals_b"]):
        recs.append((n, 0x4d, s.decode(), q.decode()))
        recs.append((nb, 0x8d, sb.decode(), qb.decode()))
    if name == "bam_pe_odd_tail":
        recs = recs[:-1]                              # the last pair loses its second mate
    bam = os.path.join(td, "pe.bam")
    bamio.write_bam(bam, recs)
    out = os.path.join(td, "out.sam")
    return case.cli(bam, bam, fa, out), out
