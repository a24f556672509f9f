"""Minimal writers for unaligned SAM / BAM read files (test inputs for the BAM/SAM ingest, reads.cpp:120-143).
BGZF = gzip members with the 'BC' extra subfield (block size), as samtools' bgzf.c requires."""
from __future__ import annotations

import struct
import zlib

NT16 = "=ACMGRSVTWYHKDBN"


def bgzf_block(data: bytes) -> bytes:
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25                      # total block size - 1
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, bsize)
    return head + comp + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def bgzf(data: bytes, block=60000) -> bytes:
    out = b"".join(bgzf_block(data[i:i + block]) for i in range(0, len(data), block))
    return out + bgzf_block(b"")                 # EOF marker


def bam_record(name: str, flag: int, seq: str, qual) -> bytes:
    """unmapped record; qual: str (phred+33) or None"""
    l = len(seq)
    codes = [NT16.index(ch.upper()) if ch.upper() in NT16 else 15 for ch in seq]
    if l & 1:
        codes.append(0)
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    q = bytes([0xff] * l) if qual is None else bytes(ord(c) - 33 for c in qual)
    qn = name.encode() + b"\0"
    core = struct.pack("<iiBBHHHiiii", -1, -1, len(qn), 0, 4680, 0, flag, l, -1, -1, 0)
    body = core + qn + packed + q
    return struct.pack("<i", len(body)) + body


def write_bam(path, records, header_text="@HD\tVN:1.0\tSO:unsorted\n"):
    """records: iterable of (name, flag, seq, qual)"""
    raw = b"BAM\1" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", 0)
    raw += b"".join(bam_record(*r) for r in records)
    open(path, "wb").write(bgzf(raw))


def write_sam(path, records):
    """headerless SAM text (the reference's CheckFile takes a leading '@' for FASTQ)"""
    with open(path, "w") as f:
        for name, flag, seq, qual in records:
            f.write("\t".join([name, str(flag), "*", "0", "0", "*", "*", "0", "0", seq or "*", qual if qual is not None else "*"]) + "\n")
