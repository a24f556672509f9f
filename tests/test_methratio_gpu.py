"""methratio on the device (bsx_meth.cu + the `methratio` command line) against the reference script's own outputs
(tests/golden/methratio/) and against the numpy oracle."""
from __future__ import annotations

import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases as CS
import runners as R
from test_methratio_cpu import GOLD, MANIFEST, alignment_files, opts_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import methratio_oracle as MO  # noqa: E402

import bsmap_b200 as B
from bsmap_b200 import lib as BL

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(BL.LIB_PATH), "methratio")


@pytest.mark.parametrize("key", sorted(MANIFEST))
def test_methratio_cli_writes_the_reference_table(tmp_path, key):
    """same command line as methratio.py, same bytes out, same summary line"""
    ent = MANIFEST[key]
    case = CS.BY_NAME[ent["case"]]
    fa, _, _ = CS.write_inputs(case, str(tmp_path))
    files = alignment_files(case, str(tmp_path), bam=ent.get("bam", False))     # .bam: decoded in place of `samtools view -X`
    out = str(tmp_path / "meth.txt")
    r = subprocess.run([EXE, "-o", out, "-d", fa, "-q"] + ent["opts"] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = open(out, "rb").read()
    exp = gzip.open(os.path.join(GOLD, key + ".txt.gz"), "rb").read()
    assert got == exp, R.first_diff(got, exp)
    assert r.stdout.strip() == ent["stdout"]


def test_methratio_cli_option_grammar(tmp_path):
    case = CS.BY_NAME["se_cfg1"]
    fa, _, _ = CS.write_inputs(case, str(tmp_path))
    files = alignment_files(case, str(tmp_path))
    exp = gzip.open(os.path.join(GOLD, "se_cfg1.g_z.txt.gz"), "rb").read()
    out = str(tmp_path / "m.txt")
    for argv in (["--out=" + out, "--ref", fa, "-gzq"], ["-o" + out, "-d" + fa, "--combine-CpG", "--zero-meth", "--quiet"]):
        r = subprocess.run([EXE] + argv + files, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == exp
    for argv, msg in ((["-d", fa] + files, "Missing output file"), (["-o", out] + files, "Missing reference file"),
                      (["-o", out, "-d", fa], "at least one"),
                      (["-o", out, "-d", fa, "-t", "x"] + files, "invalid integer"), (["-o", out, "-d", fa, "-Q"] + files, "no such option")):
        r = subprocess.run([EXE] + argv, capture_output=True, text=True)
        assert r.returncode == 2 and msg in r.stderr, (argv, r.stderr)
    (tmp_path / "x.bam").write_bytes(b"not a bam")
    r = subprocess.run([EXE, "-o", out, "-d", fa, str(tmp_path / "x.bam")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a BAM file" in r.stderr


@pytest.mark.parametrize("name,kw", [("pe_readthrough", dict(pair=True)), ("se_n1", dict(trim_fillin=7, combine_cpg=True)), ("pe_bsp_r0", dict(unique=True)),
                                     ("rrbs_se_A", dict(rm_dup=True)), ("pe_bsp_r0", dict(rm_dup=True, pair=True)), ("pe_sam", dict(rm_dup=True, trim_fillin=0))])
def test_meth_api_counters_equal_the_oracle(tmp_path, name, kw):
    """bsx_meth_add / bsx_meth_download through the C ABI: counters per position == the restatement's arrays"""
    case = CS.BY_NAME[name]
    d = case.data()
    files = alignment_files(case, str(tmp_path))
    names = d["gnames"]
    idx = {n: k for k, n in enumerate(names)}
    ix = B.Index.packed(names, d["gseqs"])
    mh = B.Meth(ix, B.meth_opts(**kw))
    seqs, chrs, pos, strand, ins, mate, flags = [], [], [], [], [], [], []
    for path in files:
        for seq, st, cr, p, insert, mate_pos, sam in MO.parse_alignments(path, set(names)):
            seqs.append(seq.encode()); chrs.append(idx[cr]); pos.append(p); ins.append(insert); mate.append(mate_pos if sam else 0)
            strand.append((1 if st[0] == "-" else 0) | (2 if st[1] == "-" else 0))
            flags.append(4 if sam else 0)     # filters are exercised through the command line; here everything is primary
    # flags for -u / -p need the raw columns: recompute them the way the parser does
    k = 0
    for path in files:
        sam = path.upper().endswith(".SAM")
        for line in open(path):
            col = line.rstrip("\n").split("\t")
            if sam:
                if line.startswith("@") or int(col[1]) & 4 or col[2] not in idx:
                    continue
                f = int(col[1]); flags[k] |= (1 if f & 0x100 else 0) | (2 if f & 0x2 else 0)
            else:
                if col[3][:2] in ("NM", "QC") or col[4] not in idx:
                    continue
                flags[k] |= (0 if col[3][:2] == "UM" else 1) | (0 if col[7] == "0" else 2)
            k += 1
    assert k == len(seqs)
    if kw.get("rm_dup"):                                        # several batches: a later batch loses against an earlier one
        cut = [0, len(seqs) // 3, len(seqs) // 3 + 1, len(seqs)]
        for b, e in zip(cut, cut[1:]):
            mh.add(seqs[b:e], chrs[b:e], pos[b:e], strand[b:e], ins[b:e], mate[b:e], flags[b:e])
    else:
        mh.add(seqs, chrs, pos, strand, ins, mate, flags)
    txt, (nmap, nc, nd) = MO.methratio(names, d["gseqs"], files, **kw)
    assert mh.n_valid == nmap
    out = tmp_path / "api.txt"
    fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    w, st = mh.write(d["gseqs"], fd)
    os.close(fd)
    assert out.read_bytes() == txt.encode() and st == (nc, nd) and w == len(txt)
    mh.close(); ix.close()


def test_packed_index_cannot_map():
    ix = B.Index.packed(["chr1"], [b"ACGT" * 1000])
    with pytest.raises(B.BsxError, match="no seed table"):
        B.Mapper(ix, B.make_params(), max_batch=16, stride=64)
    ix.close()


@pytest.mark.parametrize("key", sorted(k for k in MANIFEST if not MANIFEST[k].get("bam")))
def test_in_process_pileup_equals_the_reference_table(tmp_path, key):
    """reads -> Mapper with attached Meth (no SAM text in between) -> table == methratio.py on the reference's own output"""
    ent = MANIFEST[key]
    case = CS.BY_NAME[ent["case"]]
    d = case.data()
    kw = opts_kwargs(ent["opts"])
    chroms = kw.pop("chroms", None)
    p = B.make_params(**case.param_kwargs())
    ix = B.Index(p, d["gnames"], d["gseqs"])
    mp = B.Mapper(ix, p, max_batch=1024, stride=160)          # several sub-batches: the hook runs per batch
    mh = B.Meth(ix, B.meth_opts(**kw))
    if kw.get("rm_dup") and case.paired and not p.out_sam:     # two output files: the script's -r order is not the mapping order
        with pytest.raises(B.BsxError, match="order of two files"):
            mh.attach(mp, sam_rules=False)
        mh.close(); mp.close(); ix.close()
        return
    mh.attach(mp, sam_rules=bool(p.out_sam))
    if not case.paired:
        buf, lens = B.pack_reads(R.clip(case, d["seqs"]), stride=160)
        mp.map_se(buf, lens)
    else:
        ba, la = B.pack_reads(R.clip(case, d["seqs"]), stride=160)
        bb, lb = B.pack_reads(R.clip(case, d["seqs_b"]), stride=160)
        mp.map_pe(ba, la, bb, lb)
    out = tmp_path / "inproc.txt"
    fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    sel = None if chroms is None else [1 if n in chroms else 0 for n in d["gnames"]]
    w, (nc, nd) = mh.write(d["gseqs"], fd, chroms=sel)
    os.close(fd)
    exp = gzip.open(os.path.join(GOLD, key + ".txt.gz"), "rb").read()
    got = out.read_bytes()
    assert got == exp, R.first_diff(got, exp)
    if chroms is None:
        assert ent["stdout"] == "total %d valid mappings, %d covered cytosines, average coverage: %.2f fold." % (mh.valid(), nc, float(nd) / nc)
    mh.close(); mp.close(); ix.close()


@pytest.mark.parametrize("key", ["se_cfg2_r0_uR.default", "pe_sam.p_u", "pe_bsp_r0.default", "se_n1.t_5_g", "rrbs_se_A.r", "pe_sam_v5_R.r_p"])
def test_bsmap_cli_methratio_extension(tmp_path, key):
    """bsmap --methratio: the table of methratio.py without running it -- next to the alignment file, or instead of it"""
    ent = MANIFEST[key]
    case = CS.BY_NAME[ent["case"]]
    exe = os.path.join(os.path.dirname(BL.LIB_PATH), "bsmap")
    fa, a, b = CS.write_inputs(case, str(tmp_path))
    o = str(tmp_path / ("out." + case.out_ext))
    o2 = str(tmp_path / "out_unpair.bsp") if (case.paired and case.out_ext != "sam") else None
    flags = {"-u": ["--meth-unique"], "-p": ["--meth-pair"], "-z": ["--meth-zero"], "-g": ["--meth-cpg"], "-r": ["--meth-rmdup"]}
    ext, it = [], iter(ent["opts"])
    for x in it:
        ext += flags[x] if x in flags else [{"-t": "--meth-trim", "-m": "--meth-min-depth"}[x], next(it)]
    exp = gzip.open(os.path.join(GOLD, key + ".txt.gz"), "rb").read()
    exp_main, exp_un = R.golden_load(case)
    m1 = str(tmp_path / "m1.txt")
    r = subprocess.run([exe] + case.cli(a, b, fa, o, o2) + ["--methratio", m1] + ext, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(m1, "rb").read() == exp and open(o, "rb").read() == exp_main
    assert ent["stdout"] in r.stdout
    if case.out_ext == "sam":                                   # no alignment text at all
        m2 = str(tmp_path / "m2.txt")
        argv = [x for x in case.cli(a, b, fa, o, o2)]
        k = argv.index("-o"); del argv[k:k + 2]
        r = subprocess.run([exe] + argv + ["--methratio", m2] + ext, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert open(m2, "rb").read() == exp and ent["stdout"] in r.stdout
