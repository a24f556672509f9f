"""CPU-side checks of the product's host code: the C-ABI library loads and exports every symbol
include/bsmap_b200.h declares, compute entry points fail loudly without a GPU, and the host text
layer (bsx_format_*) turns records into exactly the reference's bytes.  No compute is done here:
the records fed to the formatter come from the oracle (the checker)."""
import os
import re

import numpy as np
import pytest

import cases as CS
import runners as R

import bsmap_b200 as B
from bsmap_b200 import lib as BL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from bsmap_b200 import build
    build.build()
    return BL.load()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "bsmap_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bsx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/bsmap_b200.h but not exported: {missing}"
    assert set(BL.EXPORTS) <= declared | {"bsx_cli_main"}


def test_struct_layouts_match_oracle():
    import oracle_lib as O
    import ctypes as C
    assert C.sizeof(BL.Params) == C.sizeof(O.Params) == 18 * 4 + 32 + 640
    assert BL.REC == O.REC and BL.REC.itemsize == 16
    assert BL.PAIR_REC == O.PAIR_REC and BL.PAIR_REC.itemsize == 28


@pytest.mark.skipif(BL.load().bsx_device_count() > 0 if os.path.exists(BL.LIB_PATH) else False, reason="has a GPU")
def test_no_cpu_fallback(L):
    p = B.make_params()
    with pytest.raises(B.BsxError, match="no CUDA device"):
        B.Index(p, ["chr1"], [b"ACGT" * 100])
    ix = B.Index.text_only(p, ["chr1"], [b"ACGT" * 100])
    with pytest.raises(B.BsxError, match="cannot map"):
        B.Mapper(ix, p, max_batch=16, stride=64)


def test_param_validation(L):
    ix = B.Index.text_only(B.make_params(), ["chr1"], [b"ACGT" * 100])
    for bad in (dict(s=7), dict(I=17), dict(v=16), dict(w=1001)):
        with pytest.raises(B.BsxError):
            B.Mapper(ix, B.make_params(**bad), max_batch=16, stride=64)


@pytest.mark.parametrize("case", CS.CASES, ids=lambda c: c.name)
def test_formatter_reproduces_reference_text(L, case):
    """oracle records -> product formatter == bytes written by the unmodified reference"""
    d = case.data()
    got = R.oracle_run(case)
    p = B.make_params(**case.param_kwargs())
    ix = B.Index.text_only(p, d["gnames"], d["gseqs"])
    fm = B.Mapper.__new__(B.Mapper)
    fm.index, fm.p, fm.h = ix, p, None
    head = ix.header() if p.out_sam else b""
    exp_main, exp_un = R.golden_load(case)
    if not case.paired:
        txt, na = fm.format_se(d["names"], d["seqs"], R.case_quals(case), got["recs"].astype(BL.REC), got["counts"])
        assert head + txt == exp_main, R.first_diff(head + txt, exp_main)
        assert na == got["n_aligned"]
    else:
        txt, un, st = fm.format_pe(d["names"], d["seqs"], d["quals"], d["names_b"], d["seqs_b"], d["quals_b"],
                                   got["pr"].astype(BL.PAIR_REC), got["ra"].astype(BL.REC), got["rb"].astype(BL.REC),
                                   got["ca"], got["cb"])
        assert head + txt == exp_main, R.first_diff(head + txt, exp_main)
        assert un == exp_un, R.first_diff(un, exp_un)
        assert st == got["n_aligned"]


def test_pack_reads_2bit_layout():
    """bsx_pack_reads: four bases per byte (first in bits 7:6, A0 C1 G2 T3, 0 otherwise), then one valid bit per base"""
    import bsmap_b200 as B
    reads = [b"ACGTNacgtn.X", b"T" * 37, b"", b"GATTACA" * 14]
    buf, lens = B.pack_reads(reads, stride=104)
    pk, low = B.pack_reads_2bit(buf, lens, threads=2)
    assert pk.shape == (4, 40) and low == 4
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    for r, rd in enumerate(reads):
        for i, ch in enumerate(rd[:104]):
            got = (pk[r, i >> 2] >> (6 - 2 * (i & 3))) & 3
            valid = (pk[r, 26 + (i >> 3)] >> (7 - (i & 7))) & 1
            assert valid == (ch in code) and got == code.get(ch, 0), (r, i, ch)
        tail = np.unpackbits(pk[r, 26:39])[len(rd):]
        assert not tail.any()
