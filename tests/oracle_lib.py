"""ctypes binding of oracle/libbsmap_oracle.so -- TEST INFRASTRUCTURE (the checker, never the product).

Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libbsmap_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "bsmap")


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "seed_size", "index_interval", "max_snp_num", "max_num_hits", "report_repeat_hits",
        "min_insert", "max_insert", "chains", "pairend", "rrbs", "randseed", "max_ns",
        "max_readlen", "out_sam", "out_unmap", "out_ref", "digest_pos", "n_adapter")] + [
        ("digest_site", C.c_char * 32), ("adapter", (C.c_char * 64) * 10)]


REC = np.dtype([("loc", "<u4"), ("chr", "<u4"), ("nhits", "<u4"), ("nm", "u1"), ("chain", "u1"),
                ("status", "u1"), ("len", "u1")])
PAIR_REC = np.dtype([("a_loc", "<u4"), ("a_chr", "<u4"), ("b_loc", "<u4"), ("b_chr", "<u4"),
                     ("insert", "<i4"), ("npairs", "<u4"), ("na", "u1"), ("nb", "u1"),
                     ("chain", "u1"), ("paired", "u1")])


def make_params(s=16, I=4, v=2, w=1000, r=1, m=28, x=500, n=0, pairend=0, S=0, f=5, L=144,
                out_sam=1, u=0, R=0, D=None, A=()):
    """mirror of Param defaults (param.cpp:6-83) + mGetOptions side effects (main.cpp:234-289)"""
    p = Params()
    p.seed_size, p.index_interval, p.max_snp_num, p.max_num_hits = s, I, v, w
    p.report_repeat_hits, p.min_insert, p.max_insert, p.chains = r, m, x, n
    p.pairend, p.randseed, p.max_ns, p.max_readlen = pairend, S, f, L
    p.out_sam, p.out_unmap, p.out_ref = out_sam, u, R
    if D:
        pos = D.index("-")
        p.digest_site = D.replace("-", "").encode()
        p.digest_pos = pos
        p.rrbs, p.index_interval, p.seed_size = 1, 1, 12   # SetDigestionSite, param.cpp:95-106
    p.n_adapter = len(A)
    for i, a in enumerate(A):
        p.adapter[i].value = a.encode()
    return p


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "liboracle"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
                os.path.join(ORACLE_DIR, "bsmap_oracle.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.bso_ref_create.restype = C.c_void_p
        L.bso_ref_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p]
        L.bso_ref_import.restype = C.c_void_p
        L.bso_ref_import.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_uint64]
        L.bso_ref_destroy.argtypes = [C.c_void_p]
        for f in ("n_words", "n_keys", "n_entries"):
            getattr(L, "bso_ref_" + f).restype = C.c_uint64
            getattr(L, "bso_ref_" + f).argtypes = [C.c_void_p]
        for f in ("refcat", "crefcat", "anchor", "tab", "pos", "pos_tag"):
            getattr(L, "bso_ref_" + f).restype = C.POINTER(C.c_uint32)
            getattr(L, "bso_ref_" + f).argtypes = [C.c_void_p]
        L.bso_map_se.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                 C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bso_map_pe.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                 C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.bso_format_header.restype = C.c_size_t
        L.bso_format_header.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.bso_format_se.restype = C.c_size_t
        L.bso_format_se.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                    C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t,
                                    C.POINTER(C.c_uint32)]
        L.bso_format_pe.restype = C.c_size_t
        L.bso_format_pe.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32] + [C.POINTER(C.c_char_p)] * 6 + [
            C.c_void_p] * 5 + [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.bso_xt.restype = C.c_uint32; L.bso_xt.argtypes = [C.c_uint32]
        L.bso_pack16.restype = C.c_uint32; L.bso_pack16.argtypes = [C.c_char_p]
        L.bso_mismatch_cell.restype = C.c_uint32; L.bso_mismatch_cell.argtypes = [C.c_uint32, C.c_uint32]
        L.bso_myrand.restype = C.c_uint32; L.bso_myrand.argtypes = [C.c_int32, C.c_int32]
        L.bso_profile_a.restype = C.c_int; L.bso_profile_a.argtypes = [C.c_int] * 4
        _lib = L
    return _lib


def _strs(xs):
    arr = (C.c_char_p * len(xs))()
    arr[:] = [x if isinstance(x, bytes) else x.encode() for x in xs]
    return arr


def pack_reads(seqs, stride=None):
    """list of bytes -> (uint8[n, stride] zero padded, uint16 lens)"""
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.uint16)
    stride = stride or max(160, int(lens.max()) if n else 160)
    buf = np.zeros((n, stride), dtype=np.uint8)
    for i, s in enumerate(seqs):
        buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    return buf, lens


class OracleRef:
    def __init__(self, params: Params, names, seqs):
        """names: list[str]; seqs: list[bytes | np.uint8 array]"""
        self.p = params
        self._seqs = [s.tobytes() if isinstance(s, np.ndarray) else bytes(s) for s in seqs]
        lens = np.array([len(s) for s in self._seqs], dtype=np.uint32)
        self.h = lib().bso_ref_create(C.byref(params), len(names), _strs(names), _strs(self._seqs),
                                      lens.ctypes.data)
        self.n_seq = len(names)

    @classmethod
    def imported(cls, params, names, lens, refcat, crefcat, tab, pos):
        """adopt index arrays built elsewhere (numpy uint32 arrays, kept alive by this object)"""
        self = cls.__new__(cls)
        self.p, self.n_seq = params, len(names)
        self._keep = (refcat, crefcat, tab, pos)
        ln = np.asarray(lens, dtype=np.uint32)
        self.h = lib().bso_ref_import(C.byref(params), len(names), _strs(names), ln.ctypes.data, refcat.ctypes.data,
                                      crefcat.ctypes.data, tab.ctypes.data, pos.ctypes.data, len(pos))
        if not self.h:
            raise RuntimeError("bso_ref_import failed")
        return self

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.bso_ref_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, name, n):
        return np.ctypeslib.as_array(getattr(lib(), "bso_ref_" + name)(self.h), shape=(int(n),))

    @property
    def n_words(self): return int(lib().bso_ref_n_words(self.h))
    @property
    def n_keys(self): return int(lib().bso_ref_n_keys(self.h))
    @property
    def n_entries(self): return int(lib().bso_ref_n_entries(self.h))
    @property
    def refcat(self): return self._arr("refcat", self.n_words)
    @property
    def crefcat(self): return self._arr("crefcat", self.n_words)
    @property
    def anchor(self): return self._arr("anchor", self.n_seq + 1)
    @property
    def tab(self): return self._arr("tab", 2 * self.n_keys + 1)
    @property
    def pos(self): return self._arr("pos", self.n_entries)
    @property
    def pos_tag(self): return self._arr("pos_tag", self.n_entries)

    def map_se(self, buf, lens, first_index=0, readset=0, want_counts=True, params=None):
        p = params or self.p
        n = len(lens)
        out = np.zeros(n, dtype=REC)
        counts = np.zeros((n, 16), dtype=np.uint16) if want_counts else None
        stats = np.zeros(4, dtype=np.uint64)
        lib().bso_map_se(self.h, C.byref(p), n, buf.ctypes.data, buf.shape[1], lens.ctypes.data, first_index,
                         readset, out.ctypes.data, counts.ctypes.data if want_counts else None, stats.ctypes.data)
        return out, counts, stats

    def map_pe(self, buf_a, lens_a, buf_b, lens_b, first_index=0, params=None):
        p = params or self.p
        n = len(lens_a)
        out = np.zeros(n, dtype=PAIR_REC)
        ra, rb = np.zeros(n, dtype=REC), np.zeros(n, dtype=REC)
        ca, cb = np.zeros((n, 16), dtype=np.uint16), np.zeros((n, 16), dtype=np.uint16)
        stats = np.zeros(4, dtype=np.uint64)
        assert buf_a.shape[1] == buf_b.shape[1]
        lib().bso_map_pe(self.h, C.byref(p), n, buf_a.ctypes.data, buf_b.ctypes.data, buf_a.shape[1],
                         lens_a.ctypes.data, lens_b.ctypes.data, first_index, out.ctypes.data, ra.ctypes.data,
                         rb.ctypes.data, ca.ctypes.data, cb.ctypes.data, stats.ctypes.data)
        return out, ra, rb, ca, cb, stats

    def header(self):
        cap = 1 << 20
        b = C.create_string_buffer(cap)
        n = lib().bso_format_header(self.h, b, cap)
        return b.raw[:n]

    def format_se(self, names, seqs, quals, recs, counts, readset=0, params=None):
        p = params or self.p
        n = len(names)
        cap = 1024 * n + 4096
        b = C.create_string_buffer(cap)
        na = C.c_uint32(0)
        ln = lib().bso_format_se(self.h, C.byref(p), n, _strs(names), _strs(seqs), _strs(quals), readset,
                                 recs.ctypes.data, counts.ctypes.data if counts is not None else None, b, cap,
                                 C.byref(na))
        assert ln < cap
        return b.raw[:ln], na.value

    def format_pe(self, names_a, seqs_a, quals_a, names_b, seqs_b, quals_b, pr, ra, rb, ca, cb, params=None):
        p = params or self.p
        n = len(names_a)
        cap = 2048 * n + 4096
        b, bu = C.create_string_buffer(cap), C.create_string_buffer(cap)
        nu = C.c_size_t(0)
        st = (C.c_uint32 * 3)()
        ln = lib().bso_format_pe(self.h, C.byref(p), n, _strs(names_a), _strs(seqs_a), _strs(quals_a),
                                 _strs(names_b), _strs(seqs_b), _strs(quals_b), pr.ctypes.data, ra.ctypes.data,
                                 rb.ctypes.data, ca.ctypes.data, cb.ctypes.data, b, cap, bu, cap, C.byref(nu), st)
        assert ln < cap and nu.value < cap
        return b.raw[:ln], bu.raw[:nu.value], tuple(st)


def run_reference(args, cwd=None, timeout=3600):
    """run the unmodified reference binary (oracle/_ref/bsmap); returns stdout"""
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(REF_BIN)
    r = subprocess.run([REF_BIN] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference bsmap failed ({r.returncode}): {r.stderr[-2000:]}")
    return r.stdout
