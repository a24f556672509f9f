"""Headline-scale parity (3.1 Gb genome, coordinates past 2^31), non-circular: the device-built index and the records
the CUDA path produces on it are compared with sha256 digests that were computed in the CPU container from the
ORACLE'S OWN index build (bso_ref_create) and mapping -- tests/golden/make_scale_golden.py -- and the SAM text with
the digest of the UNMODIFIED reference binary's output for the same 1 M reads.  Nothing is imported from the
device into the checker or the other way round.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import scale_cases as SC

pytestmark = pytest.mark.gpu


class _DevBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


@pytest.fixture(scope="module")
def golden():
    if not os.path.exists(SC.DIGESTS):
        pytest.skip("tests/golden/scale digests not generated")
    return json.load(open(SC.DIGESTS))


@pytest.fixture(scope="module")
def world(golden):
    import torch
    import bsmap_b200 as B
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, _ = torch.cuda.mem_get_info(0)
    if free < 60 << 30:
        pytest.skip("needs 60 GB of free HBM")
    dev = torch.device("cuda", 0)
    genome = SC.make_genome(dev)
    p = B.make_params(**SC.INDEX_OPTS)
    host = [torch.empty(ln, dtype=torch.uint8, pin_memory=True) for ln in SC.LENS]
    for h, g in zip(host, genome):
        h.copy_(g)
    torch.cuda.synchronize()
    ix = B.Index.from_pointers(p, SC.NAMES, [h.data_ptr() for h in host], SC.LENS, device=0)
    del host
    yield dict(torch=torch, B=B, dev=dev, genome=genome, ix=ix)
    ix.close()


def _digest_device_array(torch, dev, ptr, nbytes, chunk_bytes, pool):
    """sha256 of consecutive chunk_bytes slices of a device array: D2H through two pinned buffers, hashed on host threads"""
    view = torch.as_tensor(_DevBuf(ptr, nbytes), device=dev)
    futs = []
    for lo in range(0, nbytes, chunk_bytes):
        hi = min(nbytes, lo + chunk_bytes)
        futs.append(pool.submit(SC.sha, view[lo:hi].cpu().numpy()))
    return [f.result() for f in futs]


def test_index_equals_the_oracles_own_build(world, golden):
    torch, ix, dev = world["torch"], world["ix"], world["dev"]
    info = ix.info
    assert (info.n_words, info.n_keys, info.n_entries) == (golden["n_words"], golden["n_keys"], golden["n_entries"])
    exp = golden["arrays"]
    bufs = ix.device_buffers()      # refcat, crefcat, tab, pos, tag, ctx
    assert SC.sha(ix.download("anchor")) == exp["anchor"]
    with ThreadPoolExecutor(8) as pool:
        for name, k, per_entry in (("refcat", 0, 0), ("crefcat", 1, 0), ("tab", 2, 0), ("pos", 3, 4)):
            ptr, nbytes = bufs[k]
            assert ptr and nbytes, name
            if per_entry:
                got = _digest_device_array(torch, dev, ptr, nbytes, golden["chunk"] * per_entry, pool)
                bad = [i for i, (a, b) in enumerate(zip(got, exp[name])) if a != b]
                assert len(got) == len(exp[name]) and not bad, f"{name}: chunks {bad[:8]} of {len(got)} differ from the oracle's build"
            else:
                got = _digest_device_array(torch, dev, ptr, nbytes, nbytes, pool)[0]
                assert got == exp[name], f"{name} differs from the oracle's build"
        # the index is built with -v >= 8: 16-byte context entries {outer before, before, after, outer after}.  The oracle's
        # digests are of the inner pair ("ctx") and the outer pair ("ctx2") as separate arrays: split on the host.
        assert info.ctx_words == 4
        ptr, nbytes = bufs[5]
        view = torch.as_tensor(_DevBuf(ptr, nbytes), device=dev).view(torch.int32).view(-1, 4)
        step = golden["chunk"]
        futs = []
        for lo in range(0, view.shape[0], step):
            c = view[lo:lo + step].cpu().numpy().view(np.uint32)
            futs.append((pool.submit(SC.sha, np.ascontiguousarray(c[:, 1:3])), pool.submit(SC.sha, np.ascontiguousarray(c[:, [0, 3]]))))
        got_in, got_out = [a.result() for a, _ in futs], [b.result() for _, b in futs]
        assert got_in == exp["ctx"], "inner context differs from the oracle's build"
        assert got_out == exp["ctx2"], "outer context differs from the oracle's build"


def _map_se(world, w, want_counts=True):
    torch, B, ix = world["torch"], world["B"], world["ix"]
    p = B.make_params(**w["opts"])
    reads = SC.se_reads(world["genome"], w).cpu().numpy()
    lens = np.full(w["n"], w["L"], dtype=np.uint16)
    mp = B.Mapper(ix, p, max_batch=1 << 18, stride=w["stride"])
    recs, counts = mp.map_se(reads, lens, want_counts=want_counts)
    st = mp.stats()
    mp.close()
    return p, reads, recs, counts, st


def test_se_cfg2_records_equal_the_oracles(world, golden):
    w, exp = SC.SE, golden["workloads"][SC.SE["name"]]
    _, _, recs, counts, st = _map_se(world, w)
    assert st["candidates"] == exp["candidates"]
    assert int((recs["nhits"] > 0).sum()) == exp["mapped"]
    assert SC.sha(recs) == exp["recs"] and SC.sha(counts) == exp["counts"]


def test_se_cfg2_sam_equals_the_reference_binarys(world):
    """sorted SAM lines of the CUDA path == sorted SAM lines of oracle/_ref/bsmap -p 8 on the same genome and reads"""
    if not os.path.exists(SC.REFRUN):
        pytest.skip("reference-binary golden not generated")
    ref = json.load(open(SC.REFRUN))
    w = SC.SE
    p, reads, recs, counts, _ = _map_se(world, w)
    B, ix = world["B"], world["ix"]
    names = SC.se_read_names(world["genome"], w)
    seqs = [bytes(r[:w["L"]]) for r in reads]
    quals = [b"I" * w["L"]] * w["n"]
    mp = B.Mapper(ix, p, max_batch=1024, stride=w["stride"])
    txt, _ = mp.format_se(names, seqs, quals, recs, counts)
    mp.close()
    dig, nl = SC.sorted_sam_digest(txt)
    assert nl == ref["sam_lines"]
    assert dig == ref["sorted_sam_sha256"]


def test_pe_cfg3_records_equal_the_oracles(world, golden):
    B, ix = world["B"], world["ix"]
    w, exp = SC.PE, golden["workloads"][SC.PE["name"]]
    a, b = (x.cpu().numpy() for x in SC.pe_reads(world["genome"], w))
    lens = np.full(w["n"], w["L"], dtype=np.uint16)
    p = B.make_params(**w["opts"])
    mp = B.Mapper(ix, p, max_batch=1 << 17, stride=w["stride"])
    pr, ra, rb, _, _ = mp.map_pe(a, lens, b, lens)
    st = mp.stats()
    mp.close()
    assert st["candidates"] == exp["candidates"]
    assert int(pr["paired"].sum()) == exp["paired"]
    assert SC.sha(pr) == exp["pairs"] and SC.sha(ra) == exp["recs_a"] and SC.sha(rb) == exp["recs_b"]


def test_wide_context_kernel_records_equal_the_oracles(world, golden):
    w, exp = SC.WIDE, golden["workloads"][SC.WIDE["name"]]
    _, _, recs, counts, st = _map_se(world, w)
    assert st["candidates"] == exp["candidates"]
    assert SC.sha(recs) == exp["recs"] and SC.sha(counts) == exp["counts"]
