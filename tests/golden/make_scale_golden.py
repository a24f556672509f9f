#!/usr/bin/env python
"""Generate tests/golden/scale/*.json -- headline-scale (3.1 Gb) goldens.  Run in the CPU container (needs
/root/reference for the unmodified binary; ~62 GB RAM, ~25 min on 8 cores):

    python tests/golden/make_scale_golden.py [--skip-reference] [--skip-oracle]

1. cfg2_index_and_records.json: the ORACLE'S OWN index build (bso_ref_create, the plain-C restatement of
   dbseq.cpp:215-523) of bench.py's 3.1 Gb genome: sha256 of refcat / crefcat / anchor / tab and of pos (and of the
   inline context arrays, restated in numpy from refcat + pos) in 64 M-entry chunks; then the oracle maps three
   workloads on that index (SE config 2: 1 M reads; PE config 3: 200 k pairs; wide-context: 50 k 144-nt reads at
   -v 12) and the digests of its records, per-level counts and the candidate counters are stored.
   The GPU test (tests/test_scale_gpu.py) builds the same genome on the device and must reproduce every digest,
   so neither side imports anything from the other.
2. cfg2_reference_binary.json: the UNMODIFIED reference binary (oracle/_ref/bsmap) on the same genome and the same
   1 M SE reads, `-s 16 -v 5 -I 4 -S 7 -p <nproc>`: sha256 of the sorted SAM lines and the wall
   clock of each phase (FASTA load, seed table, mapping); the oracle's SAM for the same reads is compared with it
   here, which pins the oracle at this scale as well.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (TESTS, ROOT):
    sys.path.insert(0, p)

import oracle_lib as O          # noqa: E402
import scale_cases as SC        # noqa: E402
from bsmap_b200 import synth    # noqa: E402


def log(*a):
    print(time.strftime("%H:%M:%S"), *a, flush=True)


def par_map_se(oref, p, buf, lens, threads):
    from concurrent.futures import ThreadPoolExecutor
    n = len(lens)
    cuts = np.linspace(0, n, 4 * threads + 1).astype(int)
    recs = np.zeros(n, dtype=O.REC); counts = np.zeros((n, 16), dtype=np.uint16); stats = []

    def work(i):
        a, b = int(cuts[i]), int(cuts[i + 1])
        r, c, st = oref.map_se(buf[a:b], lens[a:b], first_index=a, params=p)
        recs[a:b] = r; counts[a:b] = c; stats.append(st)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(4 * threads)))
    return recs, counts, np.sum(stats, axis=0)


def par_map_pe(oref, p, ba, bb, lens, threads):
    from concurrent.futures import ThreadPoolExecutor
    n = len(lens)
    cuts = np.linspace(0, n, 4 * threads + 1).astype(int)
    pr = np.zeros(n, dtype=O.PAIR_REC); ra = np.zeros(n, dtype=O.REC); rb = np.zeros(n, dtype=O.REC); stats = []

    def work(i):
        a, b = int(cuts[i]), int(cuts[i + 1])
        o = oref.map_pe(ba[a:b], lens[a:b], bb[a:b], lens[a:b], first_index=a, params=p)
        pr[a:b] = o[0]; ra[a:b] = o[1]; rb[a:b] = o[2]; stats.append(o[5])
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(4 * threads)))
    return pr, ra, rb, np.sum(stats, axis=0)


def timed_reference(args, cwd):
    """run oracle/_ref/bsmap, stamping every stdout line with the wall clock -> (lines, total seconds)"""
    t0 = time.perf_counter()
    p = subprocess.Popen([O.REF_BIN] + [str(a) for a in args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    lines = []
    for ln in p.stdout:
        lines.append((time.perf_counter() - t0, ln.rstrip("\n")))
    rc = p.wait()
    if rc != 0:
        raise RuntimeError(f"reference failed rc={rc}: {lines[-5:]}")
    return lines, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/bsx_scale")
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--skip-oracle", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True); os.makedirs(SC.GOLDEN_DIR, exist_ok=True)
    T = len(os.sched_getaffinity(0))

    log("genome")
    genome = SC.make_genome("cpu")
    gsha = SC.sha(np.concatenate([g.numpy()[:1 << 20] for g in genome]))
    w = SC.SE
    log("SE reads")
    se = SC.se_reads(genome, w).numpy()
    se_lens = np.full(w["n"], w["L"], dtype=np.uint16)
    se_names = SC.se_read_names(genome, w)
    fa, fq, sam = (os.path.join(a.work, f) for f in ("ref.fa", "reads.fq", "ref_out.sam"))

    ref_info = None
    if not a.skip_reference:
        log("writing FASTA / FASTQ")
        if not os.path.exists(fa):
            synth.write_fasta(fa, genome, SC.NAMES)
        synth.write_fastq(fq, se[:, :w["L"]], se_names)
        log("reference binary")
        lines, tot = timed_reference(["-a", fq, "-d", fa, "-o", sam, "-s", 16, "-v", 5, "-I", 4, "-S", 7, "-p", T], a.work)
        t_load = next(t for t, ln in lines if ln.startswith("Load in"))
        t_tab = next(t for t, ln in lines if ln.startswith("Create seed table") or ln.startswith("max mismatches"))
        t_first = next(t for t, ln in lines if "reads finished" in ln)
        t_done = next(t for t, ln in lines if ln.startswith("Done."))
        text = open(sam, "rb").read()
        dig, nl = SC.sorted_sam_digest(text)
        ref_info = dict(command=f"bsmap -a reads.fq -d ref.fa -o out.sam -s 16 -v 5 -I 4 -S 7 -p {T}", host_threads=T,
                        reads=w["n"], fasta_load_s=t_load, seed_table_s=t_tab - t_load, mapping_s=t_done - t_tab,
                        first_batch_report_s=t_first - t_tab, total_s=tot, reads_per_s=w["n"] / (t_done - t_tab),
                        sam_bytes=len(text), sam_lines=nl, sorted_sam_sha256=dig,
                        stdout_tail=[ln for _, ln in lines[-6:]], where="CPU container (8 cores), tests/golden/make_scale_golden.py")
        log("reference:", json.dumps(ref_info)[:400])

    if a.skip_oracle:
        if ref_info:
            json.dump(ref_info, open(SC.REFRUN, "w"), indent=1)
        return

    log("oracle index build (bso_ref_create)")
    t0 = time.perf_counter()
    op = O.make_params(**SC.INDEX_OPTS)
    oref = O.OracleRef(op, SC.NAMES, [g.numpy() for g in genome])
    build_s = time.perf_counter() - t0
    log(f"built in {build_s:.0f} s, entries {oref.n_entries}")
    refcat, crefcat, tab, pos = oref.refcat, oref.crefcat, oref.tab, oref.pos
    n = oref.n_entries
    dg = dict(genome=dict(seed=SC.GENOME_SEED, chroms=SC.CHROMS, chrom_len=SC.CHROM_LEN, first_mib_of_each_chrom_sha256=gsha),
              index_opts=SC.INDEX_OPTS, n_words=oref.n_words, n_keys=oref.n_keys, n_entries=n, chunk=SC.CHUNK,
              oracle_build_seconds=build_s,
              arrays=dict(refcat=SC.sha(refcat), crefcat=SC.sha(crefcat), anchor=SC.sha(oref.anchor), tab=SC.sha(tab)))
    log("digest pos")
    dg["arrays"]["pos"] = SC.sha_chunks(lambda lo, hi: pos[lo:hi], n, threads=T)
    log("digest ctx")
    dg["arrays"]["ctx"] = SC.sha_chunks(lambda lo, hi: SC.inline_context(refcat, crefcat, tab, pos, lo, hi, 16, 0), n, threads=4)
    log("digest ctx2")
    dg["arrays"]["ctx2"] = SC.sha_chunks(lambda lo, hi: SC.inline_context(refcat, crefcat, tab, pos, lo, hi, 16, 16), n, threads=4)

    wl = {}
    log("oracle SE")
    p = O.make_params(**w["opts"])
    t0 = time.perf_counter()
    recs, counts, st = par_map_se(oref, p, se, se_lens, T)
    dt = time.perf_counter() - t0
    wl[w["name"]] = dict(n=w["n"], opts=w["opts"], recs=SC.sha(recs), counts=SC.sha(counts), candidates=int(st[0]),
                         mapped=int((recs["nhits"] > 0).sum()), unique=int((recs["nhits"] == 1).sum()), oracle_seconds=dt, threads=T)
    log(json.dumps(wl[w["name"]]))
    if ref_info:
        # the oracle's SAM for the same reads must be the reference binary's (sorted: -p T interleaves batches)
        quals = [b"I" * w["L"]] * w["n"]
        seqs = [bytes(r[:w["L"]]) for r in se]
        txt, na = oref.format_se(se_names, seqs, quals, recs, counts, params=p)
        odig, onl = SC.sorted_sam_digest(txt)
        ref_info["oracle_sorted_sam_sha256"] = odig
        ref_info["oracle_equals_reference"] = bool(odig == ref_info["sorted_sam_sha256"] and onl == ref_info["sam_lines"])
        log("oracle SAM == reference SAM:", ref_info["oracle_equals_reference"])
        json.dump(ref_info, open(SC.REFRUN, "w"), indent=1)

    log("oracle PE")
    w = SC.PE
    pa, pb = (x.numpy() for x in SC.pe_reads(genome, w))
    p = O.make_params(**w["opts"])
    t0 = time.perf_counter()
    pr, ra, rb, st = par_map_pe(oref, p, pa, pb, np.full(w["n"], w["L"], dtype=np.uint16), T)
    wl[w["name"]] = dict(n=w["n"], opts=w["opts"], pairs=SC.sha(pr), recs_a=SC.sha(ra), recs_b=SC.sha(rb), candidates=int(st[0]),
                         paired=int(pr["paired"].sum()), oracle_seconds=time.perf_counter() - t0, threads=T)
    log(json.dumps(wl[w["name"]]))

    log("oracle wide")
    w = SC.WIDE
    wd = SC.se_reads(genome, w).numpy()
    p = O.make_params(**w["opts"])
    t0 = time.perf_counter()
    recs, counts, st = par_map_se(oref, p, wd, np.full(w["n"], w["L"], dtype=np.uint16), T)
    wl[w["name"]] = dict(n=w["n"], opts=w["opts"], recs=SC.sha(recs), counts=SC.sha(counts), candidates=int(st[0]),
                         mapped=int((recs["nhits"] > 0).sum()), oracle_seconds=time.perf_counter() - t0, threads=T)
    log(json.dumps(wl[w["name"]]))
    dg["workloads"] = wl
    json.dump(dg, open(SC.DIGESTS, "w"), indent=1)
    log("wrote", SC.DIGESTS)


if __name__ == "__main__":
    main()
