#!/usr/bin/env python
"""bench.py -- WGBS 100-nt single-end reads mapped per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload "cfg2"): synthetic 3.1 Gb genome (25 x 124 Mb, counter-based random,
seed 2), simulated directional bisulfite reads, 100 nt, -s 16 -v 5 -I 4 (BASELINE.json configs[1]).
One step = one pass of the hot path (seed selection + probe + extension + best-hit selection,
bsx_map_se) over 20 M reads per GPU.  Reads shard across ranks (independent units, no collective
on the data path): weak scaling, every rank maps its own 20 M reads against its own replica of the
index, which rank 0 builds and broadcasts once over NVLink (NCCL broadcast, outside the timed region).

    value     reads/s, inputs resident in HBM, CUDA events around K launches of the mapping kernel
    e2e       same metric through bsx_map_se with pinned HOST buffers: H2D + kernel + D2H inside the
              timed region (two-stream sub-batch pipeline)
    roofline  algorithmic bytes (B = 2*ceil(L/4) + 12 P + C (4 + L/4) + 32 per read, with
              the kernel's exact C counter and the distinct-header count P) / kernel time vs the
              measured HBM copy peak (MEASURED_PEAKS.json)
    cpu_baseline  the oracle port (oracle/bsmap_oracle.c) on the host cores over a bounded sample of
              the same reads (index arrays imported, so only MAPPING is timed)

--impl reference times that same CPU port with all host threads (see DESIGN.md for why the real
reference binary, whose single-threaded table build alone takes 250-300 s at 3.1 Gb, is not run
inside the bench).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_READ = 100
STRIDE = 104      # bytes per read slot: the 100 bases rounded up to the ABI's 8-byte alignment
OPTS = dict(s=16, v=5, I=4, S=7)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chroms", type=int, default=25)
    ap.add_argument("--chrom-mb", type=float, default=124.0)
    ap.add_argument("--reads", type=int, default=20_000_000, help="reads per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--batch", type=int, default=1 << 20, help="sub-batch of the end-to-end pipeline")
    return ap.parse_args()


class _DevBuf:
    """zero-copy torch view of a raw device pointer (for the NCCL index broadcast)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_map_parallel(oref, buf, lens, first_index, threads):
    """oracle port over `threads` host threads (ctypes releases the GIL); returns (seconds, stats, records)"""
    from concurrent.futures import ThreadPoolExecutor
    n = len(lens)
    cuts = np.linspace(0, n, threads + 1).astype(int)
    stats, recs = [], [None] * threads

    def work(i):
        a, b = int(cuts[i]), int(cuts[i + 1])
        if b > a:
            recs[i], _, st = oref.map_se(buf[a:b], lens[a:b], first_index=first_index + a, want_counts=False)
            stats.append(st)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return dt, np.sum(stats, axis=0), np.concatenate([r for r in recs if r is not None])


def bind_near_gpu(torch, local):
    """Host side of the end-to-end path: run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA
    node its GPU hangs off, so that eight ranks do not pull their input through one socket.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(torch.cuda.get_device_properties(local).pci_bus_id.encode()
                                                 if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id")
                                                 else pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference" and rank != 0:
        return 0
    import torch
    import torch.distributed as dist
    import bsmap_b200 as B
    from bsmap_b200 import shard, synth
    from bsmap_b200.lib import REC

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_near_gpu(torch, local) if world > 1 else 0
    multi = world > 1 and a.impl == "ours"
    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    chrom_len = int(a.chrom_mb * 1_000_000)
    lens = [chrom_len] * a.chroms
    names = [f"chr{i + 1}" for i in range(a.chroms)]
    n = a.reads
    t_setup = time.perf_counter()

    # ---- synthetic genome on the GPU; host copy only where the index is built
    genome = synth.make_genome(2, lens, device=dev)
    p = B.make_params(**OPTS)
    build_s, bcast_s = None, None
    if rank == 0 or not multi:
        host_g = [torch.empty(ln, dtype=torch.uint8, pin_memory=True) for ln in lens]
        for h, g in zip(host_g, genome):
            h.copy_(g)
        torch.cuda.synchronize()
        ix = B.Index.from_pointers(p, names, [h.data_ptr() for h in host_g], lens, device=local)
        build_s = ix.info.build_seconds
        del host_g
    if multi:
        # one-time index broadcast over NVLink: metadata by object broadcast, arrays by NCCL broadcast
        meta = shard.broadcast_blob(ix.meta() if rank == 0 else None, src=0, device=dev)
        if rank != 0:
            ix = B.Index.shell(p, meta, local)
        t_b = time.perf_counter()
        shard.broadcast_buffers([torch.as_tensor(_DevBuf(ptr, nbytes), device=dev) for ptr, nbytes in ix.device_buffers() if nbytes], src=0)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - t_b

    # ---- simulated reads for this rank (distinct per rank), generated in HBM
    first_index = rank * n
    seq_dev = torch.zeros((n, STRIDE), dtype=torch.uint8, device=dev)
    CH = 1 << 21
    for s0 in range(0, n, CH):
        m = min(CH, n - s0)
        sim = synth.simulate_reads(genome, m, L_READ, seed=2024, subs="cfg2", first_index=first_index + s0)
        seq_dev[s0:s0 + m, :L_READ] = sim["seq"]
        del sim
    if os.environ.get("BSX_BENCH_REPEAT_READS"):
        # diagnostic only (never a bench value): the first K reads repeated, so that every table / list / image access
        # hits cache -- same instruction stream, no DRAM; tells how far the memory system holds the kernel back
        k = int(os.environ["BSX_BENCH_REPEAT_READS"])
        seq_dev[:] = seq_dev[:k].repeat((n + k - 1) // k, 1)[:n]
    del genome
    torch.cuda.empty_cache()
    len_dev = torch.full((n,), L_READ, dtype=torch.int16, device=dev)
    seq_host = torch.empty((n, STRIDE), dtype=torch.uint8, pin_memory=True); seq_host.copy_(seq_dev)
    len_host = torch.empty((n,), dtype=torch.int16, pin_memory=True); len_host.copy_(len_dev)
    rec_host = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    del seq_dev, len_dev
    torch.cuda.empty_cache()
    setup_s = time.perf_counter() - t_setup

    cfg = {"workload": f"cfg2: {a.chroms}x{a.chrom_mb:g}Mb synthetic genome, {n} x {L_READ}nt SE reads per GPU per step, -s 16 -v 5 -I 4",
           "l2_policy": "working set larger than L2 (index 21 GB + 2.2 GB reads per step vs 126 MB L2)",
           "reads_per_gpu_per_step": n, "genome_bp": sum(lens)}
    if numa_cpus:
        cfg["host_binding"] = f"each rank pinned to the {numa_cpus} CPUs NVML reports as local to its GPU (pinned buffers first-touched there)"

    # =====================================================================================
    if a.impl == "reference":
        # CPU arm: the oracle port with every host thread, on bounded samples of the same reads
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        op = O.make_params(**OPTS)
        T = host_threads()
        sample = a.cpu_sample or min(n, 40_000 * T)
        buf = seq_host.numpy(); ln = len_host.numpy().view(np.uint16)
        # the GPU's records for the same reads (outside the timed region): every record the CPU arm produces is compared
        gm = B.Mapper(ix, p, max_batch=1 << 20, stride=STRIDE)
        gpu_recs = np.empty(n, dtype=REC)
        gm.map_se_ptr(n, seq_host.data_ptr(), len_host.data_ptr(), gpu_recs.ctypes.data, first_index=first_index)
        gm.close()
        arrs = [ix.download(w) for w in ("refcat", "crefcat", "tab", "pos")]
        ix.close()
        oref = O.OracleRef.imported(op, names, lens, *arrs)
        times, compared, mismatching = [], 0, 0
        for it in range(a.warmup + a.steps):
            s0 = (it * sample) % max(1, n - sample + 1)
            dt, _, orec = cpu_map_parallel(oref, buf[s0:s0 + sample], ln[s0:s0 + sample], first_index + s0, T)
            compared += sample
            mismatching += int((orec != gpu_recs[s0:s0 + sample].astype(orec.dtype)).sum())
            if it >= a.warmup:
                times.append(dt)
        tot = sum(times)
        val = sample * a.steps / tot
        line = {"metric": "wgbs_100nt_reads_mapped_per_sec", "value": val, "unit": "reads/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "impl": "reference",
                "cpu_baseline": {"value": val, "unit": "reads/s", "cores": T, "kind": "port",
                                 "sample": f"{sample} reads per step of the same read set; index arrays imported (mapping only)",
                                 "records_compared_with_gpu": compared, "records_differing_from_gpu": mismatching},
                "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # =====================================================================================
    mp = B.Mapper(ix, p, max_batch=n, stride=STRIDE)
    # the kernel is launched on this (non-default) torch stream, so torch.cuda.Event brackets it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    mp.upload(n, seq_host.data_ptr(), len_host.data_ptr(), stream=stream)
    torch.cuda.synchronize()

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM
    for _ in range(a.warmup):
        mp.run_se(n, first_index=first_index, stream=stream)
    barrier()
    mp.stats(reset=True)
    l0 = mp.launches
    clocks = ClockSampler(local); clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(tstream)
    for _ in range(a.steps):
        mp.run_se(n, first_index=first_index, stream=stream)
    e1.record(tstream)
    barrier()
    ms = e0.elapsed_time(e1)
    st = mp.stats(reset=True)
    launches_value = mp.launches - l0
    recs_dev, _ = mp.download_se(n, stream=stream)
    mapped_frac = float((recs_dev["nhits"] > 0).mean())

    # ---- end to end: pinned host buffers in, host records out, every step
    small = B.Mapper(ix, p, max_batch=a.batch, stride=STRIDE)
    for _ in range(max(1, a.warmup // 2)):
        small.map_se_ptr(n, seq_host.data_ptr(), len_host.data_ptr(), rec_host.data_ptr(), first_index=first_index)
    barrier()
    l1 = small.launches
    t0 = time.perf_counter()
    for _ in range(a.steps):
        small.map_se_ptr(n, seq_host.data_ptr(), len_host.data_ptr(), rec_host.data_ptr(), first_index=first_index)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = small.launches - l1
    clk = clocks.stop()
    e2e_recs = rec_host.numpy().view(REC).reshape(-1)
    same = bool(np.array_equal(e2e_recs, recs_dev))

    # max over ranks
    ms, e2e_s = shard.reduce_max([ms, e2e_s], device=dev)
    tot_c, tot_p, tot_over, tot_full, tot_list, tot_gather = shard.reduce_sum(
        [st[k] for k in ("candidates", "probes", "overfetch", "full_extensions", "list_entries", "gathers")], device=dev)
    reads_total = n * world * a.steps
    value = reads_total / (ms * 1e-3)
    e2e_val = reads_total / e2e_s

    # ---- roofline of the mapping kernel (per launch, this rank's view scaled by world)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    c_per_read = tot_c / reads_total
    p_per_read = tot_p / reads_total
    bytes_per_read = 2 * ((L_READ + 3) // 4) + 12 * p_per_read + c_per_read * (4 + L_READ / 4) + 32
    per_launch_bytes = bytes_per_read * n
    kernel_s = (ms * 1e-3) / a.steps
    achieved = per_launch_bytes / kernel_s / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "kernel": "bsx_map_se_wgbs_kernel", "algorithmic_bytes_per_read": bytes_per_read,
            "candidates_per_read": c_per_read, "headers_per_read": p_per_read,
            "overfetch_per_read": tot_over / reads_total, "full_extensions_per_read": tot_full / reads_total,
            "hbm_gathers_per_read": tot_gather / reads_total}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roof["traffic"] = prof.get("dram_bytes_per_read", 0) * n or None
    except Exception:
        pass

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on a bounded sample, mapping only
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        arrs = [ix.download(w) for w in ("refcat", "crefcat", "tab", "pos")]
        oref = O.OracleRef.imported(O.make_params(**OPTS), names, lens, *arrs)
        T = host_threads()
        sample = a.cpu_sample or min(n, 40_000 * T)
        buf = seq_host.numpy(); ln = len_host.numpy().view(np.uint16)
        dt, ost, orec = cpu_map_parallel(oref, buf[:sample], ln[:sample], first_index, T)
        cpu = {"value": sample / dt, "unit": "reads/s", "cores": T, "kind": "port",
               "sample": f"first {sample} reads of the step ({dt:.1f} s wall on {T} threads); index arrays imported (mapping only)",
               "candidates_per_read": float(ost[0]) / sample, "headers_per_read": float(ost[1]) / sample,
               "records_compared_with_gpu": int(sample),
               "records_differing_from_gpu": int((orec != recs_dev[:sample].astype(orec.dtype)).sum())}
        oref.close()
        del arrs

    if rank == 0:
        line = {"metric": "wgbs_100nt_reads_mapped_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "config": cfg, "clocks": clk,
                "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": n * (STRIDE + 2), "d2h_bytes_per_step": n * 16,
                        "ms_per_step": 1e3 * e2e_s / a.steps, "records_identical_to_resident_run": same},
                "gpu_launches": int(launches_value + launches_e2e),
                "roofline": roof, "cpu_baseline": cpu,
                "mapped_fraction": mapped_frac, "index_build_seconds": build_s, "index_broadcast_seconds": bcast_s,
                "setup_seconds": setup_s}
        print(json.dumps(line))
    mp.close(); small.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
