"""The N>1 host path on CPU: two gloo ranks shard a read set, 'map' their chunks with the oracle
(standing in for the device: this test is about sharding, index-metadata broadcast, reductions and
the in-order merge, not about kernels), and the merged records must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases as CS
    import oracle_lib as O
    from bsmap_b200 import shard
    case = CS.BY_NAME["se_cfg1"]
    d = case.data()
    n, chunk = len(d["seqs"]), 257
    # metadata blob travels from rank 0 (here: the case name stands in for the index metadata)
    blob = shard.broadcast_blob(b"index-meta:" + case.name.encode() if rank == 0 else None)
    assert blob == b"index-meta:se_cfg1"
    # the "index arrays": rank 0 fills, everyone receives in place
    t = torch.arange(1000, dtype=torch.int32) if rank == 0 else torch.zeros(1000, dtype=torch.int32)
    shard.broadcast_buffers([t, torch.zeros(0)])
    assert int(t.sum()) == 499500
    oref = O.OracleRef(O.make_params(**case.param_kwargs()), d["gnames"], d["gseqs"])
    buf, lens = O.pack_reads(d["seqs"])
    mine = []
    for s, c in shard.my_chunks(n, rank, world, chunk):
        recs, _, _ = oref.map_se(buf[s:s + c], lens[s:s + c], first_index=s, want_counts=False)
        mine.append(recs)
    local = np.concatenate(mine) if mine else np.zeros(0, dtype=O.REC)
    parts = shard.gather_records(local)
    merged = shard.merge_in_order(n, world, chunk, parts)
    mx = shard.reduce_max([float(rank + 1), 5.0 - rank])
    sm = shard.reduce_sum([float(len(local))])
    if rank == 0:
        full, _, _ = oref.map_se(buf, lens, first_index=0, want_counts=False)
        q.put((bool(np.array_equal(merged, full)), mx, sm, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, mx, sm, n = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "merged records differ from the single-process run"
    assert mx == [2.0, 5.0] and sm == [float(n)]


def test_chunk_plan_is_a_partition():
    from bsmap_b200 import shard
    for n, w, c in [(0, 2, 5), (1, 8, 5), (1000, 3, 7), (4096, 8, 512)]:
        plan = shard.chunk_plan(n, w, c)
        assert sum(x[2] for x in plan) == n
        assert [x[1] for x in plan] == list(np.cumsum([0] + [x[2] for x in plan[:-1]]))[:len(plan)]
        assert all(x[0] == i % w for i, x in enumerate(plan))
