"""Multi-GPU host logic: reads shard across ranks, the index is replicated, records merge in input order.

The path has no exchange step: one process per GPU, contiguous read chunks dealt
round-robin, every chunk carrying its global read index base (needed by myrand and by output order).
The only collectives are (a) the one-time index broadcast and (b) the timing / counter reductions of
the bench.  Everything here is backend-agnostic torch.distributed (NCCL on the GPU box, gloo in the CPU
tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def chunk_plan(n_reads: int, world: int, chunk: int) -> List[Tuple[int, int, int]]:
    """(rank, first_index, count) for every chunk: contiguous chunks dealt round-robin"""
    plan, start, k = [], 0, 0
    while start < n_reads:
        cnt = min(chunk, n_reads - start)
        plan.append((k % world, start, cnt))
        start += cnt
        k += 1
    return plan


def my_chunks(n_reads: int, rank: int, world: int, chunk: int) -> List[Tuple[int, int]]:
    return [(s, c) for r, s, c in chunk_plan(n_reads, world, chunk) if r == rank]


def merge_in_order(n_reads: int, world: int, chunk: int, per_rank: Sequence[np.ndarray]) -> np.ndarray:
    """records of every rank (each in its own chunk order) -> one array in input read order"""
    out = np.empty(n_reads, dtype=per_rank[0].dtype)
    cursor = [0] * world
    for r, s, c in chunk_plan(n_reads, world, chunk):
        out[s:s + c] = per_rank[r][cursor[r]:cursor[r] + c]
        cursor[r] += c
    return out


def broadcast_blob(blob: bytes | None, src: int = 0, device="cpu") -> bytes:
    """index metadata from the building rank to everyone"""
    box = [blob]
    dist.broadcast_object_list(box, src=src, device=torch.device(device) if device != "cpu" else None)
    return box[0]


def broadcast_buffers(tensors: Sequence[torch.Tensor], src: int = 0) -> None:
    """the one-time index broadcast: every replica array, in place"""
    for t in tensors:
        if t.numel():
            dist.broadcast(t, src=src)


def gather_records(local: np.ndarray, device="cpu") -> List[np.ndarray]:
    """all ranks' record arrays on every rank (host merge in input order follows)"""
    world = dist.get_world_size()
    t = torch.from_numpy(local.view(np.uint8).reshape(-1).copy()).to(device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device=device))
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, dtype=torch.uint8, device=device)
    pad[:t.numel()] = t
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[:int(s)].cpu().numpy().view(local.dtype) for b, s in zip(bufs, sizes)]


def reduce_max(values: Sequence[float], device="cpu") -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def reduce_sum(values: Sequence[float], device="cpu") -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]
