"""Multi-GPU plumbing: one process per GPU, torch.distributed for the one exchange.

Clips are independent units: feature extraction shards by
contiguous clip ranges with no communication.  Retrieval has exactly one
exchange -- every rank needs the whole database of clip embeddings
([N_db, 2*n_mfcc] float32: 104 MB for a million clips) -- done with a single
all-gather (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).  Each
rank then ranks its own query shard against the full database; hit counts are
summed with one all-reduce of len(k_list) integers.
"""
from __future__ import annotations

import os
from typing import Iterable, Sequence

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of n units for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when absent."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE; returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def all_gather_rows(local, total_rows: int | None = None):
    """Concatenate per-rank row blocks [n_r, d] (ragged n_r allowed) into [sum n_r, d] on every rank.

    torch tensor in -> torch tensor out (same device).  With equal shards this is one
    all_gather_into_tensor; ragged shards are padded to the largest and trimmed.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    if total_rows is not None and sum(counts) != total_rows:
        raise RuntimeError(f"shards hold {sum(counts)} rows, expected {total_rows}")
    width = local.shape[1:]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0], *width), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    biggest = max(counts)
    padded = torch.zeros((biggest, *width), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * biggest, *width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * biggest: r * biggest + counts[r]] for r in range(world)], dim=0)


def all_reduce_sum_ints(values: Sequence[int], device=None) -> list[int]:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def sharded_retrieval(local_db_emb, local_db_targets, local_q_emb, local_q_targets, k_list: Iterable[int],
                      topk_fn=None, hits_fn=None, profile: dict | None = None):
    """Distributed evaluate_retrieval (src/retrieval/retrieval.py:52-72 at scale).

    Each rank holds a shard of database embeddings/targets and a shard of queries.
    Returns ([(k, hits_global, n_queries_global)], local_topk_idx).  topk_fn/hits_fn
    default to the GPU kernels; the gloo CPU tests inject reference callables so the
    exchange logic is tested without a GPU.  With `profile` (a dict) and CUDA tensors,
    the device time of the three phases is recorded with CUDA events on the current
    stream: profile["gather_ms" | "topk_ms" | "reduce_ms"], plus "gather_bytes" (what
    this rank receives) and "db_rows".
    """
    import torch

    k_list = [int(k) for k in k_list]
    on_gpu = isinstance(local_q_emb, torch.Tensor) and local_q_emb.is_cuda
    fast_hits = hits_fn is None and on_gpu          # hit counters stay on the device: one all-reduce, one host read
    if topk_fn is None or hits_fn is None:
        from . import retrieval as R

        topk_fn = topk_fn or R.cosine_topk
        hits_fn = hits_fn or R.hits_at_k
    marks = []

    def mark():
        if profile is not None and on_gpu:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(local_q_emb.device))
            marks.append(ev)

    mark()
    db = all_gather_rows(local_db_emb)
    tdb = all_gather_rows(local_db_targets.reshape(-1, 1)).reshape(-1)
    mark()
    kmax = min(max(k_list), db.shape[0])
    idx = topk_fn(local_q_emb, db, kmax)
    mark()
    if fast_hits:
        import torch.distributed as dist

        from .retrieval import hits_at_k_device

        counts = hits_at_k_device(idx, [min(k, kmax) for k in k_list], tdb, local_q_targets)
        t = torch.cat([counts, torch.tensor([int(local_q_emb.shape[0])], dtype=torch.int64, device=counts.device)])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        mark()
        tot = [int(v) for v in t.tolist()]
    else:
        hits = [hits_fn(idx, min(k, kmax), tdb, local_q_targets) for k in k_list]
        dev = local_q_emb.device if isinstance(local_q_emb, torch.Tensor) else None
        tot = all_reduce_sum_ints(hits + [int(local_q_emb.shape[0])], device=dev)
        mark()
    if profile is not None:
        profile["db_rows"] = int(db.shape[0])
        profile["gather_bytes"] = int((db.shape[0] - local_db_emb.shape[0]) * (db.shape[1] * db.element_size() + tdb.element_size()))
        if len(marks) == 4:
            marks[-1].synchronize()
            profile["gather_ms"] = marks[0].elapsed_time(marks[1])
            profile["topk_ms"] = marks[1].elapsed_time(marks[2])
            profile["reduce_ms"] = marks[2].elapsed_time(marks[3])
    return [(k, tot[i], tot[-1]) for i, k in enumerate(k_list)], idx
