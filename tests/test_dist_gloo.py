"""world_size-2 gloo tests of the multi-GPU host logic (sharding, the one all-gather, hit reduction).

The GPU kernels are replaced by the CPU oracle's callables here so that the exchange logic can
run in the dev container; on the B200 box the same code path runs over NCCL (bench.py --gpus N).
"""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def test_shard_range_covers_everything():
    from dsp_final_b200.dist import shard_range, shard_sizes

    for n in (0, 1, 7, 2000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_db: int, n_q: int, q):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    from dsp_final_b200 import dist as D
    from oracle import oracle as O

    try:
        r, w, _ = D.init_process_group("gloo")
        assert (r, w) == (rank, world)
        rng = np.random.default_rng(11)
        db = rng.standard_normal((n_db, 26)).astype(np.float32)
        qs = rng.standard_normal((n_q, 26)).astype(np.float32)
        tdb = rng.integers(0, 50, n_db).astype(np.int32)
        tq = rng.integers(0, 50, n_q).astype(np.int32)
        b0, b1 = D.shard_range(n_db, rank, world)
        q0, q1 = D.shard_range(n_q, rank, world)
        gathered = D.all_gather_rows(torch.as_tensor(db[b0:b1]))
        assert np.array_equal(gathered.numpy(), db)                       # ragged shards reassemble in rank order
        topk = lambda a, b, k: O.cosine_topk(np.asarray(a), np.asarray(b), k)          # noqa: E731
        hits = lambda idx, k, a, b: O.hits_at_k(idx, k, np.asarray(a), np.asarray(b))  # noqa: E731
        res, idx = D.sharded_retrieval(torch.as_tensor(db[b0:b1]), torch.as_tensor(tdb[b0:b1]),
                                       torch.as_tensor(qs[q0:q1]), torch.as_tensor(tq[q0:q1]), (10, 20),
                                       topk_fn=topk, hits_fn=hits)
        want_idx = O.cosine_topk(qs, db, 20)
        assert np.array_equal(idx, want_idx[q0:q1])
        want = [(k, O.hits_at_k(want_idx, k, tdb, tq), n_q) for k in (10, 20)]
        assert res == want, (res, want)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:                                                  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))


@pytest.mark.parametrize("n_db,n_q", [(1600, 400), (1601, 399)])
def test_sharded_retrieval_world2(n_db, n_q):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_db, n_q, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out
