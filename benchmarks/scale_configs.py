#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] at scale, one process per GPU (torchrun or plain python).

  --config4  stream N synthetic clips (default 100 000) through log-mel, 128 mels, frame 1024 / hop 512,
             in device-resident chunks (the train_cnn input shape); clips shard across ranks.
  --config5  N clips (default 1 000 000): MFCC clip embeddings per rank, ONE all-gather of the database
             embeddings, top-20 of every query against the whole database, hit@10/20 all-reduced.
             80/20 database/query split (clip i is a query when (i // 50) % 5 == 4), labels i % 50.
             A sample of queries is re-ranked by the CPU oracle against the full database: indices must be identical.

Synthetic clips are generated on the device chunk by chunk (never on the host); generation is outside the
timed regions, which use CUDA events around our kernels only.  One JSON line per config from rank 0.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config4", action="store_true")
    ap.add_argument("--config5", action="store_true")
    ap.add_argument("--clips", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=4000)
    ap.add_argument("--oracle-queries", type=int, default=256)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from dsp_final_b200 import dist as D
    from dsp_final_b200 import retrieval as R
    from dsp_final_b200 import synth
    from dsp_final_b200.batch import features_batch
    from dsp_final_b200.dsp.mfcc import MfccConfig

    rank, world, local = D.init_process_group()
    dev = torch.device("cuda", local)

    def timed_stream(n_total, cfg, want, seed, keep=None):
        """Run `want` over this rank's shard chunk by chunk; returns (kernel seconds, wall seconds, kept outputs)."""
        b0, b1 = D.shard_range(n_total, rank, world)
        kept, ker_ms = [], 0.0
        t_wall = time.perf_counter()
        for s in range(b0, b1, args.chunk):
            e = min(b1, s + args.chunk)
            clips = synth.device_clips(e - s, seed=seed, device=dev, first=s)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = features_batch(clips, cfg, want)
            e1.record()
            torch.cuda.synchronize(dev)
            ker_ms += e0.elapsed_time(e1)
            if keep:
                kept.append(out[keep])
            del clips, out
        wall = time.perf_counter() - t_wall
        return ker_ms * 1e-3, wall, kept, (b0, b1)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config4:
        n = args.clips or 100_000
        cfg = MfccConfig(sample_rate=44100, frame_length=1024, hop_length=512, n_mels=128)
        features_batch(synth.device_clips(64, seed=1, device=dev), cfg, ("log_mel",))        # warm-up / plan
        if world > 1:
            dist.barrier()
        ker, wall, _, _ = timed_stream(n, cfg, ("log_mel",), seed=4)
        ker, wall = max_over_ranks(ker), max_over_ranks(wall)
        if rank == 0:
            print(json.dumps({"bench": "config4_logmel128_stream", "n_gpus": world, "clips": n, "chunk": args.chunk,
                              "audio_s_per_s_kernels": n * 5.0 / ker, "kernel_seconds": ker,
                              "wall_seconds_incl_generation": wall,
                              "bytes_per_clip": 4 * 220500 + 4 * 429 * 128, "algorithmic_GBps_per_gpu":
                              (n / world) * (4 * 220500 + 4 * 429 * 128) / ker / 1e9}), flush=True)

    if args.config5:
        n = args.clips or 1_000_000
        cfg = MfccConfig(sample_rate=44100, frame_length=1024, hop_length=512)
        features_batch(synth.device_clips(64, seed=1, device=dev), cfg, ("embed",))
        if world > 1:
            dist.barrier()
        ker, wall, kept, (b0, b1) = timed_stream(n, cfg, ("embed",), seed=5, keep="embed")
        emb = torch.cat(kept, dim=0)
        idx_global = torch.arange(b0, b1, device=dev)
        targets = (idx_global % 50).to(torch.int32)
        is_q = ((idx_global // 50) % 5) == 4          # every fifth block of 50 clips: all classes on both sides
        q, tq = emb[is_q].contiguous(), targets[is_q].contiguous()
        db, tdb = emb[~is_q].contiguous(), targets[~is_q].contiguous()
        ker, wall = max_over_ranks(ker), max_over_ranks(wall)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res, idx = D.sharded_retrieval(db, tdb, q, tq, (10, 20))
        torch.cuda.synchronize(dev)
        t_ret = max_over_ranks(time.perf_counter() - t0)
        n_q_total, n_db_total = res[0][2], n - res[0][2]
        # parity: rank 0 re-ranks a sample of its queries on the CPU against the full database
        parity = None
        full_db = D.all_gather_rows(db)                      # every rank takes part in the collective
        if rank == 0 and args.oracle_queries > 0:
            from oracle import oracle as O

            m = min(args.oracle_queries, q.shape[0])
            want = O.cosine_topk(q[:m].cpu().numpy(), full_db.cpu().numpy(), 20)
            parity = {"queries_checked": int(m), "identical_indices": bool(np.array_equal(idx[:m].cpu().numpy(), want))}
        if rank == 0:
            print(json.dumps({"bench": "config5_embed_plus_retrieval", "n_gpus": world, "clips": n,
                              "n_queries": n_q_total, "n_db": n_db_total,
                              "mfcc_embed_audio_s_per_s_kernels": n * 5.0 / ker, "feature_kernel_seconds": ker,
                              "feature_wall_seconds_incl_generation": wall,
                              "retrieval_seconds": t_ret, "queries_per_s": n_q_total / t_ret,
                              "pair_scores_per_s": n_q_total * n_db_total / t_ret,
                              "hit_at_10": res[0][1] / res[0][2], "hit_at_20": res[1][1] / res[1][2],
                              "retrieval_includes": "all-gather of DB embeddings (NCCL), tensor-core filter + exact FP64 re-score, top-20, hit@k, all-reduce",
                              "oracle_parity": parity}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
