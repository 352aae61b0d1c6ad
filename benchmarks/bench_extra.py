#!/usr/bin/env python
"""Secondary measurements (not the driver's headline bench.py contract).

  --sweep       MFCC audio-s/s for the frame x hop grid of configs/experiments.yaml (BASELINE.json configs[2])
  --retrieval   queries/s of cosine top-k at ESC-50 size (400 x 1600, k = 10/20) and at scale
  --logmel128   log-mel (128 mels) streaming throughput, train_cnn input shape (configs[3])

Each measurement prints one JSON line; CUDA-event timing, >= 3 warm-ups, inputs resident in HBM.
Under torchrun the retrieval leg shards queries and the database across ranks and performs the
one all-gather of database embeddings (configs[4]).
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


def _time_cuda(fn, steps, warmup, torch):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e-3


def sweep(args, torch):
    from dsp_final_b200 import synth
    from dsp_final_b200.batch import features_batch
    from dsp_final_b200.dsp.mfcc import MfccConfig
    from dsp_final_b200.plan import get_plan

    clips = synth.device_clips(args.clips, seed=1234, device=torch.device("cuda"))
    for fl in (512, 1024, 2048):
        for hop in (256, 512, 1024):
            cfg = MfccConfig(sample_rate=44100, frame_length=fl, hop_length=hop)
            plan = get_plan(cfg)
            dt = _time_cuda(lambda: features_batch(clips, cfg, ("mfcc",)), args.steps, 3, torch)
            t = plan.num_frames(clips.shape[1])
            nbytes = args.clips * (clips.shape[1] * 4 + t * 13 * 4)
            print(json.dumps({"bench": "mfcc_sweep", "frame_length": fl, "hop_length": hop, "kernel": plan.kernel,
                              "clips": args.clips, "audio_s_per_s": args.clips * 5.0 / dt, "ms": dt * 1e3,
                              "algorithmic_GBps": nbytes / dt / 1e9}), flush=True)


def logmel128(args, torch):
    from dsp_final_b200 import synth
    from dsp_final_b200.batch import features_batch
    from dsp_final_b200.dsp.mfcc import MfccConfig
    from dsp_final_b200.plan import get_plan

    clips = synth.device_clips(args.clips, seed=4, device=torch.device("cuda"))
    cfg = MfccConfig(sample_rate=44100, frame_length=1024, hop_length=512, n_mels=128)
    plan = get_plan(cfg)
    dt = _time_cuda(lambda: features_batch(clips, cfg, ("log_mel",)), args.steps, 3, torch)
    nbytes = args.clips * (clips.shape[1] * 4 + 429 * 128 * 4)
    print(json.dumps({"bench": "logmel128", "kernel": plan.kernel, "clips": args.clips,
                      "audio_s_per_s": args.clips * 5.0 / dt, "ms": dt * 1e3, "algorithmic_GBps": nbytes / dt / 1e9}),
          flush=True)


def retrieval(args, torch):
    import torch.distributed as dist

    from dsp_final_b200 import dist as D
    from dsp_final_b200 import retrieval as R

    rank, world, local = D.init_process_group()
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)
    for nq, ndb in ((400, 1600), (args.nq, args.ndb)):
        q0, q1 = D.shard_range(nq, rank, world)
        b0, b1 = D.shard_range(ndb, rank, world)
        q = torch.randn((q1 - q0, 26), generator=gen, device=dev)
        db_local = torch.randn((b1 - b0, 26), generator=gen, device=dev)
        tq = torch.randint(0, 50, (q1 - q0,), generator=gen, device=dev, dtype=torch.int32)
        tdb_local = torch.randint(0, 50, (b1 - b0,), generator=gen, device=dev, dtype=torch.int32)

        def step():
            res, _ = D.sharded_retrieval(db_local, tdb_local, q, tq, (10, 20))
            return res

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"bench": "retrieval_top20", "n_gpus": world, "n_queries": nq, "n_db": ndb, "dim": 26,
                              "queries_per_s": nq / float(tt.item()), "ms": float(tt.item()) * 1e3,
                              "pair_scores_per_s": nq * ndb / float(tt.item()),
                              "hit_at_10": res[0][1] / res[0][2], "hit_at_20": res[1][1] / res[1][2],
                              "includes": "all-gather of DB embeddings + FP64 scoring + top-20 + hit@k + all-reduce"}),
                  flush=True)
        # spot parity against the CPU oracle on a few queries (rank 0 only, small case only)
        if rank == 0 and ndb <= 2000 and world == 1:
            from oracle import oracle as O

            idx = R.cosine_topk(q, db_local, 20).cpu().numpy()
            assert np.array_equal(idx, O.cosine_topk(q.cpu().numpy(), db_local.cpu().numpy(), 20))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--retrieval", action="store_true")
    ap.add_argument("--logmel128", action="store_true")
    ap.add_argument("--clips", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--nq", type=int, default=20_000)
    ap.add_argument("--ndb", type=int, default=1_000_000)
    args = ap.parse_args()
    import torch

    if args.sweep:
        sweep(args, torch)
    if args.logmel128:
        logmel128(args, torch)
    if args.retrieval:
        retrieval(args, torch)


if __name__ == "__main__":
    main()
