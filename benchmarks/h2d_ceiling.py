#!/usr/bin/env python
"""What the box gives N concurrent ranks that do nothing but copy: the ceiling of every end-to-end number.

    python benchmarks/h2d_ceiling.py                                        one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/h2d_ceiling.py

Each rank copies 2000 clips x 220 500 float32 (1.764 GB, bench.py's input) from pinned host memory to its GPU
with one cudaMemcpyAsync per pass, all ranks at once: host->device alone, device->host alone and both directions
overlapped.  Rank 0 prints one JSON line with the GB/s per rank (min / mean over ranks) and each rank's
GPU -> PCI bus -> NUMA node -> CPU set, which is what bench.py's `e2e.copy_ceiling` and `topology` summarise.
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import bind_to_gpu_numa  # noqa: E402
from dsp_final_b200.dist import init_process_group  # noqa: E402


def main():
    rank, world, local = init_process_group()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    topo = bind_to_gpu_numa(local)
    n = 2000 * 220_500
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.fill_(1.0)                                                   # touch: the pages exist where this rank runs
    d = torch.empty(n, dtype=torch.float32, device=dev)
    h2 = torch.empty(n // 8, dtype=torch.float32, pin_memory=True)
    d2 = torch.zeros(n // 8, dtype=torch.float32, device=dev)
    up, dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(mode, reps=5):
        def step():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(up):
                    d.copy_(h, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(dn):
                    h2.copy_(d2, non_blocking=True)
            up.synchronize()
            dn.synchronize()

        step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = time.perf_counter() - t0
        nbytes = (n * 4 if mode in ("h2d", "both") else 0) + (n // 2 if mode in ("d2h", "both") else 0)
        return nbytes * reps / dt / 1e9

    res = {m: run(m) for m in ("h2d", "d2h", "both")}
    rows = [dict(rank=rank, **topo, **{f"{k}_gb_s": v for k, v in res.items()})]
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, dict(rank=rank, **topo, **{f"{k}_gb_s": v for k, v in res.items()}))
    if rank == 0:
        summary = {"n_ranks": world, "bytes_h2d_per_pass": n * 4}
        for m in ("h2d", "d2h", "both"):
            vals = [r[f"{m}_gb_s"] for r in rows]
            summary[m] = {"min_gb_s_per_rank": min(vals), "mean_gb_s_per_rank": sum(vals) / len(vals), "aggregate_gb_s": sum(vals)}
        summary["ranks"] = rows
        print(json.dumps(summary), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
