"""Quick timing of dspx_cosine_topk (top-20, dim 26) on the three shapes the retrieval notes quote:
Gaussian 20 000 x 1 000 000 and 8 192 x 262 144, and class-clustered 25 000 x 800 000 (every cosine > 0.9999).
Checks 48 queries of each case against the CPU oracle."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dsp_final_b200 import retrieval as R  # noqa: E402
from oracle import oracle as O  # noqa: E402


def clustered(n, dim, gen, dev):
    """50 class centres around a common offset (every cosine > 0.999), like MFCC mean/std embeddings."""
    cg = torch.Generator(device=dev)
    cg.manual_seed(3)
    centres = torch.randn((50, dim), generator=cg, device=dev)
    offset = 4.0 * torch.randn((1, dim), generator=cg, device=dev)
    cls = torch.randint(0, 50, (n,), generator=gen, device=dev)
    return offset + centres[cls] + 0.02 * torch.randn((n, dim), generator=gen, device=dev)


def main():
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    cases = [("gauss 20000x1000000", torch.randn((20_000, 26), generator=gen, device=dev), torch.randn((1_000_000, 26), generator=gen, device=dev)),
             ("gauss 8192x262144", torch.randn((8192, 26), generator=gen, device=dev), torch.randn((262_144, 26), generator=gen, device=dev)),
             ("clustered 25000x800000", clustered(25_000, 26, gen, dev), clustered(800_000, 26, gen, dev))]
    for name, q, db in cases:
        idx = R.cosine_topk(q, db, 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            idx = R.cosine_topk(q, db, 20)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        sel = torch.arange(0, q.shape[0], max(1, q.shape[0] // 48), device=dev)[:48]
        ok = np.array_equal(idx[sel].cpu().numpy(), O.cosine_topk(q[sel].cpu().numpy(), db.cpu().numpy(), 20))
        print(f"{name}: {ms:.2f} ms  {q.shape[0] * db.shape[0] / ms / 1e9:.2f}e12 pair-scores/s  oracle-identical={ok}", flush=True)


if __name__ == "__main__":
    main()
