"""One-off confidence check at BASELINE scale: top-20 of 1024 queries against 1 000 000 rows, Gaussian and tightly
clustered embeddings (thousands of rows within 1e-5 of the k-th score), indices and scores against the CPU oracle."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from dsp_final_b200 import retrieval as R
from oracle import oracle as O

g = torch.Generator(device="cuda")
g.manual_seed(11)
nq, ndb, dim, classes = 1024, 1_000_000, 26, 50
centers = torch.randn((classes, dim), generator=g, device="cuda")
offset = 4.0 * torch.randn((1, dim), generator=g, device="cuda")
for name, noise in (("gaussian", None), ("clustered", 0.02), ("very clustered", 0.002)):
    if noise is None:
        q = torch.randn((nq, dim), generator=g, device="cuda")
        db = torch.randn((ndb, dim), generator=g, device="cuda")
    else:
        q = offset + centers[torch.arange(nq, device="cuda") % classes] + noise * torch.randn((nq, dim), generator=g, device="cuda")
        db = offset + centers[torch.arange(ndb, device="cuda") % classes] + noise * torch.randn((ndb, dim), generator=g, device="cuda")
    idx, sc = R.cosine_topk(q, db, 20, return_scores=True)
    t0 = time.perf_counter()
    want_idx, want_sc = O.cosine_topk(q.cpu().numpy(), db.cpu().numpy(), 20, return_scores=True)
    dt = time.perf_counter() - t0
    ok_i = np.array_equal(idx.cpu().numpy(), want_idx)
    ok_s = np.array_equal(sc.cpu().numpy(), want_sc)
    gap = float(np.min(want_sc[:, 18] - want_sc[:, 19]))
    print(f"{name}: indices identical {ok_i}, scores identical {ok_s}; smallest gap between 19th and 20th score {gap:.3e}; oracle {dt:.1f} s", flush=True)
