"""Retrieval on class-structured, tightly clustered embeddings (every cosine close to 1, thousands of near-ties at the
k-th score) -- the regime of real MFCC statistics and of BASELINE configs[4] -- next to the usual Gaussian case."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(sys.argv[1] if len(sys.argv) > 1 else Path(__file__).resolve().parents[1])))
import torch

from dsp_final_b200 import retrieval as R

g = torch.Generator(device="cuda")
g.manual_seed(3)
nq, ndb, dim, classes = 25_000, 800_000, 26, 50
centers = torch.randn((classes, dim), generator=g, device="cuda")
offset = 4.0 * torch.randn((1, dim), generator=g, device="cuda")
for name, noise in (("clustered", 0.02), ("gaussian", None)):
    if noise is None:
        q = torch.randn((nq, dim), generator=g, device="cuda")
        db = torch.randn((ndb, dim), generator=g, device="cuda")
    else:
        q = offset + centers[torch.arange(nq, device="cuda") % classes] + noise * torch.randn((nq, dim), generator=g, device="cuda")
        db = offset + centers[torch.arange(ndb, device="cuda") % classes] + noise * torch.randn((ndb, dim), generator=g, device="cuda")
    R.cosine_topk(q, db, 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        idx = R.cosine_topk(q, db, 20)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: top20 {nq} x {ndb}: {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
