"""Tuning aid: cosine top-20 time as a function of the database split count (DSPX_TOPK_SPLITS)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from dsp_final_b200 import retrieval as R

g = torch.Generator(device="cuda")
g.manual_seed(0)
for nq, ndb in ((8192, 262144), (20000, 1000000), (2048, 1000000), (100000, 100000)):
    q = torch.randn((nq, 26), generator=g, device="cuda")
    db = torch.randn((ndb, 26), generator=g, device="cuda")
    out = []
    for sp in (0, 1, 2, 3, 4, 5, 6, 8, 9, 12, 16):
        if sp:
            os.environ["DSPX_TOPK_SPLITS"] = str(sp)
        else:
            os.environ.pop("DSPX_TOPK_SPLITS", None)
        R.cosine_topk(q, db, 20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            R.cosine_topk(q, db, 20)
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{sp}:{e0.elapsed_time(e1) / 3:.2f}")
    print(f"{nq} x {ndb}  splits:ms  " + "  ".join(out), flush=True)
