"""Small driver for ncu: one cosine top-20 of 8192 queries against 262144 database rows (dim 26)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from dsp_final_b200 import retrieval as R

g = torch.Generator(device="cuda")
g.manual_seed(0)
q = torch.randn((8192, 26), generator=g, device="cuda")
db = torch.randn((262144, 26), generator=g, device="cuda")
for _ in range(3):
    idx = R.cosine_topk(q, db, 20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
idx = R.cosine_topk(q, db, 20)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"top20 8192 x 262144: {ms:.3f} ms -> {8192 * 262144 / ms / 1e6:.1f} G pair-scores/s, {8192 / ms * 1e3:.0f} queries/s")
