"""Per-role cycles of cosine_topk_tc_kernel's CTA (0, 0) per tile, cold (first call) and in steady state (second call
with DSPX_EXPERIMENT_KEEP_THR=1: every split starts from the final thresholds).  Needs the -DDSPX_TC_PROFILE library
(see retr_tc_roles.py)."""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from dsp_final_b200 import retrieval as R
g = torch.Generator(device="cuda"); g.manual_seed(0)
names = ["mma wait acc_empty", "mma wait full_b", "mma issue+commit", "prod wait empty", "prod store", "epi wait acc_full", "epi masks", "epi enqueue+drain", "tiles", "mma total"]
q = torch.randn((20000, 26), generator=g, device="cuda"); db = torch.randn((1000000, 26), generator=g, device="cuda")
h = ctypes.CDLL(os.environ["DSPX_LIBRARY"])
for call in range(3):
    R.cosine_topk(q, db, 20); torch.cuda.synchronize()
    out = (ctypes.c_longlong * 16)()
    h.dspx_debug_tc_prof(out)
    tiles = out[8]
    print("call", call, "tiles", tiles, " ".join(f"[{n}: {out[i]/max(tiles,1):.0f}]" for i, n in enumerate(names) if i != 8), flush=True)
    if hasattr(h, "dspx_debug_tc_trace") and call == 2:
        tr = (ctypes.c_longlong * (32 * 18))()
        h.dspx_debug_tc_trace(tr)
        t0 = tr[0]
        print("tile: mma(acc_empty seen, issue done) | epilogue warps: acc_full seen ... | released ...   (cycles from the first event)")
        for i in range(12):
            row = [tr[i * 18 + j] - t0 for j in range(18)]
            print(f"{64 + i}: mma {row[0]:6d} {row[1]:6d} | full " + " ".join(f"{x:6d}" for x in row[2:10]) + " | rel " + " ".join(f"{x:6d}" for x in row[10:18]), flush=True)
