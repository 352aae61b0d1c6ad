#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture for profiles/.

    python tools/ncu_summary.py REPORT.ncu-rep [--items N] [--title "..."] [--command "..."] [--traffic KEY ALGO_BYTES]

Prints (markdown-ready): a header with the commit the binary was built from, the command that was profiled, the
raw metrics the roofline discussion uses, the instruction mix per item (item = frame pair for the feature kernel),
and -- split at the kernel's warp-level sync points -- stall samples, instructions and shared-memory wavefronts per
phase.  With --traffic the measured DRAM bytes of the launch are merged into profiles/traffic.json under KEY, which
is what bench.py reports as `roofline.traffic`.
"""
from __future__ import annotations

import argparse
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected",
          "stall_dispatch", "stall_branch_resolving", "stall_barrier", "stall_lg", "stall_no_inst"]


def ncu(*args: str) -> str:
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=False).stdout


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--items", type=float, default=430_000.0, help="work items per launch (frame pairs: 2000 clips x 215)")
    ap.add_argument("--title", default="")
    ap.add_argument("--command", default="")
    ap.add_argument("--traffic", nargs=2, metavar=("KEY", "ALGO_BYTES"))
    ap.add_argument("--no-phases", action="store_true")
    args = ap.parse_args()

    commit = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = subprocess.run(["git", "-C", str(ROOT), "status", "--porcelain", "--", "dsp_final_b200/csrc", "include"],
                           capture_output=True, text=True).stdout.strip()
    print(f"# {args.title or Path(args.report).name}")
    print(f"# built from commit {commit}{' + uncommitted csrc changes' if dirty else ''}; capture: ncu --set full --clock-control none --import-source on")
    if args.command:
        print(f"# command: {args.command}")
    rows = list(csv.reader(ncu("-i", args.report, "--page", "raw", "--csv").splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if name_col is not None:
        print(f"# kernel: {vals[name_col]}")
    print("```")
    for w in WANT:
        if w in d:
            print(f"{w} = {d[w][0]} {d[w][1]}")
    print("```")
    if args.traffic and "dram__bytes_read.sum" in d:
        total = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        path = ROOT / "profiles" / "traffic.json"
        cur = json.loads(path.read_text()) if path.exists() else {}
        cur[args.traffic[0]] = {"dram_bytes_per_launch": int(total), "algorithmic_bytes_per_launch": int(float(args.traffic[1])),
                                "source": f"{Path(args.report).name} (commit {commit}): dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
        path.write_text(json.dumps(cur, indent=1) + "\n")
        print(f"DRAM traffic of the launch: {total / 1e9:.4f} GB against {float(args.traffic[1]) / 1e9:.4f} GB algorithmic "
              f"(ratio {total / float(args.traffic[1]):.3f}) -> profiles/traffic.json[{args.traffic[0]}]")

    src = ncu("-i", args.report, "--page", "source", "--csv", "--print-source", "sass")
    rows = list(csv.reader(src.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) >= len(hdr)]
    byop, samp, wf, wfex = (collections.Counter() for _ in range(4))
    phases: dict = collections.defaultdict(collections.Counter)
    phase, tot, tot_samples = 0, 0, 0
    for r in data:
        text = r[ix["Source"]].strip()
        toks = text.split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        n, s = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
        w = int(r[ix["L1 Wavefronts Shared"]] or 0)
        tot += n
        tot_samples += s
        byop[op] += n
        samp[op] += s
        wf[op] += w
        wfex[op] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        if "BRA.DIV" in text or "BAR.SYNC" in text:           # __syncwarp() / __syncthreads(): a phase boundary
            phase += 1
        ph = phases[phase]
        ph["inst"] += n
        ph["samples"] += s
        ph["wf"] += w
        for st in STALLS:
            if st in ix:
                ph[st] += int(r[ix[st]] or 0)
    print(f"\nwarp instructions per item: {tot / args.items:.1f}; stall samples: {tot_samples}")
    print("```")
    for op, n in byop.most_common(22):
        print(f"  {op:10s} {n / args.items:7.1f}/item  samples {samp[op]:6d}  smem wavefronts {wf[op] / args.items:6.1f}  excess {wfex[op] / args.items:6.1f}")
    print("```")
    if not args.no_phases:
        print("\nPhases (split at the warp-level sync points of the SASS, in program order):")
        print("```")
        for ph in sorted(phases):
            a = phases[ph]
            if a["samples"] == 0 and a["inst"] == 0:
                continue
            top = sorted(((a[s], s) for s in STALLS), reverse=True)[:4]
            print(f"phase {ph:2d}: {100 * a['samples'] / max(1, tot_samples):5.1f} % of samples  {a['inst'] / args.items:7.1f} inst/item  "
                  f"{a['wf'] / args.items:6.1f} smem wf/item  " + " ".join(f"{s[6:]} {100 * v / max(1, a['samples']):.0f}%" for v, s in top))
        print("```")


if __name__ == "__main__":
    main()
