#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` capture of bench.py into profiles/rNN_launches.csv.

    python tools/launch_list.py RAW.csv OUT.csv --command "..." [--bench-json BENCH.json]

Keeps every dspx:: launch in order (the torch kernels that generate the synthetic inputs are counted and dropped),
writes a header with the commit the library was built from and the profiled command, checks that every kernel
name in the list exists in the shipped libdspx.so (cuobjdump), and prints the dominant kernel's share of the step
next to bench.py's own ms_per_step (ncu times are cold-cache and serialised: shares must agree, not absolutes).
"""
from __future__ import annotations

import argparse
import collections
import csv
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("out")
    ap.add_argument("--command", required=True)
    ap.add_argument("--bench-json")
    args = ap.parse_args()
    lines = Path(args.raw).read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.reader(lines[start:]))
    hdr, data = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
    name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ours = [r for r in data if "dspx::" in r[name_i]]
    commit = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "dsp_final_b200" / "_native" / "libdspx.so")], capture_output=True, text=True).stdout
    mangled = set(re.findall(r"Function : (\S+)", sass))
    demangled = subprocess.run(["c++filt"], input="\n".join(sorted(mangled)), capture_output=True, text=True).stdout
    shipped = {re.sub(r"\(.*", "", d).replace("void ", "").replace("dspx::", "").strip() for d in demangled.splitlines()}
    missing = set()
    by_kernel = collections.defaultdict(list)
    for r in ours:
        base = re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("dspx::", "").strip()
        by_kernel[base].append(float(r[val_i].replace(",", "")))
        stem = base.split("<")[0]
        if not any(s.split("<")[0] == stem for s in shipped):
            missing.add(base)
    with open(args.out, "w", newline="") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none --csv {args.command}\n")
        f.write(f"# libdspx.so built from commit {commit}; {len(data)} launches in total, the {len(data) - len(ours)} torch kernels that generate "
                f"the synthetic inputs are omitted, every dspx:: launch is listed in order\n")
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(hdr)
        w.writerows(ours)
    print(f"{len(ours)} dspx launches of {len(data)}; kernels not found in the shipped library: {sorted(missing) or 'none'}")
    total = sum(sum(v) for v in by_kernel.values())
    for k, v in sorted(by_kernel.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:60s} x{len(v):4d}  mean {sum(v) / len(v) / 1e6:9.4f} ms  share of dspx time {100 * sum(v) / total:5.1f} %")
    if args.bench_json:
        line = [l for l in Path(args.bench_json).read_text().splitlines() if l.startswith("{")][-1]
        b = json.loads(line)
        dom = max(by_kernel.items(), key=lambda kv: len(kv[1]) if "feat_warp8" in kv[0] else 0)
        # the same kernel also runs on the small chunks of the host pipeline (e2e leg): the timed steps are the
        # full-size launches, i.e. the ones within 20 % of the longest
        full = [v for v in dom[1] if v >= 0.8 * max(dom[1])]
        mean_ms = sum(full) / len(full) / 1e6
        print(f"dominant kernel {dom[0]}: {len(full)} full-size launches (warm-up + timed steps), ncu mean {mean_ms:.4f} ms vs "
              f"bench.py ms_per_step {b['ms_per_step']:.4f} ms ({100 * (mean_ms / b['ms_per_step'] - 1):+.1f} %); "
              f"{len(dom[1]) - len(full)} chunk launches of the host pipeline")


if __name__ == "__main__":
    main()
