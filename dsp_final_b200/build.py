"""Builds the native libraries in-tree with nvcc for sm_100a.

  dsp_final_b200/_native/libdspx.so      the product: C ABI of include/dspx.h
  dsp_final_b200/_native/libdspx_emu.so  test-only host replay of the kernels' phase functions

The .so files are git-ignored but travel to the GPU box with the gpurun
snapshot, so nothing is JIT-compiled there.  `python -m dsp_final_b200.build`
rebuilds when a source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "_native"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libdspx.so cannot be built (there is no CPU fallback)")


def _stale(target: Path) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dspx.h"]
    return any(s.stat().st_mtime > t for s in srcs)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    lib = OUT / "libdspx.so"
    if force or _stale(lib):
        # NOT -split-compile: it splits the module before optimisation and the 96-register feature kernels come out with
        # local-memory spills and rolled loops (2344 instead of 2648 SASS instructions, 26 LDL/STL; -2 % measured)
        cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "128", "-Xcompiler", "-fPIC", "-shared",
               "-o", str(lib), str(CSRC / "dspx.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return lib


def build_emu(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    lib = OUT / "libdspx_emu.so"
    if force or _stale(lib):
        cmd = [_nvcc(), *ARCH, "-O2", "-std=c++17", "-diag-suppress", "128", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
               "-o", str(lib), str(CSRC / "emu.cu")]
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_native(force=force, verbose="-v" in sys.argv))
    print(build_emu(force=force))
