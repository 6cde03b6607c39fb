"""Builds em-spec_b200/emspec/libemspec.so (the C-ABI library of include/emspec.h)
in-tree with nvcc for sm_100a.  Usage: python em-spec_b200/build_emspec.py [--force] [-v]"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "emspec", "libemspec.so")
SOURCES = ["engine.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "emspec.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """`out` / `defines`: experimental builds for A/B runs (tools/variant_check.py), e.g.
    build(out="gpurun_out/libx.so", defines=["EMS_R64_BASES=6"]); the product build takes neither."""
    if out is None and not force and not _stale():
        return OUT
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines]]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out or OUT, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libemspec.so")
    return out or OUT


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None, defines=defs))
