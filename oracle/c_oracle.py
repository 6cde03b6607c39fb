"""ctypes binding of oracle/reassign_oracle.c (the plain-C, multi-threaded restatement of the stand-in
oracle).  TEST INFRASTRUCTURE ONLY — same rule as oracle/reassign_oracle.py: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

The functions mirror reassign_oracle.py (same names, same arrays) with a `threads` argument.
The library is built into oracle/_build/ (git-ignored, travels to the GPU box): an AVX2 + FMA build
used when /proc/cpuinfo shows both, and a portable one otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import reassign_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "reassign_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIBS = {"avx2": os.path.join(OUT_DIR, "libreassign_oracle_avx2.so"),
        "generic": os.path.join(OUT_DIR, "libreassign_oracle.so")}
# -fno-math-errno / -fno-trapping-math let gcc inline rint and vectorise; neither changes a value
_FLAGS = {"avx2": ["-O3", "-mavx2", "-mfma", "-fno-math-errno", "-fno-trapping-math"],
          "generic": ["-O2", "-fno-math-errno", "-fno-trapping-math"]}


class CParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32),
                ("sample_rate", C.c_double), ("db_range", C.c_double), ("gain", C.c_double),
                ("low_end_boost", C.c_double), ("smoothing", C.c_double), ("noise_gate_db", C.c_double),
                ("reassign", C.c_int32), ("display_rows", C.c_int32),
                ("freq_scale", C.c_double), ("agc_strength", C.c_double), ("brightness", C.c_double)]


def build(force: bool = False) -> dict:
    """gcc -std=c11 -shared of reassign_oracle.c, both variants; rebuilt when the source is newer."""
    os.makedirs(OUT_DIR, exist_ok=True)
    for kind, out in LIBS.items():
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
            continue
        cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-pthread",
               *_FLAGS[kind], SRC, "-o", out, "-lm"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    return dict(LIBS)


def _cpu_has(*flags) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    have = set(ln.split(":", 1)[1].split())
                    return all(x in have for x in flags)
    except OSError:
        pass
    return False


_lib = None
_kind = None


def load() -> C.CDLL:
    global _lib, _kind
    if _lib is None:
        _kind = "avx2" if _cpu_has("avx2", "fma") else "generic"
        path = LIBS[_kind]
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.orc_abi.restype = C.c_int
        assert lib.orc_abi() == 1
        lib.orc_frame_count.restype = C.c_int64
        lib.orc_frame_count.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        P, VP = C.POINTER(CParams), C.c_void_p
        lib.orc_points.argtypes = [VP, C.c_int64, P, VP, VP, VP, VP, C.c_int]
        lib.orc_scatter.argtypes = [VP, VP, VP, C.c_int64, P, VP, C.c_int]
        lib.orc_postpass.argtypes = [VP, C.c_int64, P, VP, C.c_int]
        lib.orc_process.argtypes = [VP, C.c_int64, P, VP, VP, C.c_int]
        for fn in (lib.orc_points, lib.orc_scatter, lib.orc_postpass, lib.orc_process):
            fn.restype = C.c_int
        _lib = lib
    return _lib


def build_kind() -> str:
    load()
    return _kind


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cp(prm: orc.Params) -> CParams:
    return CParams(prm.n_fft, prm.hop, prm.sample_rate, prm.db_range, prm.gain, prm.low_end_boost,
                   prm.smoothing, prm.noise_gate_db, 1 if prm.flags & orc.FLAG_REASSIGN else 0,
                   prm.display_rows, prm.freq_scale, prm.agc_strength, prm.brightness)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what}: status {rc}")


def reassign_points(x, prm: orc.Params, threads: int = 0, return_raw: bool = False):
    """reassign_oracle.reassign_points in C: (dt_cols, dk_bins, energy[, raw]) float64 [F][B]."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 1
    F = orc.frame_count(x.shape[0], prm.n_fft, prm.hop)
    B = prm.n_bins
    dcol, dbin, en = (np.empty((F, B), np.float64) for _ in range(3))
    raw = np.empty((F, B), np.float64) if return_raw else None
    cp = _cp(prm)
    _check(lib.orc_points(_ptr(x), x.shape[0], C.byref(cp), _ptr(dcol), _ptr(dbin), _ptr(en),
                          _ptr(raw) if return_raw else None, threads or host_threads()), "orc_points")
    return (dcol, dbin, en, raw) if return_raw else (dcol, dbin, en)


def scatter_grid(dcol, dbin, energy, prm: orc.Params, threads: int = 0) -> np.ndarray:
    lib = load()
    dcol, dbin, energy = (np.ascontiguousarray(a, dtype=np.float64) for a in (dcol, dbin, energy))
    F = energy.shape[0]
    grid = np.empty((F, prm.n_rows), np.float64)
    cp = _cp(prm)
    _check(lib.orc_scatter(_ptr(dcol), _ptr(dbin), _ptr(energy), F, C.byref(cp), _ptr(grid),
                           threads or host_threads()), "orc_scatter")
    return grid


def postpass(grid, prm: orc.Params, threads: int = 0) -> np.ndarray:
    lib = load()
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    idx = np.empty(grid.shape, np.uint8)
    cp = _cp(prm)
    _check(lib.orc_postpass(_ptr(grid), grid.shape[0], C.byref(cp), _ptr(idx), threads or host_threads()), "orc_postpass")
    return idx


def process(x, prm: orc.Params, threads: int = 0, out=None):
    """reassign_oracle.process in C: (grid float64 [F][R], index u8 [F][R]).  out: optional (grid, index)
    arrays with at least F rows to write into (a caller walking a stream in slices reuses them)."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    F = orc.frame_count(x.shape[0], prm.n_fft, prm.hop)
    if out is not None:
        grid, idx = out[0][:F], out[1][:F]
        assert grid.shape == idx.shape == (F, prm.n_rows) and grid.dtype == np.float64 and idx.dtype == np.uint8
        assert grid.flags.c_contiguous and idx.flags.c_contiguous
    else:
        grid = np.empty((F, prm.n_rows), np.float64)
        idx = np.empty((F, prm.n_rows), np.uint8)
    cp = _cp(prm)
    _check(lib.orc_process(_ptr(x), x.shape[0], C.byref(cp), _ptr(grid), _ptr(idx), threads or host_threads()), "orc_process")
    return grid, idx


if __name__ == "__main__":
    print(build(force=True))
