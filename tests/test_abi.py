"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/emspec.h declares.  No compute calls without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "emspec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ems_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(lib_built):
    import emspec
    lib = emspec.load()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/emspec.h but not exported"
    assert sorted(emspec.SYMBOLS) == syms, "binding and header disagree"


def test_abi_version_and_status_strings(lib_built):
    import emspec
    lib = emspec.load()
    assert lib.ems_abi_version() == 5
    assert lib.ems_status_str(0) == b"ok"
    for s in range(1, 6):
        assert len(lib.ems_status_str(s)) > 0


def test_default_params_are_the_settings_png_preset(lib_built):
    import emspec
    p = emspec.default_params()
    assert (p.n_fft, p.hop, p.channels) == (4096, 128, 1)
    assert abs(p.db_range - 58) < 1e-6 and abs(p.gain - 3.5) < 1e-6
    assert abs(p.low_end_boost - 3.9) < 1e-6 and p.smoothing == 0 and p.noise_gate_db == -65
    assert p.flags & emspec.FLAG_REASSIGN and p.flags & emspec.FLAG_DETERMINISTIC
    assert p.display_rows == 0 and abs(p.freq_scale - 1.0) < 1e-6
    assert p.agc_strength == 0 and abs(p.brightness - 0.44) < 1e-6        # settings.png "Brightness 44 %"


def test_invalid_arguments_rejected_before_any_cuda_call(lib_built):
    import emspec
    lib = emspec.load()
    h = ctypes.c_void_p()
    assert lib.ems_create(None, ctypes.byref(h)) == emspec.ERR_INVALID_ARG
    for bad in (dict(n_fft=1000), dict(n_fft=128), dict(n_fft=65536), dict(hop=0),
                dict(hop=8192), dict(channels=0), dict(smoothing=1.0), dict(db_range=0.0),
                dict(display_rows=1), dict(display_rows=-3), dict(freq_scale=-1.0), dict(agc_strength=1.5),
                dict(brightness=0.0), dict(brightness=1.5)):
        p = emspec.default_params()
        for k, v in bad.items():
            setattr(p, k, v)
        assert lib.ems_create(ctypes.byref(p), ctypes.byref(h)) == emspec.ERR_INVALID_ARG, bad
    assert lib.ems_destroy(None) == emspec.ERR_INVALID_ARG
    assert lib.ems_process_points(None, None, 0, None, None, None, None) == emspec.ERR_INVALID_ARG


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device the engine refuses to exist; with one it must be created."""
    import torch
    import emspec
    if torch.cuda.is_available():
        emspec.Engine().close()
    else:
        with pytest.raises(emspec.EmspecError) as ei:
            emspec.Engine()
        assert ei.value.status == emspec.ERR_CUDA


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under em-spec_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "em-spec_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import reassign_oracle" not in txt and "from reassign_oracle" not in txt, f
                assert "oracle/" not in txt.replace("oracle/reassign_oracle.py", "").replace("oracle/reassign_oracle.py::", ""), f


def _build_abi_check(tmp_path):
    """tests/c/abi_check.c compiled with gcc as plain C against include/emspec.h and linked to the library."""
    import subprocess
    import build_emspec
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(build_emspec.OUT)
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-o", exe, "-L", libdir, "-l:libemspec.so",
           f"-Wl,-rpath,{libdir}", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_plain_c_and_layout_matches_the_ctypes_mirror(lib_built, tmp_path):
    """the struct layout is checked by a C compiler (static asserts in abi_check.c), and
    the offsets it prints equal those of the hand-written ctypes mirror."""
    import subprocess
    import emspec
    exe = _build_abi_check(tmp_path)
    out = subprocess.run([exe, "layout"], capture_output=True, text=True, check=True).stdout.split("\n")
    got = dict(ln.split() for ln in out if ln.strip())
    assert int(got.pop("sizeof")) == ctypes.sizeof(emspec.Params)
    assert int(got.pop("abi")) == emspec.load().ems_abi_version()
    names = [f[0] for f in emspec.Params._fields_]
    assert sorted(got) == sorted(names)
    for n in names:
        assert int(got[n]) == getattr(emspec.Params, n).offset, n


@pytest.mark.gpu
def test_c_caller_end_to_end(lib_built, tmp_path):
    """A C program (no Python, no torch) drives ems_create / ems_process_host / ems_destroy."""
    import subprocess
    exe = _build_abi_check(tmp_path)
    res = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "bad_columns 0" in res.stdout


def test_builtin_colour_maps_match_the_oracle_bit_for_bit(lib_built):
    """"Multiple Color Maps" (README.md:15,45): host-side integer work, callable without a GPU; every
    built-in table equals the oracle's restatement, is opaque, starts dark and ends bright."""
    import numpy as np
    import emspec
    import reassign_oracle as orc
    names = emspec.colormap_names()
    assert names == list(orc.COLORMAPS) and len(names) == emspec.load().ems_colormap_count()
    for i, name in enumerate(names):
        lut = emspec.builtin_colormap(i)
        assert lut.dtype == np.uint32 and lut.shape == (256,)
        assert (lut == orc.builtin_colormap(name)).all(), name
        assert (emspec.builtin_colormap(name) == lut).all()
        assert ((lut >> 24) == 0xFF).all()
        luma = (lut & 0xFF).astype(int) * 2 + ((lut >> 8) & 0xFF) * 5 + ((lut >> 16) & 0xFF)
        assert luma[0] < 300 and luma[255] > 1400 and luma[255] == luma.max(), name
    lib = emspec.load()
    buf = np.zeros(256, np.uint32)
    for bad in (-1, len(names)):
        assert lib.ems_colormap_name(bad) is None
        assert lib.ems_colormap_builtin(bad, ctypes.c_void_p(buf.ctypes.data)) == emspec.ERR_INVALID_ARG
    assert lib.ems_colormap_builtin(0, None) == emspec.ERR_INVALID_ARG
    assert lib.ems_cursor_info(None, 0.0, 0.0, None) == emspec.ERR_INVALID_ARG
