"""ctypes binding of include/emspec.h — used by the parity tests and bench.py.

The product is libemspec.so (C-ABI, hand-written CUDA for sm_100a).  This module only
marshals pointers: torch is used for device memory and streams, never for compute.
There is no CPU path: if the library is missing or no CUDA device is present every
entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EMS_LIB_PATH: A/B runs of two builds of the same C-ABI (tools/ab_kernel.py); never a CPU path
LIB_PATH = os.environ.get("EMS_LIB_PATH") or os.path.join(_HERE, "libemspec.so")

FLAG_REASSIGN = 1
FLAG_DETERMINISTIC = 2
FLAG_SYNC = 4
FLAG_BOUNDED_SCRATCH = 8
FLAG_SORTED_SCATTER = 16
STAGE_POINTS, STAGE_SCATTER, STAGE_POST = 0, 1, 2

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_STATE = range(6)


class EmspecError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"emspec status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    """ems_params (include/emspec.h); defaults = assets/settings.png "Default" preset."""
    _fields_ = [
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("sample_rate", C.c_float),
        ("channels", C.c_int32), ("db_range", C.c_float), ("gain", C.c_float),
        ("low_end_boost", C.c_float), ("smoothing", C.c_float),
        ("noise_gate_db", C.c_float), ("flags", C.c_uint32),
        ("display_rows", C.c_int32), ("freq_scale", C.c_float), ("agc_strength", C.c_float),
        ("brightness", C.c_float),
    ]


class Cursor(C.Structure):
    """ems_cursor (include/emspec.h): what lies under an output cell."""
    _fields_ = [("time_s", C.c_double), ("freq_hz", C.c_double), ("midi_note", C.c_int32),
                ("cents", C.c_float), ("name", C.c_char * 8)]


# name -> (restype, argtypes); every symbol include/emspec.h declares.
_VP, _FP, _U8P = C.c_void_p, C.c_void_p, C.c_void_p
_SIG = {
    "ems_abi_version": (C.c_int, []),
    "ems_status_str": (C.c_char_p, [C.c_int]),
    "ems_last_error": (C.c_char_p, [_VP]),
    "ems_default_params": (C.c_int, [C.POINTER(Params)]),
    "ems_create": (C.c_int, [C.POINTER(Params), C.POINTER(_VP)]),
    "ems_destroy": (C.c_int, [_VP]),
    "ems_update_display": (C.c_int, [_VP, C.POINTER(Params)]),
    "ems_set_stream": (C.c_int, [_VP, _VP]),
    "ems_get_stream": (C.c_int, [_VP, C.POINTER(_VP)]),
    "ems_synchronize": (C.c_int, [_VP]),
    "ems_output_rows": (C.c_int, [_VP, C.POINTER(C.c_size_t)]),
    "ems_frame_count": (C.c_int, [_VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ems_cursor_info": (C.c_int, [_VP, C.c_double, C.c_double, C.POINTER(Cursor)]),
    "ems_hz_to_row": (C.c_int, [_VP, C.c_double, C.POINTER(C.c_double)]),
    "ems_colormap_count": (C.c_int, []),
    "ems_colormap_name": (C.c_char_p, [C.c_int]),
    "ems_colormap_builtin": (C.c_int, [C.c_int, _VP]),
    "ems_process_points": (C.c_int, [_VP, _FP, C.c_size_t, _FP, _FP, _FP, C.POINTER(C.c_size_t)]),
    "ems_process_grid": (C.c_int, [_VP, _FP, C.c_size_t, _FP, _U8P, C.POINTER(C.c_size_t)]),
    "ems_scatter_points": (C.c_int, [_VP, _FP, _FP, _FP, C.c_size_t, _FP, _U8P]),
    "ems_process_host": (C.c_int, [_VP, _FP, C.c_size_t, _FP, _U8P, C.POINTER(C.c_size_t)]),
    "ems_process_host_i16": (C.c_int, [_VP, _FP, C.c_size_t, _FP, _U8P, C.POINTER(C.c_size_t)]),
    "ems_process_host_i24": (C.c_int, [_VP, _U8P, C.c_size_t, _FP, _U8P, C.POINTER(C.c_size_t)]),
    "ems_scratch_bytes": (C.c_int, [_VP, C.POINTER(C.c_size_t)]),
    "ems_image_summary": (C.c_int, [_VP, _U8P, C.c_size_t, _VP]),
    "ems_colorize": (C.c_int, [_VP, _U8P, C.c_size_t, _VP, _VP]),
    "ems_stage_ms": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_float)]),
    "ems_launch_count": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "ems_stream_push": (C.c_int, [_VP, _FP, _U8P, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "ems_stream_push_i16": (C.c_int, [_VP, _VP, _U8P, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "ems_stream_push_i24": (C.c_int, [_VP, _VP, _U8P, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "ems_stream_set_colormap": (C.c_int, [_VP, _VP]),
    "ems_stream_column_rgba": (C.c_int, [_VP, _VP]),
    "ems_stream_reset": (C.c_int, [_VP]),
    "ems_stream_state_size": (C.c_int, [_VP, C.POINTER(C.c_size_t)]),
    "ems_stream_save": (C.c_int, [_VP, _VP, C.c_size_t]),
    "ems_stream_load": (C.c_int, [_VP, _VP, C.c_size_t]),
}
SYMBOLS = tuple(_SIG)

_lib = None


def load() -> C.CDLL:
    """Loads libemspec.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python em-spec_b200/build_emspec.py` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def default_params() -> Params:
    p = Params()
    load().ems_default_params(C.byref(p))
    return p


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def colormap_names() -> list:
    """Names of the built-in colour maps, by id."""
    lib = load()
    return [lib.ems_colormap_name(i).decode() for i in range(lib.ems_colormap_count())]


def builtin_colormap(which):
    """256 uint32 pixels (0xAABBGGRR) of a built-in colour map, by id or name."""
    import numpy as np
    lib = load()
    idx = colormap_names().index(which) if isinstance(which, str) else int(which)
    lut = np.zeros(256, np.uint32)
    st = lib.ems_colormap_builtin(idx, C.c_void_p(lut.ctypes.data))
    if st != OK:
        raise EmspecError(st, lib.ems_status_str(st).decode())
    return lut


class Engine:
    """One ems_handle.  Tensors in / out are torch CUDA tensors (device memory only)."""

    def __init__(self, **kw):
        self.lib = load()
        self.params = default_params()
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self.params, k, v)
        self.h = C.c_void_p()
        st = self.lib.ems_create(C.byref(self.params), C.byref(self.h))
        if st != OK:
            self.h = C.c_void_p()
            raise EmspecError(st, self.lib.ems_status_str(st).decode())

    # ------------------------------------------------------------------ plumbing
    def _check(self, st: int):
        if st != OK:
            raise EmspecError(st, f"{self.lib.ems_status_str(st).decode()}: "
                                  f"{self.lib.ems_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.ems_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_bins(self) -> int:
        return self.params.n_fft // 2 + 1

    @property
    def n_rows(self) -> int:
        """Output rows per column of grid / index / streamed columns (n_bins or display_rows)."""
        n = C.c_size_t()
        self._check(self.lib.ems_output_rows(self.h, C.byref(n)))
        return n.value

    def frame_count(self, n_samples: int) -> int:
        n = C.c_size_t()
        self._check(self.lib.ems_frame_count(self.h, n_samples, C.byref(n)))
        return n.value

    def cursor_info(self, column: float, row: float) -> dict:
        """Time, frequency and nearest note under output cell (column, row)."""
        c = Cursor()
        self._check(self.lib.ems_cursor_info(self.h, float(column), float(row), C.byref(c)))
        return {"time_s": c.time_s, "freq_hz": c.freq_hz, "midi_note": c.midi_note,
                "cents": c.cents, "name": c.name.decode()}

    def hz_to_row(self, freq_hz: float) -> float:
        """Fractional output row of a frequency (axis ticks, note grid lines)."""
        r = C.c_double()
        self._check(self.lib.ems_hz_to_row(self.h, float(freq_hz), C.byref(r)))
        return r.value

    def use_torch_stream(self):
        """Run on torch's current stream so torch.cuda.Event timing and ordering apply."""
        import torch
        self._check(self.lib.ems_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def synchronize(self):
        self._check(self.lib.ems_synchronize(self.h))

    def update_display(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._check(self.lib.ems_update_display(self.h, C.byref(self.params)))

    def stage_ms(self, stage: int) -> float:
        ms = C.c_float()
        self._check(self.lib.ems_stage_ms(self.h, stage, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.ems_launch_count(self.h, C.byref(n)))
        return n.value

    def _pcm(self, pcm):
        import torch
        if pcm.dim() == 1:
            pcm = pcm[None, :]
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous()
        assert pcm.shape[0] == self.params.channels
        return pcm

    # ------------------------------------------------------------------ offline
    def process_points(self, pcm, out=None):
        """-> (dt_cols, dk_bins, energy) fp32 CUDA tensors [channels][F][B]."""
        import torch
        pcm = self._pcm(pcm)
        S = pcm.shape[1]
        F = self.frame_count(S)
        shape = (self.params.channels, F, self.n_bins)
        if out is None:
            out = tuple(torch.empty(shape, dtype=torch.float32, device=pcm.device) for _ in range(3))
        n = C.c_size_t()
        self._check(self.lib.ems_process_points(self.h, _ptr(pcm), S, _ptr(out[0]), _ptr(out[1]),
                                                _ptr(out[2]), C.byref(n)))
        assert n.value == F
        return out

    def process_grid(self, pcm, want_grid=True, want_index=True, out=None):
        """-> (grid fp32 | None, index u8 | None), CUDA tensors [channels][F][B]."""
        import torch
        pcm = self._pcm(pcm)
        S = pcm.shape[1]
        F = self.frame_count(S)
        shape = (self.params.channels, F, self.n_rows)
        if out is None:
            grid = torch.empty(shape, dtype=torch.float32, device=pcm.device) if want_grid else None
            idx = torch.empty(shape, dtype=torch.uint8, device=pcm.device) if want_index else None
        else:
            grid, idx = out
        n = C.c_size_t()
        self._check(self.lib.ems_process_grid(self.h, _ptr(pcm), S, _ptr(grid), _ptr(idx), C.byref(n)))
        return grid, idx

    def scatter_points(self, dt, dk, en, want_grid=True, want_index=True):
        import torch
        F = en.shape[-2]
        shape = (self.params.channels, F, self.n_rows)
        grid = torch.empty(shape, dtype=torch.float32, device=en.device) if want_grid else None
        idx = torch.empty(shape, dtype=torch.uint8, device=en.device) if want_index else None
        self._check(self.lib.ems_scatter_points(self.h, _ptr(dt), _ptr(dk), _ptr(en), F,
                                                _ptr(grid), _ptr(idx)))
        return grid, idx

    def process_host(self, pcm_host, want_grid=False, index_out=None, grid_out=None):
        """Host (CPU, ideally pinned) fp32 tensor [channels][S] -> (grid | None, index) CPU tensors."""
        import torch
        if pcm_host.dim() == 1:
            pcm_host = pcm_host[None, :]
        assert (not pcm_host.is_cuda) and pcm_host.dtype == torch.float32 and pcm_host.is_contiguous()
        S = pcm_host.shape[1]
        F = self.frame_count(S)
        shape = (self.params.channels, F, self.n_rows)
        pin = torch.cuda.is_available()
        if index_out is None:
            index_out = torch.empty(shape, dtype=torch.uint8, pin_memory=pin)
        if want_grid and grid_out is None:
            grid_out = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
        n = C.c_size_t()
        self._check(self.lib.ems_process_host(self.h, _ptr(pcm_host), S, _ptr(grid_out),
                                              _ptr(index_out), C.byref(n)))
        return grid_out, index_out

    def process_host_i16(self, pcm_i16, want_grid=False, index_out=None, grid_out=None):
        """Host int16 tensor [S][channels] (interleaved) -> (grid | None, index) CPU tensors."""
        import torch
        if pcm_i16.dim() == 1:
            pcm_i16 = pcm_i16[:, None]
        assert (not pcm_i16.is_cuda) and pcm_i16.dtype == torch.int16 and pcm_i16.is_contiguous()
        assert pcm_i16.shape[1] == self.params.channels
        S = pcm_i16.shape[0]
        F = self.frame_count(S)
        shape = (self.params.channels, F, self.n_rows)
        pin = torch.cuda.is_available()
        if index_out is None:
            index_out = torch.empty(shape, dtype=torch.uint8, pin_memory=pin)
        if want_grid and grid_out is None:
            grid_out = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
        n = C.c_size_t()
        self._check(self.lib.ems_process_host_i16(self.h, _ptr(pcm_i16), S, _ptr(grid_out),
                                                  _ptr(index_out), C.byref(n)))
        return grid_out, index_out

    def process_host_i24(self, pcm_i24, want_grid=False, index_out=None, grid_out=None):
        """Host uint8 tensor [S][channels][3] (packed little-endian int24, interleaved) ->
        (grid | None, index) CPU tensors."""
        import torch
        assert (not pcm_i24.is_cuda) and pcm_i24.dtype == torch.uint8 and pcm_i24.is_contiguous()
        assert pcm_i24.dim() == 3 and pcm_i24.shape[1] == self.params.channels and pcm_i24.shape[2] == 3
        S = pcm_i24.shape[0]
        F = self.frame_count(S)
        shape = (self.params.channels, F, self.n_rows)
        pin = torch.cuda.is_available()
        if index_out is None:
            index_out = torch.empty(shape, dtype=torch.uint8, pin_memory=pin)
        if want_grid and grid_out is None:
            grid_out = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
        n = C.c_size_t()
        self._check(self.lib.ems_process_host_i24(self.h, _ptr(pcm_i24), S, _ptr(grid_out),
                                                  _ptr(index_out), C.byref(n)))
        return grid_out, index_out

    def image_summary(self, index, out=None):
        """index: CUDA u8 [channels][F][R] -> CUDA int64 [channels][2]: (sum of bytes, position-weighted sum)."""
        import torch
        assert index.is_cuda and index.dtype == torch.uint8 and index.is_contiguous() and index.shape[0] == self.params.channels
        if out is None:
            out = torch.empty((self.params.channels, 2), dtype=torch.int64, device=index.device)
        self._check(self.lib.ems_image_summary(self.h, _ptr(index), index.shape[1], _ptr(out)))
        return out

    def scratch_bytes(self) -> int:
        n = C.c_size_t()
        self._check(self.lib.ems_scratch_bytes(self.h, C.byref(n)))
        return n.value

    def colorize(self, index, lut_rgba):
        """index: CUDA u8 tensor (any shape); lut_rgba: 256 uint32 (numpy / CPU tensor, 0xAABBGGRR)
        -> CUDA int32 tensor of the same shape holding the packed pixels."""
        import numpy as np
        import torch
        lut = np.ascontiguousarray(np.asarray(lut_rgba, dtype=np.uint32))
        assert lut.size == 256 and index.is_cuda and index.dtype == torch.uint8 and index.is_contiguous()
        out = torch.empty(index.shape, dtype=torch.int32, device=index.device)
        self._check(self.lib.ems_colorize(self.h, _ptr(index), index.numel(),
                                          C.c_void_p(lut.ctypes.data), _ptr(out)))
        return out

    # ------------------------------------------------------------------ streaming
    def stream_reset(self):
        self._check(self.lib.ems_stream_reset(self.h))

    def stream_save(self) -> bytes:
        """Checkpoint of the streaming state (ring, rolling columns, smoothing / AGC state)."""
        n = C.c_size_t()
        self._check(self.lib.ems_stream_state_size(self.h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._check(self.lib.ems_stream_save(self.h, buf, n.value))
        return buf.raw

    def stream_load(self, blob: bytes):
        self._check(self.lib.ems_stream_load(self.h, blob, len(blob)))

    def stream_set_colormap(self, lut_rgba):
        """lut_rgba: 256 uint32 (0xAABBGGRR) or None to switch the streaming colour map off."""
        import numpy as np
        if lut_rgba is None:
            self._check(self.lib.ems_stream_set_colormap(self.h, None))
            return
        lut = np.ascontiguousarray(np.asarray(lut_rgba, dtype=np.uint32))
        assert lut.size == 256
        self._check(self.lib.ems_stream_set_colormap(self.h, C.c_void_p(lut.ctypes.data)))

    def stream_column_rgba(self, rgba_host):
        """rgba_host: CPU int32 [channels][n_rows]; receives the pixels of the column the last push delivered."""
        self._check(self.lib.ems_stream_column_rgba(self.h, _ptr(rgba_host)))
        return rgba_host

    def stream_push(self, pcm_host, column_host):
        """pcm_host: CPU fp32, int16 or uint8 (packed int24: 3 bytes per sample) [hop*channels] interleaved, capture
        formats; column_host: CPU u8 [channels][n_rows].  -> (ready: bool, column_index: int)"""
        import torch
        ready, idx = C.c_int(0), C.c_int64(-1)
        fn = {torch.int16: self.lib.ems_stream_push_i16, torch.uint8: self.lib.ems_stream_push_i24}.get(pcm_host.dtype, self.lib.ems_stream_push)
        if pcm_host.dtype == torch.uint8:
            assert pcm_host.numel() == 3 * self.params.hop * self.params.channels
        self._check(fn(self.h, _ptr(pcm_host), _ptr(column_host), C.byref(ready), C.byref(idx)))
        return bool(ready.value), idx.value
