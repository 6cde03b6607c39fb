"""fp32 emulation of the CUDA kernels' algebra in NumPy (complex64 FFT, float32 operators): predicts,
on a machine without a GPU, whether a signal meets the parity tolerances before GPU time is spent.
Not the product and not the oracle: one FFT Z = FFT(x + j x th'), untangle, Hann / dh stencils,
Auger-Flandrin operators and the drop rule, as in em-spec_b200/csrc/stft_r16.cuh::bin_tail."""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.fft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import reassign_oracle as orc  # noqa: E402


def points_fp32(x: np.ndarray, prm: orc.Params):
    N, H = prm.n_fft, prm.hop
    F = orc.frame_count(len(x), N, H)
    n = np.arange(N, dtype=np.float64)
    thw = ((n - N / 2) * (0.5 - 0.5 * np.cos(2 * np.pi * n / N)) * (2.0 / N)).astype(np.float32)
    idx = (np.arange(F)[:, None] * H) + np.arange(N)[None, :]
    fr = x.astype(np.float32)[idx]
    z = (fr + 1j * (fr * thw)).astype(np.complex64)
    Z = scipy.fft.fft(z, axis=1)
    assert Z.dtype == np.complex64
    k = np.arange(N // 2 + 1)
    Zk, Zn = Z[:, k], np.conj(Z[:, (N - k) % N])
    X2 = (Zk + Zn).astype(np.complex64)                      # 2 X[k]
    T2 = ((Zk - Zn) * np.complex64(-1j)).astype(np.complex64)  # 2 X_th'[k]
    Xm = np.concatenate([np.conj(X2[:, 1:2]), X2[:, :-1]], axis=1)
    Xp = np.concatenate([X2[:, 1:], np.conj(X2[:, -2:-1])], axis=1)
    A4 = (X2 - np.float32(0.5) * (Xm + Xp)).astype(np.complex64)          # 4 X_h
    p4 = (A4.real * A4.real + A4.imag * A4.imag).astype(np.float32)
    e = (p4 * np.float32(1.0 / (N * N))).astype(np.float32)
    d = (Xm - Xp).astype(np.complex64)
    D2 = (np.float32(0.5) * d.imag - 1j * np.float32(0.5) * d.real).astype(np.complex64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / p4).astype(np.float32)
        dts = ((T2.real * A4.real + T2.imag * A4.imag) * inv * np.float32(N)).astype(np.float32)
        dk = (-(D2.imag * A4.real - D2.real * A4.imag) * inv).astype(np.float32)
    dtc = (dts * np.float32(1.0 / H)).astype(np.float32)
    rc = np.rint(dtc)
    m = np.arange(F, dtype=np.float64)[:, None]
    rowf = k[None, :] + np.rint(dk)
    ok = (e > np.float32(prm.gate_lin)) & (np.abs(dts) <= N / 2) & (rowf >= 0) & (rowf <= N / 2) \
        & (m + rc >= 0) & (m + rc <= F - 1)
    if not (prm.flags & orc.FLAG_REASSIGN):
        ok = e > np.float32(prm.gate_lin)
        dtc = np.zeros_like(dtc); dk = np.zeros_like(dk)
    return np.where(ok, dtc, 0).astype(np.float32), np.where(ok, dk, 0).astype(np.float32), np.where(ok, e, 0).astype(np.float32)


def grid_index_fp32(x: np.ndarray, prm: orc.Params):
    dt, dk, e = points_fp32(x, prm)
    grid = orc.scatter_grid(dt.astype(np.float64), dk.astype(np.float64), e.astype(np.float64), prm)
    return grid.astype(np.float32), orc.postpass(grid.astype(np.float32).astype(np.float64), prm)


if __name__ == "__main__":
    from parity_util import check_grid_dense, check_index, check_points
    for n_fft, hop in ((4096, 128), (8192, 256)):
        x = orc.synth_music(int(1.5 * 48000), 48000.0, seed=31)
        prm = orc.Params(n_fft=n_fft, hop=hop)
        print(n_fft, hop, check_points(points_fp32(x, prm), x, prm))
        g, i = grid_index_fp32(x, prm)
        err, grid_o, amb = check_grid_dense(g, x, prm)
        print("  grid", err, "ambiguous", amb.mean(), "index off-by-one", check_index(i, grid_o, prm, x, max_excused=1.0))
