"""The plain-C restatement of the oracle (oracle/reassign_oracle.c) against the NumPy one.

Both are float64 and follow the same definitions, but share no code: NumPy runs three scipy
rfft per frame, the C file packs the real signals into complex radix-2/4 FFTs of its own.  They
must agree to float64 round-off; every decision (kept / dropped, destination cell, colour index)
must be the same except on values that sit on a threshold to within that round-off."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import reassign_oracle as orc  # noqa: E402
import c_oracle as co  # noqa: E402

SR = 48000
GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = [
    (4096, 128, {}),
    (2048, 512, dict(smoothing=0.4)),                                            # configs[0] geometry
    (1024, 100, dict(display_rows=300, freq_scale=0.8, agc_strength=0.6, brightness=0.7)),
    (8192, 256, dict(flags=orc.FLAG_DETERMINISTIC)),                             # plain mode, odd stage count
    (256, 64, dict(noise_gate_db=-120.0)),
    (512, 512, dict(low_end_boost=1.0, gain=1.0, db_range=90.0)),
]


def _points_agree(a, b):
    kept_a, kept_b = a[2] > 0, b[2] > 0
    assert (kept_a != kept_b).sum() <= 1e-6 * max(1, kept_a.sum())
    k = kept_a & kept_b
    if k.any():
        assert np.abs(a[0] - b[0])[k].max() < 1e-7 and np.abs(a[1] - b[1])[k].max() < 1e-7
        assert (np.abs(a[2] - b[2])[k] / a[2][k]).max() < 1e-8


@pytest.mark.parametrize("n_fft,hop,kw", CASES)
def test_c_oracle_matches_numpy_oracle(n_fft, hop, kw):
    prm = orc.Params(n_fft=n_fft, hop=hop, **kw)
    x = orc.synth_signal(SR, SR, seed=n_fft + hop)
    a = orc.reassign_points(x, prm, return_raw=True)
    b = co.reassign_points(x, prm, return_raw=True)
    _points_agree(a, b)
    assert (np.abs(a[3] - b[3]) <= 1e-9 * a[3].max()).all()          # un-gated energy, relative to the peak
    # the same points scatter to the same grid, bit for bit (same order of additions)
    ga, gb = orc.scatter_grid(*a[:3], prm), co.scatter_grid(*a[:3], prm)
    assert np.array_equal(ga, gb)
    # the same grid shapes to the same image (a value on a rounding boundary may differ by one step)
    ia, ib = orc.postpass(ga, prm), co.postpass(ga, prm)
    d = np.abs(ia.astype(int) - ib.astype(int))
    assert d.max() <= 1 and (d > 0).sum() <= 1e-6 * d.size
    # whole path
    g2, i2 = co.process(x, prm)
    assert np.linalg.norm(g2 - ga) <= 1e-9 * np.linalg.norm(ga)
    d = np.abs(i2.astype(int) - ia.astype(int))
    assert (d > 0).sum() <= 1e-5 * d.size


def test_thread_count_does_not_change_a_bit():
    prm = orc.Params(n_fft=1024, hop=64, smoothing=0.3, agc_strength=0.5)
    x = orc.synth_music(SR // 2, SR, seed=2)
    ref = co.reassign_points(x, prm, threads=1)
    g_ref = co.scatter_grid(*ref, prm, threads=1)
    i_ref = co.postpass(g_ref, prm, threads=1)
    for th in (2, 3, 7, 16):
        pts = co.reassign_points(x, prm, threads=th)
        assert all(np.array_equal(u, v) for u, v in zip(ref, pts)), th
        assert np.array_equal(co.scatter_grid(*ref, prm, threads=th), g_ref), th
        assert np.array_equal(co.postpass(g_ref, prm, threads=th), i_ref), th
    g, i = co.process(x, prm, threads=5)
    assert np.array_equal(g, g_ref) and np.array_equal(i, i_ref)


def test_edges_short_streams_and_odd_frame_counts():
    lib = co.load()
    assert lib.orc_frame_count(1023, 1024, 256) == 0 and lib.orc_frame_count(1024, 1024, 256) == 1
    assert lib.orc_frame_count(172800000, 4096, 128) == 1349969
    prm = orc.Params(n_fft=1024, hop=256)
    for S in (0, 1000, 1024, 1279, 1280, 1024 + 2 * 256, 1024 + 6 * 256):        # F = 0, 0, 1, 1, 2, 3, 7
        x = orc.synth_signal(max(S, 1), SR, seed=S)[:S]
        a, b = orc.reassign_points(x, prm), co.reassign_points(x, prm, threads=4)
        assert a[0].shape == b[0].shape == (orc.frame_count(S, 1024, 256), 513)
        _points_agree(a, b)
        g, i = co.process(x, prm, threads=4)
        assert g.shape == i.shape == a[0].shape
    # caller-made points that jump further than a window can: the scatter still matches NumPy
    rng = np.random.default_rng(0)
    F, B = 40, 513
    e = np.where(rng.random((F, B)) < 0.05, rng.random((F, B)), 0.0)
    dc = rng.integers(-30, 30, (F, B)).astype(np.float64)
    dc = np.clip(dc, -np.arange(F)[:, None], (F - 1 - np.arange(F))[:, None])
    dk = np.clip(rng.normal(0, 3, (F, B)), -np.arange(B)[None, :], (B - 1 - np.arange(B))[None, :])
    assert np.array_equal(co.scatter_grid(dc, dk, e, prm, threads=4), orc.scatter_grid(dc, dk, e, prm))


def test_c_oracle_known_answers():
    """The analytic KATs of tests/test_oracle_kats.py (fixture tests/golden/kats.npz) on the C restatement:
    an off-bin tone is reassigned to its true frequency, a unit impulse to its true time, a linear chirp
    onto its instantaneous-frequency line.  (The C entry point takes float32 samples, as the GPU does.)"""
    z = np.load(os.path.join(GOLD, "kats.npz"))
    t = np.arange(SR) / SR
    prm = orc.Params(n_fft=2048, hop=512, noise_gate_db=-200)
    x = np.sin(2 * np.pi * float(z["tone_hz"]) * t).astype(np.float32)
    dt, dk, e = co.reassign_points(x, prm)
    k = int(np.argmax(e[10]))
    f_hat = (k + dk[10, k]) * SR / 2048
    assert abs(f_hat - float(z["tone_hz"])) < 2e-3
    assert abs(f_hat - float(z["tone_wrong_sign_hz"])) > 10
    assert abs(e[10, k] - 1.0) < 0.2

    x = np.zeros(SR, np.float32)
    m, pos = 5, int(z["impulse_pos"])
    x[pos] = 1.0
    dt, dk, e = co.reassign_points(x, prm)
    that = (m + dt[m, 1:-1]) * 512 + 1024
    assert np.abs(that - pos).max() < 1e-6 and np.abs(dk[m, 1:-1]).max() < 1e-6

    f0, f1 = float(z["chirp_f0"]), float(z["chirp_f1"])
    x = np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t)).astype(np.float32)
    dt, dk, e = co.reassign_points(x, orc.Params(n_fft=4096, hop=128, noise_gate_db=-200))
    for m in (100, 150, 250):
        k = int(np.argmax(e[m]))
        for kk in range(k - 3, k + 4):
            tt = ((m + dt[m, kk]) * 128 + 2048) / SR
            assert abs((kk + dk[m, kk]) * SR / 4096 - (f0 + (f1 - f0) * tt)) < 2e-3      # float32 samples


def test_c_oracle_matches_golden_fixture():
    z = np.load(os.path.join(GOLD, "reassign_n512_h128.npz"))
    prm = orc.Params(n_fft=int(z["n_fft"]), hop=int(z["hop"]), noise_gate_db=float(z["gate_db"]))
    dt, dk, e = co.reassign_points(z["x"], prm)
    assert np.allclose(dt, z["dt_cols"], atol=1e-5) and np.allclose(dk, z["dk_bins"], atol=1e-5)
    assert np.allclose(e, z["energy"], rtol=1e-6, atol=1e-12)
    grid, index = co.process(z["x"], prm)
    assert np.allclose(grid, z["grid"], rtol=1e-6, atol=1e-12)
    assert (index == z["index"]).all()


from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402
from test_properties import geom, kinds, make_signal  # noqa: E402


@settings(max_examples=40, derandomize=True, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(geom, kinds, st.integers(0, 2 ** 16), st.integers(0, 700), st.integers(1, 9),
       st.sampled_from([(0, 1.0), (64, 0.0), (200, 1.0), (97, 1.7)]), st.sampled_from([0.0, 0.5]), st.sampled_from([0.0, 0.8]))
def test_c_and_numpy_oracle_agree_on_random_cases(g, kind, seed, extra, threads, rows_scale, smoothing, agc):
    """Random geometries and degenerate signals (silence, DC, square wave, impulses: exact zeros, ties and
    knife-edge decisions are the rule there, not the exception)."""
    n_fft, div = g
    hop = max(1, n_fft // div)
    x = make_signal(kind, n_fft + 7 * hop + extra, seed)
    prm = orc.Params(n_fft=n_fft, hop=hop, display_rows=rows_scale[0], freq_scale=rows_scale[1],
                     smoothing=smoothing, agc_strength=agc)
    a = orc.reassign_points(x, prm, return_raw=True)
    b = co.reassign_points(x, prm, threads=threads, return_raw=True)
    # decisions may differ only where the deciding quantity sits on its threshold to round-off: such points
    # exist by construction here (a square wave's zero bins, an impulse's |dt| = N/2 edge)
    diff = (a[2] > 0) != (b[2] > 0)
    if diff.any():
        scale = max(a[3].max(), 1e-300)
        on_gate = np.abs(a[3] - prm.gate_lin) <= 1e-9 * max(prm.gate_lin, scale)
        Xh, Xth, Xdh = orc.stft3(x, n_fft, hop, 0, a[0].shape[0])
        _, dts, dkb = orc.reassign_operators(Xh, Xth, Xdh, n_fft)
        weak = a[3] <= 1e-18 * scale                                        # operators of a numerically zero bin are noise
        dc = dts / hop
        on_edge = (np.abs(np.abs(dts) - n_fft / 2) < 1e-6) | (np.abs(np.abs(dkb - np.rint(dkb)) - 0.5) < 1e-6) \
            | (np.abs(np.abs(dc - np.rint(dc)) - 0.5) < 1e-6)
        assert not (diff & ~(on_gate | weak | on_edge)).any()
    k = (a[2] > 0) & (b[2] > 0) & (a[3] > 1e-12 * max(a[3].max(), 1e-300))
    if k.any():
        assert np.abs(a[0] - b[0])[k].max() < 1e-6 and np.abs(a[1] - b[1])[k].max() < 1e-6
        assert (np.abs(a[2] - b[2])[k] / a[2][k]).max() < 1e-7
    ga, gb = orc.scatter_grid(*a[:3], prm), co.scatter_grid(*a[:3], prm, threads=threads)
    assert np.array_equal(ga, gb)
    d = np.abs(orc.postpass(ga, prm).astype(int) - co.postpass(ga, prm, threads=threads).astype(int))
    assert d.max() <= 1 and (d > 0).sum() <= max(1, 1e-4 * d.size)
