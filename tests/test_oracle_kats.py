"""CPU tests pinning the float64 stand-in oracle by analytic known answers.
STAND-IN: EM-Spec publishes no vectors (/root/reference/README.md:73); the KATs are
mathematics of the reassignment method the README names (README.md:3,11)."""
import os

import numpy as np
import pytest

import reassign_oracle as orc

SR = 48000
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_frame_count_edges():
    assert orc.frame_count(0, 1024, 256) == 0
    assert orc.frame_count(1023, 1024, 256) == 0
    assert orc.frame_count(1024, 1024, 256) == 1
    assert orc.frame_count(1279, 1024, 256) == 1
    assert orc.frame_count(1280, 1024, 256) == 2
    assert orc.frame_count(480000, 2048, 512) == 934          # configs[0]
    assert orc.frame_count(172800000, 4096, 128) == 1349969   # configs[2]
    assert orc.frame_count(2880000, 4096, 256) == 11235       # one clip of configs[3]


def test_windows_definition():
    h, th, dh = orc.windows(1024)
    n = np.arange(1024)
    assert h[0] == 0 and abs(h[512] - 1) < 1e-15
    assert np.allclose(th, (n - 512) * h)
    # analytic derivative vs a centred finite difference of the continuous Hann
    eps = 1e-5
    fd = (-0.5 * np.cos(2 * np.pi * (n + eps) / 1024) + 0.5 * np.cos(2 * np.pi * (n - eps) / 1024)) / (2 * eps)
    assert np.allclose(dh, fd, atol=1e-9)


def test_kat_offbin_tone_frequency_sign():
    z = np.load(os.path.join(GOLD, "kats.npz"))
    t = np.arange(SR) / SR
    x = np.sin(2 * np.pi * float(z["tone_hz"]) * t)
    prm = orc.Params(n_fft=2048, hop=512, noise_gate_db=-200)
    dt, dk, e = orc.reassign_points(x, prm)
    k = int(np.argmax(e[10]))
    f_hat = (k + dk[10, k]) * SR / 2048
    assert abs(f_hat - float(z["tone_hz"])) < 2e-3
    assert abs(f_hat - float(z["tone_wrong_sign_hz"])) > 10          # the wrong sign lands here
    assert abs(e[10, k] - 1.0) < 0.2                                   # ~0 dB for a full-scale sine


def test_kat_impulse_time_sign():
    z = np.load(os.path.join(GOLD, "kats.npz"))
    prm = orc.Params(n_fft=2048, hop=512, noise_gate_db=-200)
    x = np.zeros(SR)
    m = 5
    pos = m * 512 + 1024 + 300
    assert pos == int(z["impulse_pos"])
    x[pos] = 1.0
    dt, dk, e = orc.reassign_points(x, prm)
    that = (m + dt[m, 1:-1]) * 512 + 1024
    assert np.abs(that - pos).max() < 1e-6
    assert np.abs(dk[m, 1:-1]).max() < 1e-6
    assert abs(int(z["impulse_wrong_sign"]) - (pos - 600)) == 0


def test_kat_linear_chirp_on_if_line():
    z = np.load(os.path.join(GOLD, "kats.npz"))
    f0, f1 = float(z["chirp_f0"]), float(z["chirp_f1"])
    t = np.arange(SR) / SR
    x = np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t))
    prm = orc.Params(n_fft=4096, hop=128, noise_gate_db=-200)
    dt, dk, e = orc.reassign_points(x, prm)
    for m in (100, 150, 250):
        k = int(np.argmax(e[m]))
        for kk in range(k - 3, k + 4):
            tt = ((m + dt[m, kk]) * 128 + 2048) / SR
            assert abs((kk + dk[m, kk]) * SR / 4096 - (f0 + (f1 - f0) * tt)) < 1e-3


def test_operators_equal_phase_derivatives_of_the_stft():
    """Independent route to the same quantities: the reassigned coordinates are, by definition
    (Kodera et al. 1976; Auger & Flandrin 1995 eq. 1-2), the local group delay -dphi/domega and
    the instantaneous frequency dphi/dt of the STFT.  Here both come from finite differences of
    the STFT phase — frames one sample apart, and a 32x zero-padded spectrum — on the bench
    signal (log chirp + two tones + noise), and must agree with the three-window operators of
    the oracle at every strong bin.  Pins signs, scales and the frame-local phase reference
    without any analytic model of the signal."""
    N = 2048
    x = orc.synth_signal(6 * N, SR, seed=3).astype(np.float64)
    h, _, _ = orc.windows(N)
    starts = np.array([1000, 2500, 4000, 5500, 7000])
    Xh, Xth, Xdh = orc.stft3(x, N, 1, 0, starts.max() + 2)          # hop 1: frame m starts at sample m
    e, dts, dkb = orc.reassign_operators(Xh, Xth, Xdh, N)
    k = np.arange(N // 2 + 1)
    checked = 0
    errs_t = []
    for m in starts:
        strong = (e[m] > e[m].max() * 1e-2) & (k > 2) & (k < N // 2 - 2)      # main lobes: within 20 dB of the peak
        # instantaneous frequency: central difference of the phase over frames m-1, m+1
        dphi_dt = np.angle(Xh[m + 1] * np.conj(Xh[m - 1])) / 2.0                   # rad / sample, |.| < pi/2
        # the principal value sits within pi/2 of the bin's own frequency 2 pi k / N (mod pi)
        w_bin = 2 * np.pi * k / N
        dphi_dt = w_bin + (dphi_dt - w_bin + np.pi / 2) % np.pi - np.pi / 2
        if_bins = dphi_dt * N / (2 * np.pi)
        err_w = np.abs(if_bins - (k + dkb[m]))[strong]
        # group delay: central difference of the phase over +-1 bin of a 32x finer frequency grid
        P = 32
        Z = np.fft.rfft(x[m:m + N] * h, P * N)
        dphi_dw = np.angle(Z[P * k[1:-1] + 1] * np.conj(Z[P * k[1:-1] - 1])) / (2 * 2 * np.pi / (P * N))   # samples
        n_hat = -dphi_dw                                                                              # from the frame start
        err_t = np.abs((n_hat - N / 2) - dts[m][1:-1])[strong[1:-1]]
        assert err_w.max() < 2e-3, err_w.max()          # bins   (finite-difference error of a chirp: O(rate^2))
        assert np.median(err_w) < 1e-4
        errs_t.append(err_t)
        checked += int(strong.sum())
    # samples, of a 2048-sample window.  The difference quotient in frequency breaks down where two
    # components interfere (the phase turns by pi across a spectral zero), hence quantiles, not the max
    errs_t = np.concatenate(errs_t)
    assert np.median(errs_t) < 5e-3 and np.quantile(errs_t, 0.9) < 0.1, (np.median(errs_t), np.quantile(errs_t, 0.9))
    assert checked > 100


def test_energy_conservation_and_drop_rule():
    x = orc.synth_signal(SR, SR, seed=0)
    prm = orc.Params(n_fft=2048, hop=512)
    dt, dk, e = orc.reassign_points(x, prm)
    grid = orc.scatter_grid(dt, dk, e)
    assert abs(grid.sum() - e.sum()) < 1e-9 * e.sum()
    assert (np.abs(dt) <= 2048 / 2 / 512 + 1e-12).all()
    k = np.arange(1025)[None, :]
    assert ((k + dk) >= -0.5).all() and ((k + dk) <= 1024.5).all()
    assert (e[e > 0] > prm.gate_lin).all()
    # dropped points are zeroed, not NaN
    assert np.isfinite(dt).all() and np.isfinite(dk).all()


def test_silence_dc_square():
    prm = orc.Params(n_fft=1024, hop=256)
    g, idx = orc.process(np.zeros(8192), prm)
    assert g.max() == 0 and idx.max() == 0
    for x in (np.ones(8192), np.sign(np.sin(2 * np.pi * 997.0 * np.arange(8192) / SR))):
        g, idx = orc.process(x, prm)
        assert np.isfinite(g).all()


def test_postpass_controls():
    prm = orc.Params(n_fft=512, hop=128, gain=1.0, low_end_boost=1.0, db_range=60.0, noise_gate_db=-50.0)
    grid = np.zeros((4, 257))
    grid[:, 10] = 1.0          # 0 dB -> 255
    grid[:, 20] = 1e-3         # -30 dB -> 127.5 -> 128 (half-even on .5 -> 128)
    grid[:, 30] = 1e-5 * 0.99  # below the -50 dB gate -> 0
    grid[:, 40] = 1e-7         # below the floor -> 0
    idx = orc.postpass(grid, prm)
    assert idx[0, 10] == 255 and idx[0, 20] in (127, 128) and idx[0, 30] == 0 and idx[0, 40] == 0
    w = orc.low_end_weight(orc.Params(n_fft=512, low_end_boost=3.9))
    assert abs(w[0] - 3.9) < 1e-12 and abs(w[-1] - 1.0) < 1e-3 and (np.diff(w) <= 0).all()
    # smoothing: EMA step response
    prm.smoothing = 0.5
    E = orc.shaped_energy(grid, prm)
    assert np.allclose(E[:, 10], [0.5, 0.75, 0.875, 0.9375])


def test_oracle_matches_golden_fixture():
    z = np.load(os.path.join(GOLD, "reassign_n512_h128.npz"))
    prm = orc.Params(n_fft=int(z["n_fft"]), hop=int(z["hop"]), noise_gate_db=float(z["gate_db"]))
    dt, dk, e = orc.reassign_points(z["x"], prm)
    assert np.allclose(dt, z["dt_cols"], atol=1e-5) and np.allclose(dk, z["dk_bins"], atol=1e-5)
    assert np.allclose(e, z["energy"], rtol=1e-6, atol=1e-12)
    grid, index = orc.process(z["x"], prm)
    assert np.allclose(grid, z["grid"], rtol=1e-6, atol=1e-12)
    assert (index == z["index"]).all()


def test_synth_signal_is_deterministic():
    a = orc.synth_signal(4800, SR, seed=0)
    b = orc.synth_signal(4800, SR, seed=0)
    assert a.dtype == np.float32 and (a == b).all()
    assert np.abs(a).max() < 1.0
    c = orc.synth_signal(4800, SR, clip_index=3)
    assert not (a == c).all()


def test_frequency_scale_rows():
    """Warped display axis: monotone, end points pinned, energy conserved,
    freq_scale 0 is a linear resample, larger scale gives the bass more rows."""
    prm = orc.Params(n_fft=4096, hop=128, display_rows=546, freq_scale=1.0)
    k = np.arange(2049)
    rows = orc.output_row(k, np.zeros(2049), prm)
    assert rows[0] == 0 and rows[-1] == 545 and (np.diff(rows) >= 0).all()
    lin = orc.output_row(k, np.zeros(2049), orc.Params(n_fft=4096, display_rows=546, freq_scale=0.0))
    assert np.abs(lin - np.rint(k / 2048 * 545)).max() == 0
    assert rows[k == 100][0] > lin[k == 100][0]                   # 1.17 kHz sits higher on the warped axis
    f = orc.row_frequencies(prm)
    assert f[0] == 0 and abs(f[-1] - 24000) < 1e-6 and (np.diff(f) > 0).all()
    back = orc.output_row(f / (48000 / 4096), np.zeros(546), prm)  # row centres map back to their rows
    assert (back == np.arange(546)).all()
    x = orc.synth_signal(24000, SR, seed=3)
    dt, dk, e = orc.reassign_points(x, prm)
    g = orc.scatter_grid(dt, dk, e, prm)
    assert g.shape == (e.shape[0], 546) and abs(g.sum() - e.sum()) < 1e-9 * e.sum()
    w = orc.low_end_weight(prm)
    assert w.shape == (546,) and abs(w[0] - prm.low_end_boost) < 1e-12


def test_agc_level_recurrence():
    """AGC: peak hold with exponential release; strength 1 pins the loudest
    recent cell at 0 dB; strength 0 is the identity."""
    prm = orc.Params(n_fft=512, hop=128, gain=1.0, low_end_boost=1.0, agc_strength=1.0, db_range=60.0,
                     noise_gate_db=-90.0, brightness=1.0)
    grid = np.zeros((6, 257))
    grid[0, 5] = 1e-3
    grid[1, 7] = 1e-2
    grid[2:, 9] = 1e-4
    sc = orc.agc_scale(grid, prm)
    lam = np.exp(-128 / 48000.0)
    assert np.allclose(sc[:3], [1e3, 1e2, 1 / (1e-2 * lam)])
    assert np.allclose(1 / sc[3:], [1e-2 * lam ** 2, 1e-2 * lam ** 3, 1e-2 * lam ** 4])
    idx = orc.postpass(grid, prm)
    assert idx[0, 5] == 255 and idx[1, 7] == 255 and idx[2, 9] < 255
    off = orc.postpass(grid, orc.Params(**{**prm.__dict__, "agc_strength": 0.0}))
    assert off[1, 7] == int(np.rint(255 * 40 / 60))
    # "Brightness" (settings.png): the running level is drawn at index 255 * brightness
    half = orc.Params(**{**prm.__dict__, "brightness": 0.44})
    assert np.allclose(orc.agc_scale(grid, half), sc * 10 ** (-0.56 * 6.0))
    idx = orc.postpass(grid, half)
    assert idx[0, 5] == int(np.rint(255 * 0.44)) and idx[1, 7] == int(np.rint(255 * 0.44))
    assert (orc.postpass(grid, orc.Params(**{**half.__dict__, "agc_strength": 0.0})) == off).all()   # only with the AGC on


def test_drop_rule_is_decided_on_the_rounded_row():
    """a point whose w^ = k + dk sits at N/2 + 0.5 (or -0.5) must not reach row N/2 + 1
    (or -1) through round-half-even; the rule tests the rounded row itself."""
    prm = orc.Params(n_fft=4096, hop=128, noise_gate_db=-200.0)
    F = 100
    e = np.ones(6)
    dt = np.zeros(6)
    k = np.array([2047.0, 2047.0, 2046.0, 1.0, 1.0, 0.0])
    dk = np.array([1.5, 1.49, 2.5, -1.5, -1.49, -0.5])
    ok = orc.keep_mask(e, dt, dk, np.full(6, 50.0), k, F, prm)
    rows = k + np.rint(dk)
    assert ((rows >= 0) & (rows <= 2048))[ok].all()
    assert list(ok) == [False, True, True, False, True, True]      # 1.5 -> 2 (row 2049) and -1.5 -> -2 (row -1) are dropped
    # kept points always land on the grid, so the scatter never spills into a neighbouring column
    en = np.zeros((F, prm.n_bins)); dc = np.zeros_like(en); db = np.zeros_like(en)
    en[50, 2047] = 1.0; db[50, 2047] = 1.49
    g = orc.scatter_grid(dc, db, en, prm)
    assert g[50, 2048] == 1.0 and g.sum() == 1.0


def test_cursor_readout_known_notes_and_round_trip():
    """README.md:39 read-out: known pitches, and row -> Hz -> row is the identity on both axes."""
    for kw in (dict(n_fft=4096, hop=128), dict(n_fft=2048, hop=512, display_rows=546, freq_scale=1.0),
               dict(n_fft=8192, hop=256, display_rows=300, freq_scale=0.0)):
        prm = orc.Params(**kw)
        for hz, midi, name in ((440.0, 69, "A4"), (261.6256, 60, "C4"), (27.5, 21, "A0"), (4186.009, 108, "C8"),
                               (466.1638, 70, "A#4"), (8.1758, 0, "C-1")):
            r = orc.hz_to_row(hz, prm)
            c = orc.cursor_info(3, r, prm)
            assert abs(c["freq_hz"] - hz) < 1e-9 * hz and c["midi_note"] == midi and c["name"] == name
            assert abs(c["cents"]) < 0.01
            assert abs(c["time_s"] - (3 * prm.hop + prm.n_fft / 2) / prm.sample_rate) < 1e-15
        rows = np.linspace(0, prm.n_rows - 1, 257)
        back = [orc.hz_to_row(orc.cursor_info(0, r, prm)["freq_hz"], prm) for r in rows]
        assert np.abs(np.array(back) - rows).max() < 1e-8
        fr = orc.row_frequencies(prm)
        assert np.allclose([orc.cursor_info(0, r, prm)["freq_hz"] for r in range(0, prm.n_rows, 37)], fr[::37], rtol=1e-12, atol=1e-12)
    assert orc.cursor_info(0, 0, orc.Params())["midi_note"] == -1          # DC: no note


def test_builtin_colour_maps_are_ramps_through_their_control_colours():
    for name, stops in orc.COLORMAPS.items():
        lut = orc.builtin_colormap(name)
        assert lut.shape == (256,) and ((lut >> 24) == 0xFF).all()
        for pos, r, g, b in stops:
            assert lut[pos] == (0xFF000000 | (b << 16) | (g << 8) | r), (name, pos)
        for ch in (0, 8, 16):                                              # piecewise monotone between stops
            v = ((lut >> ch) & 0xFF).astype(int)
            for (p0, *_), (p1, *_) in zip(stops[:-1], stops[1:]):
                d = np.diff(v[p0:p1 + 1])
                assert (d >= 0).all() or (d <= 0).all(), (name, ch, p0)
    assert list(orc.COLORMAPS)[0] == "inferno"                             # the Default preset's map (settings.png)
