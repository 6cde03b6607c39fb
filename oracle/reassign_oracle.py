"""STAND-IN ORACLE — float64 NumPy reassigned spectrogram.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: effree/EM-Spec ships no source, no tests, no golden vectors
(/root/reference/README.md:73 "The source code is maintained in a private
repository").  Nothing here is EM-Spec's own output.  This file
restates the *published* time-frequency reassignment method the README names
(/root/reference/README.md:3,11 "reassignment method") from the literature:

  * F. Auger, P. Flandrin, "Improving the readability of time-frequency and
    time-scale representations by the reassignment method", IEEE TSP 43(5), 1995
    (the operators  t^ = t - Re(X_th X_h*)/|X_h|^2,  w^ = w + Im(X_dh X_h*)/|X_h|^2);
  * S. Fulop, K. Fitz, JASA 119(1), 2006 (discrete form).

It is pinned instead by analytic known-answer tests (tests/test_oracle_kats.py:
off-bin tone, unit impulse, linear chirp, energy conservation) — SURVEY.md §4.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (em-spec_b200/) never does.

Conventions (SURVEY.md §7 "Sign conventions", §8c):
  frame m covers samples [m*H, m*H+N), centre m*H + N/2, no padding;
  periodic Hann h[n] = 0.5 - 0.5 cos(2 pi n / N); th[n] = (n - N/2) h[n];
  dh[n] = (pi/N) sin(2 pi n / N) (analytic derivative);
  X_w[k] = sum_n x[mH+n] w[n] exp(-2 pi i k n / N), k = 0..N/2;
  energy e = |X_h|^2 (4/N)^2   (a full-scale sine gives 0 dB);
  dt [samples] = +Re(X_th conj X_h)/|X_h|^2   -> t^ [columns] = m + dt/H;
  dk [bins]    = -Im(X_dh conj X_h)/|X_h|^2 * N/(2 pi) -> w^ [bins] = k + dk;
  points are reported as displacements (dt/H columns, dk bins) from (m, k) so that
  fp32 keeps full precision on hour-long streams;
  a point is dropped (energy 0, zero displacement) when e <= gate, |dt| > N/2,
  its nearest bin k + rint(dk) outside [0, N/2] (w^ within half a bin of the spectrum: DC and
  Nyquist energy, whose w^ sits exactly on 0 and N/2, stay off a knife edge; the test is on the
  rounded row itself so that it is exact in fp32 and a kept point can never leave the grid),
  or m + rint(dt/H) outside [0, F-1];
  nearest-cell deposit at (m + rint(dt/H), k + rint(dk)), round-half-even (np.rint / rintf).
Display shaping (/root/reference/README.md:46-51): gain, low-end boost, smoothing,
noise gate, dB range -> u8 colour index.  Semantics are stand-ins.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import scipy.fft

AGC_RELEASE_SECONDS = 1.0   # stand-in release time of the automatic gain control
LOW_END_CORNER_HZ = 200.0  # stand-in shape constant of the low-end boost weight
TOP_DB = 0.0               # display ceiling; floor = TOP_DB - db_range

FLAG_REASSIGN = 1          # 0: plain |X_h|^2 columns ("Natural"), 1: reassigned ("Enhanced")
FLAG_DETERMINISTIC = 2     # no effect in the oracle (float64 sums in index order)
FLAG_SYNC = 4              # no effect in the oracle


@dataclasses.dataclass
class Params:
    """Mirror of `ems_params` (include/emspec.h).  Defaults = settings.png "Default" preset."""
    n_fft: int = 4096            # README.md:43 "FFT Size"
    hop: int = 128               # README.md:44 "Scroll Speed" maps to hop
    sample_rate: float = 48000.0
    channels: int = 1
    db_range: float = 58.0       # README.md:46
    gain: float = 3.5            # README.md:47 (linear amplitude)
    low_end_boost: float = 3.9   # README.md:49
    smoothing: float = 0.0       # README.md:50
    noise_gate_db: float = -65.0 # README.md:51
    flags: int = FLAG_REASSIGN | FLAG_DETERMINISTIC
    display_rows: int = 0        # 0: one row per bin; > 0: rows of the warped frequency axis
    freq_scale: float = 1.0      # README.md:48 "Frequency Scale" (used when display_rows > 0)
    agc_strength: float = 0.0    # README.md:14 "AGC" / settings.png "AGC Strength"; 0 = off
    brightness: float = 0.44     # settings.png "Brightness 44 %": where the AGC draws the running level

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def n_rows(self) -> int:
        return self.display_rows if self.display_rows > 0 else self.n_bins

    @property
    def warp_a(self) -> float:
        return 10.0 ** (2.0 * self.freq_scale) - 1.0

    @property
    def gate_lin(self) -> float:
        return 10.0 ** (self.noise_gate_db / 10.0)


def frame_count(n_samples: int, n_fft: int, hop: int) -> int:
    """F = 1 + floor((S - N)/H), 0 when the stream is shorter than one frame."""
    return 0 if n_samples < n_fft else 1 + (n_samples - n_fft) // hop


def windows(n_fft: int):
    """(h, th, dh) in float64, frame-local index n = 0..N-1."""
    n = np.arange(n_fft, dtype=np.float64)
    ang = 2.0 * np.pi * n / n_fft
    h = 0.5 - 0.5 * np.cos(ang)
    th = (n - n_fft / 2) * h
    dh = (np.pi / n_fft) * np.sin(ang)
    return h, th, dh


def stft3(x: np.ndarray, n_fft: int, hop: int, m0: int, m1: int, workers: int = 1):
    """X_h, X_th, X_dh for frames [m0, m1) -> three complex128 arrays [m1-m0, N/2+1]."""
    h, th, dh = windows(n_fft)
    idx = (np.arange(m0, m1)[:, None] * hop) + np.arange(n_fft)[None, :]
    fr = np.asarray(x, dtype=np.float64)[idx]
    Xh = scipy.fft.rfft(fr * h, axis=1, workers=workers)
    Xth = scipy.fft.rfft(fr * th, axis=1, workers=workers)
    Xdh = scipy.fft.rfft(fr * dh, axis=1, workers=workers)
    return Xh, Xth, Xdh


def reassign_operators(Xh, Xth, Xdh, n_fft: int):
    """(energy, dt_samples, dk_bins) from the three STFTs; zero where |X_h| = 0."""
    p = Xh.real * Xh.real + Xh.imag * Xh.imag
    safe = np.where(p > 0.0, p, 1.0)
    dt = (Xth.real * Xh.real + Xth.imag * Xh.imag) / safe
    dk = -(Xdh.imag * Xh.real - Xdh.real * Xh.imag) / safe * (n_fft / (2.0 * np.pi))
    e = p * (4.0 / n_fft) ** 2
    dt = np.where(p > 0.0, dt, 0.0)
    dk = np.where(p > 0.0, dk, 0.0)
    return e, dt, dk


def keep_mask(e, dt, dk, m, k, F: int, prm: Params):
    """The drop rule (SURVEY.md §7 "Out-of-support points"; header of this file): a point of frame m,
    bin k with energy e, dt [samples], dk [bins] is kept iff it clears the gate, stays inside the
    window support and its nearest cell (m + rint(dt/H), k + rint(dk)) lies on the grid."""
    N, H = prm.n_fft, prm.hop
    col = m + np.rint(dt / H)
    row = k + np.rint(dk)
    return (e > prm.gate_lin) & (np.abs(dt) <= N / 2) & (row >= 0) & (row <= N / 2) \
        & (col >= 0) & (col <= F - 1)


def reassign_points(x: np.ndarray, prm: Params, chunk: int = 256, workers: int = 1,
                    return_raw: bool = False):
    """The a1..a3 path: fp64 points as [F][B] arrays (dt_cols, dk_bins, energy).

    Same meaning as ems_process_points (include/emspec.h): the point of frame f, bin k
    sits at column f + dt_cols, bin k + dk_bins.  Dropped points carry energy 0 and
    zero displacement.  With return_raw=True also returns the un-gated energy (used
    by the tests to band the coordinate tolerance by level).
    """
    N, H = prm.n_fft, prm.hop
    x = np.asarray(x)
    F = frame_count(x.shape[-1], N, H)
    B = prm.n_bins
    dcol = np.zeros((F, B), np.float64)
    dbin = np.zeros((F, B), np.float64)
    en = np.zeros((F, B), np.float64)
    raw = np.empty((F, B), np.float64) if return_raw else None
    k = np.arange(B, dtype=np.float64)[None, :]
    gate = prm.gate_lin
    for m0 in range(0, F, chunk):
        m1 = min(F, m0 + chunk)
        Xh, Xth, Xdh = stft3(x, N, H, m0, m1, workers)
        e, dt, dk = reassign_operators(Xh, Xth, Xdh, N)
        m = np.arange(m0, m1, dtype=np.float64)[:, None]
        if raw is not None:
            raw[m0:m1] = e
        if prm.flags & FLAG_REASSIGN:
            dc = dt / H
            ok = keep_mask(e, dt, dk, m, k, F, prm)
            dcol[m0:m1] = np.where(ok, dc, 0.0)
            dbin[m0:m1] = np.where(ok, dk, 0.0)
        else:
            ok = e > gate
        en[m0:m1] = np.where(ok, e, 0.0)
    if return_raw:
        return dcol, dbin, en, raw
    return dcol, dbin, en


def output_row(k, dk, prm: Params):
    """Row of a point at reassigned frequency k + dk [bins].  Without display_rows: the
    nearest bin.  With it ("Frequency Scale", README.md:48; stand-in shape): rows of a
    log1p-warped axis, row = rint((R-1) log1p(a x)/log1p(a)), x = (k+dk)/(N/2),
    a = 10^(2 freq_scale) - 1 (a -> 0: linear)."""
    if prm.display_rows <= 0:
        return k + np.rint(dk).astype(np.int64)
    x = np.clip((k + dk) / (prm.n_fft / 2), 0.0, 1.0)
    a = prm.warp_a
    u = np.log1p(a * x) / np.log1p(a) if a > 1e-6 else x
    return np.rint(u * (prm.display_rows - 1)).astype(np.int64)


def row_frequencies(prm: Params) -> np.ndarray:
    """Centre frequency [Hz] of every output row."""
    if prm.display_rows <= 0:
        return np.arange(prm.n_bins, dtype=np.float64) * prm.sample_rate / prm.n_fft
    u = np.arange(prm.display_rows, dtype=np.float64) / (prm.display_rows - 1)
    a = prm.warp_a
    x = np.expm1(u * np.log1p(a)) / a if a > 1e-6 else u
    return x * prm.sample_rate / 2


def hz_to_row(freq_hz: float, prm: Params) -> float:
    """Fractional output row of a frequency: the mapping `output_row` rounds (include/emspec.h::ems_hz_to_row)."""
    R = prm.n_rows
    if prm.display_rows <= 0:
        r = freq_hz * prm.n_fft / prm.sample_rate
    else:
        x = min(max(freq_hz / (prm.sample_rate / 2), 0.0), 1.0)
        a = prm.warp_a
        r = (R - 1) * (math.log1p(a * x) / math.log1p(a) if a > 1e-6 else x)
    return min(max(r, 0.0), R - 1.0)


NOTE_NAMES = ("C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B")


def cursor_info(column: float, row: float, prm: Params) -> dict:
    """Time, frequency and nearest equal-tempered note under output cell (column, row)
    (README.md:39 "note and frequency information"; include/emspec.h::ems_cursor_info).  Fractional
    rows follow the axis between the row centres of `row_frequencies`; A4 = 440 Hz = MIDI 69."""
    R = prm.n_rows
    r = min(max(float(row), 0.0), R - 1.0)
    if prm.display_rows <= 0:
        f = r * prm.sample_rate / prm.n_fft
    else:
        u = r / (R - 1) if R > 1 else 0.0
        a = prm.warp_a
        f = (math.expm1(u * math.log1p(a)) / a if a > 1e-6 else u) * prm.sample_rate / 2
    out = {"time_s": (column * prm.hop + prm.n_fft / 2) / prm.sample_rate, "freq_hz": f,
           "midi_note": -1, "cents": 0.0, "name": ""}
    if f >= 1.0:
        pitch = 69.0 + 12.0 * math.log2(f / 440.0)
        note = math.floor(pitch + 0.5)
        out.update(midi_note=int(note), cents=100.0 * (pitch - note),
                   name=f"{NOTE_NAMES[int(note) % 12]}{int(note) // 12 - 1}")
    return out


# Stand-in colour maps (README.md:15,45 "Multiple Color Maps"): (position, r, g, b) control colours.
COLORMAPS = {
    "inferno": ((0, 0, 0, 4), (64, 87, 16, 110), (128, 188, 55, 84), (192, 249, 142, 9), (255, 252, 255, 164)),  # settings.png default
    "gray": ((0, 0, 0, 0), (255, 255, 255, 255)),
    "heat": ((0, 0, 0, 0), (85, 200, 0, 0), (170, 255, 200, 0), (255, 255, 255, 255)),
    "magma": ((0, 0, 0, 4), (64, 81, 18, 124), (128, 183, 55, 121), (192, 252, 137, 97), (255, 252, 253, 191)),
    "viridis": ((0, 68, 1, 84), (64, 59, 82, 139), (128, 33, 145, 140), (192, 94, 201, 98), (255, 253, 231, 37)),
    "ice": ((0, 0, 0, 0), (96, 0, 60, 160), (192, 80, 200, 255), (255, 255, 255, 255)),
}


def builtin_colormap(name: str) -> np.ndarray:
    """256 packed 0xAABBGGRR pixels: linear ramps between the control colours, rounded to nearest in
    integer arithmetic ((c0 (d - t) + c1 t + d // 2) // d), alpha 255."""
    lut = np.zeros(256, np.uint32)
    stops = COLORMAPS[name]
    for (p0, *c0), (p1, *c1) in zip(stops[:-1], stops[1:]):
        d = p1 - p0
        for i in range(p0, p1 + 1):
            t = i - p0
            r, g, b = ((a * (d - t) + z * t + d // 2) // d for a, z in zip(c0, c1))
            lut[i] = 0xFF000000 | (b << 16) | (g << 8) | r
    return lut


def scatter_grid(dcol, dbin, energy, prm: Params | None = None) -> np.ndarray:
    """a4: G[f + rint(dt_cols), output_row(k + dk_bins)] += e over kept (e > 0) points; fp64 [F][R]."""
    F, B = energy.shape
    if prm is None:
        prm = Params(n_fft=2 * (B - 1))
    R = prm.n_rows
    f, k = np.nonzero(energy > 0.0)
    col = f + np.rint(dcol[f, k]).astype(np.int64)
    row = output_row(k, dbin[f, k], prm)
    flat = np.bincount(col * R + row, weights=energy[f, k], minlength=F * R)
    return flat.reshape(F, R)


def low_end_weight(prm: Params) -> np.ndarray:
    """w_low = 1 + (boost-1)/(1 + (f/200 Hz)^2) per output row: `boost` at DC, -> 1 at high f."""
    f = row_frequencies(prm)
    return 1.0 + (prm.low_end_boost - 1.0) / (1.0 + (f / LOW_END_CORNER_HZ) ** 2)


def shaped_energy(grid: np.ndarray, prm: Params) -> np.ndarray:
    """E = G gain^2 w_low(k), then the temporal EMA y[m] = s y[m-1] + (1-s) E[m], y[-1] = 0."""
    E = grid * (prm.gain ** 2) * low_end_weight(prm)[None, :]
    s = prm.smoothing
    if s > 0.0:
        y = np.empty_like(E)
        acc = np.zeros(E.shape[1])
        for m in range(E.shape[0]):
            acc = s * acc + (1.0 - s) * E[m]
            y[m] = acc
        E = y
    return E


def agc_scale(E: np.ndarray, prm: Params) -> np.ndarray:
    """Automatic gain (README.md:14; SURVEY.md §8f-2, stand-in semantics): per column m
    level[m] = max(peak[m], lambda level[m-1]), peak[m] = max_r E[m, r], level[-1] = 0,
    lambda = exp(-hop / (sample_rate * 1 s)); cells are drawn at E * T / level^strength, where the
    "Brightness" control sets T = 10^(-(1 - brightness) db_range / 10): at full strength the loudest
    cell of the running level gets colour index 255 * brightness."""
    lam = math.exp(-prm.hop / (prm.sample_rate * AGC_RELEASE_SECONDS))
    peak = E.max(axis=1) if E.shape[1] else np.zeros(E.shape[0])
    scale = np.ones(E.shape[0])
    target = 10.0 ** (-(1.0 - prm.brightness) * prm.db_range / 10.0)
    lv = 0.0
    for m in range(E.shape[0]):
        lv = max(peak[m], lam * lv)
        if lv > 0.0:
            scale[m] = target * lv ** (-prm.agc_strength)
    return scale


def postpass(grid: np.ndarray, prm: Params) -> np.ndarray:
    """a5: shaped energy -> (AGC) -> dB -> gate -> u8 colour index [F][R]."""
    E = shaped_energy(grid, prm)
    Ed = E * agc_scale(E, prm)[:, None] if prm.agc_strength > 0.0 else E
    with np.errstate(divide="ignore"):
        db_gate = 10.0 * np.log10(E)
        db = 10.0 * np.log10(Ed)
    floor = TOP_DB - prm.db_range
    v = np.rint(255.0 * (db - floor) / prm.db_range)
    v = np.clip(v, 0.0, 255.0)
    v = np.where((E > 0.0) & (db_gate >= prm.noise_gate_db), v, 0.0)
    return v.astype(np.uint8)


def process(x: np.ndarray, prm: Params, workers: int = 1):
    """Whole path for one channel: (grid fp64 [F][B], index u8 [F][B])."""
    dcol, dbin, en = reassign_points(x, prm, workers=workers)
    grid = scatter_grid(dcol, dbin, en, prm)
    return grid, postpass(grid, prm)


# ---------------------------------------------------------------------------
# Synthetic input: chirp + two tones + noise, float32 samples.
# ---------------------------------------------------------------------------
def synth_signal(n_samples: int, sample_rate: float = 48000.0, seed: int = 0,
                 clip_index: int | None = None) -> np.ndarray:
    """0.5*logchirp(20 Hz -> 20 kHz over the clip) + 0.25*sin(2 pi f1 t)
    + 0.125*sin(2 pi f2 t) + 1e-3*N(0,1), cast to float32.
    Single stream: f1 = 440, f2 = 3000.5.  Batch clips: f1 = 440 + 7*(i mod 64),
    f2 = 3000.5 + 11*(i mod 128), seed = clip index."""
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    T = n_samples / sample_rate
    f0, f1c = 20.0, 20000.0
    r = math.log(f1c / f0)
    phase = 2.0 * np.pi * f0 * T / r * np.expm1(r * t / T)
    fa, fb = 440.0, 3000.5
    if clip_index is not None:
        fa += 7.0 * (clip_index % 64)
        fb += 11.0 * (clip_index % 128)
        seed = clip_index
    rng = np.random.default_rng(seed)
    x = 0.5 * np.sin(phase) + 0.25 * np.sin(2 * np.pi * fa * t) \
        + 0.125 * np.sin(2 * np.pi * fb * t) + 1e-3 * rng.standard_normal(n_samples)
    return x.astype(np.float32)


def synth_music(n_samples: int, sample_rate: float = 48000.0, seed: int = 0) -> np.ndarray:
    """Dense, music-like test signal (stand-in for "system audio", README.md:36):
    pink noise at about -26 dBFS rms, three notes with vibrato and 12 harmonics each (1/h roll-off,
    repeating attack/decay envelopes, peaks near -20 dBFS) and a decaying noise burst every 0.25 s.
    Unlike synth_signal almost every bin of every frame is above the default -65 dB gate."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    spec = np.fft.rfft(rng.standard_normal(n_samples))
    f = np.fft.rfftfreq(n_samples, 1.0 / sample_rate)
    spec[1:] /= np.sqrt(f[1:])
    spec[0] = 0.0
    pink = np.fft.irfft(spec, n_samples)
    x = 0.05 * pink / np.sqrt(np.mean(pink * pink))
    for i, f0 in enumerate((110.0, 196.0, 329.63)):
        vib = 0.005 * f0 / 5.0 * np.sin(2 * np.pi * 5.0 * t + i)            # +-0.5 % at 5 Hz, as phase deviation / (2 pi)
        env = np.exp(-3.0 * np.mod(t + 0.17 * i, 0.5)) * (1.0 - np.exp(-200.0 * np.mod(t + 0.17 * i, 0.5)))
        for hn in range(1, 13):
            if hn * f0 < 0.45 * sample_rate:
                x += 0.1 / hn * env * np.sin(2 * np.pi * hn * (f0 * t + vib) + 0.3 * hn)
    burst = np.exp(-60.0 * np.mod(t, 0.25)) * rng.standard_normal(n_samples)
    x += 0.03 * burst
    return x.astype(np.float32)
