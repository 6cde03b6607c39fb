"""CPU tests: the stand-in oracle against THIRD-PARTY code it shares nothing with.

EM-Spec publishes no vectors (/root/reference/README.md:73) and librosa is not installed here, so the
strongest pin available is (1) SciPy's own short-time Fourier transform class for the three STFTs of
rows a1/a2 and (2) the reassignment formulas in the form librosa documents for
`librosa.reassigned_spectrogram` (frequency = bin frequency - Im(S_dh/S_h) sr/2pi, time = frame
centre + Re(S_th/S_h)/sr, the window derivative taken as a cyclic central difference of the window
samples), restated here on top of SciPy's STFT.  The oracle uses rfft on gathered frames and the
analytic window derivative; agreement of the two routes fixes signs, scales and frame alignment.
Parity with EM-Spec itself stays unpinned."""
import numpy as np
import pytest
import scipy.signal as ss

import reassign_oracle as orc

SR = 48000.0


def _scipy_stft(x, w, n_fft, hop, n_frames):
    """Frames [m*hop, m*hop + n_fft) from scipy.signal.ShortTimeFFT -> [F][B].  Its slice p is centred
    on sample p*hop, the oracle's frame m on m*hop + n_fft/2."""
    assert (n_fft // 2) % hop == 0
    S = ss.ShortTimeFFT(w, hop=hop, fs=1.0, fft_mode="onesided", phase_shift=None)
    Z = S.stft(np.asarray(x, np.float64))
    p0 = (n_fft // 2) // hop - S.p_min
    return Z[:, p0:p0 + n_frames].T


@pytest.mark.parametrize("n_fft,hop", [(256, 64), (1024, 256), (2048, 512), (4096, 128)])
def test_three_stfts_match_scipy_ShortTimeFFT(n_fft, hop):
    x = orc.synth_signal(3 * n_fft + 40 * hop + 17, SR, seed=n_fft).astype(np.float64)
    F = orc.frame_count(len(x), n_fft, hop)
    got = orc.stft3(x, n_fft, hop, 0, F)
    for w, X in zip(orc.windows(n_fft), got):
        Z = _scipy_stft(x, w, n_fft, hop, F)
        assert Z.shape == X.shape
        assert np.abs(Z - X).max() <= 1e-12 * np.abs(X).max()


def _librosa_form(x, n_fft, hop, n_frames):
    """(freq_hz, time_s, |S_h|^2) the way librosa's documentation states the method; window
    derivative = cyclic central difference (librosa.util.cyclic_gradient), time weights counted from
    the centre of the periodic window (n - n_fft/2, the oracle's own centre convention)."""
    h, _, _ = orc.windows(n_fft)
    dh = (np.roll(h, -1) - np.roll(h, 1)) / 2.0
    th = (np.arange(n_fft) - n_fft / 2) * h
    Sh = _scipy_stft(x, h, n_fft, hop, n_frames)
    Sdh = _scipy_stft(x, dh, n_fft, hop, n_frames)
    Sth = _scipy_stft(x, th, n_fft, hop, n_frames)
    with np.errstate(divide="ignore", invalid="ignore"):
        freq = np.arange(n_fft // 2 + 1)[None, :] * SR / n_fft - np.imag(Sdh / Sh) * (0.5 * SR / np.pi)
        time = (np.arange(n_frames)[:, None] * hop + n_fft / 2) / SR + np.real(Sth / Sh) / SR
    return freq, time, np.abs(Sh) ** 2


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (2048, 512)])
def test_reassigned_coordinates_match_the_librosa_form(n_fft, hop):
    x = orc.synth_signal(48000, SR, seed=5).astype(np.float64)
    prm = orc.Params(n_fft=n_fft, hop=hop, noise_gate_db=-200.0)
    F = orc.frame_count(len(x), n_fft, hop)
    dcol, dbin, en, raw = orc.reassign_points(x, prm, return_raw=True)
    freq, time, p = _librosa_form(x, n_fft, hop, F)
    k = np.arange(prm.n_bins)[None, :]
    m = np.arange(F)[:, None]
    f_orc = (k + dbin) * SR / n_fft
    t_orc = ((m + dcol) * hop + n_fft / 2) / SR
    assert np.allclose(raw, p * (4.0 / n_fft) ** 2, rtol=1e-10, atol=0)
    strong = (en > 1e-6 * raw.max()) & (k > 0) & (k < n_fft // 2)
    assert strong.sum() > 1000
    # times: same formula, so round-off only
    assert np.abs(t_orc - time)[strong].max() * SR < 1e-6
    # frequencies: the sampled derivative differs from the analytic one by the factor
    # sinc-like sin(2 pi/N)/(2 pi/N) on the displacement (periodic Hann is a single cosine), i.e. by at
    # most 0.5 bin * (2 pi/N)^2/6
    g = np.sin(2 * np.pi / n_fft) / (2 * np.pi / n_fft)
    d_orc = (f_orc - k * SR / n_fft)[strong]
    d_lib = (freq - k * SR / n_fft)[strong]
    assert np.abs(d_lib - g * d_orc).max() < 1e-7 * SR / n_fft
    near = np.abs(d_orc) <= 0.5 * SR / n_fft       # points that stay in their own bin
    assert near.sum() > 500
    assert np.abs(d_lib - d_orc)[near].max() < (0.5 * (1 - g) + 1e-7) * SR / n_fft
