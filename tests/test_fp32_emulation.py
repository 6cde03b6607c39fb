"""CPU check of the kernels' ALGEBRA (not of the kernels): tools/fp32_emulate.py restates one-FFT +
untangle + Hann / dh stencils + Auger-Flandrin operators in NumPy complex64 / float32.  It must meet
the parity tolerances against the float64 oracle on the sparse bench signal and on the dense
music-like signal — which is what justifies the deviation from SURVEY.md §7 (frequency-domain Hann
stencils in fp32) without a GPU, and sizes new parity cases before GPU time is spent on them."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import reassign_oracle as orc
from fp32_emulate import grid_index_fp32, points_fp32
from parity_util import check_grid_dense, check_index, check_points

SR = 48000


@pytest.mark.parametrize("n_fft,hop,sig", [(4096, 128, "sparse"), (4096, 128, "music"), (2048, 512, "sparse"), (8192, 256, "music")])
def test_fp32_algebra_meets_the_tolerances(n_fft, hop, sig):
    n = int(0.75 * SR) if n_fft <= 4096 else SR
    x = orc.synth_signal(n, SR, seed=5) if sig == "sparse" else orc.synth_music(n, SR, seed=6)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    st = check_points(points_fp32(x, prm), x, prm)
    assert st["max_dt_strong"] < 2e-4 and st["e_rel_l2"] < 1e-6          # a tenth of the tolerance and better
    g, idx = grid_index_fp32(x, prm)
    err, grid_o, amb = check_grid_dense(g, x, prm)
    check_index(idx, grid_o, prm, x)


def test_fp32_drop_rule_matches_oracle_rule_on_ties():
    """The emulation decides on the rounded row like the kernels and the oracle."""
    prm = orc.Params(n_fft=512, hop=128, noise_gate_db=-200.0)
    x = (0.3 * np.random.default_rng(3).standard_normal(20000)).astype(np.float32)
    dt, dk, e = points_fp32(x, prm)
    k = np.arange(prm.n_bins)[None, :]
    rows = k + np.rint(dk)
    assert ((rows >= 0) & (rows <= prm.n_fft // 2))[e > 0].all()
