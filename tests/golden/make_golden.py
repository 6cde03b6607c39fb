"""Regenerates tests/golden/*.npz from the float64 stand-in oracle (oracle/reassign_oracle.py).

STAND-IN fixtures: EM-Spec ships no golden vectors; these pin the oracle
against accidental change and travel to the GPU box, where /root/reference and nothing
else outside the repo exists.  Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import reassign_oracle as orc  # noqa: E402


def main():
    sr = 48000
    prm = orc.Params(n_fft=512, hop=128, noise_gate_db=-65.0)
    x = orc.synth_signal(8192, sr, seed=11)
    dt, dk, e, raw = orc.reassign_points(x, prm, return_raw=True)
    grid, index = orc.process(x, prm)
    np.savez_compressed(
        os.path.join(HERE, "reassign_n512_h128.npz"),
        n_fft=prm.n_fft, hop=prm.hop, gate_db=prm.noise_gate_db, x=x,
        dt_cols=dt.astype(np.float32), dk_bins=dk.astype(np.float32),
        energy=e.astype(np.float32), raw=raw.astype(np.float32),
        grid=grid.astype(np.float32), index=index)
    # analytic KAT table: values any correct implementation must reproduce
    np.savez_compressed(
        os.path.join(HERE, "kats.npz"),
        tone_hz=1000.37, tone_wrong_sign_hz=1015.26, impulse_pos=3884, impulse_wrong_sign=3284,
        chirp_f0=500.0, chirp_f1=8000.0)


if __name__ == "__main__":
    main()
