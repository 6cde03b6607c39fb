import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("em-spec_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import emspec, reassign_oracle as orc
SR = 48000
x = orc.synth_signal(SR, SR, seed=4)
for n_fft, hop in ((1024, 64), (2048, 128)):
    prm = orc.Params(n_fft=n_fft, hop=hop)
    grid_o, idx_o = orc.process(x, prm)
    for force in ("0", "1"):
        os.environ["EMS_FORCE_GENERIC"] = force
        eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=prm.flags | emspec.FLAG_SYNC)
        g, idx = eng.process_grid(torch.from_numpy(x).cuda())
        g = g[0].cpu().numpy(); idx = idx[0].cpu().numpy()
        eng.close()
        d = np.abs(idx.astype(int) - idx_o.astype(int))
        bad = np.argwhere(d > 1)
        print(n_fft, hop, "force_generic", force, "cells d>1:", len(bad), "rel_l2", np.linalg.norm(g - grid_o) / np.linalg.norm(grid_o))
        for (c, k) in bad[:12]:
            print("   col", c, "row", k, "gpu", g[c, k], "oracle", grid_o[c, k], "idx", idx[c, k], idx_o[c, k],
                  "nbrs gpu", g[max(c-1,0):c+2, max(k-1,0):k+2].round(9).tolist(), "orc", grid_o[max(c-1,0):c+2, max(k-1,0):k+2].round(9).tolist())
