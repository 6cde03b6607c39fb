"""Several kernels of the family in one process, for one ncu pass over all of them:
  ncu --set full --clock-control none --import-source on -k regex:"stft_reassign|post_|scatter" -c 16 -o gpurun_out/x python tools/ncu_multi.py
points at n_fft 1024, 2048, 16384, 32768 (hop = n_fft / 4), then the image path (deposit + post-pass) at 4096 / 128,
sparse and with every bin kept, then scatter of caller-held points."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "em-spec_b200")); sys.path.insert(0, ROOT)
import torch, emspec, bench
S = 120 * 48000
pcm = bench.synth_device(S, 0, torch.device("cuda"))
fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
for n_fft in (1024, 2048, 16384, 32768):
    eng = emspec.Engine(n_fft=n_fft, hop=n_fft // 4, flags=fl)
    pts = eng.process_points(pcm)
    print(n_fft, "frames", pts[0].shape[1], float(pts[2].sum()))
    eng.close(); del pts
for gate in (-65.0, -200.0):
    eng = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=gate, flags=fl)
    g, idx = eng.process_grid(pcm)
    print("grid gate", gate, int(idx.sum()))
    if gate == -65.0:
        pts = eng.process_points(pcm)
        g2, i2 = eng.scatter_points(*pts)
        print("scatter equal", bool(torch.equal(i2, idx)))
    eng.close()
