"""One launch of the fused kernel for ncu: python tools/ncu_target.py [n_fft hop seconds mode]
mode = points | grid.  Run it once without ncu first (it must exit 0), then under
  ncu --set full --clock-control none --import-source on -k regex:stft_reassign -c 1 -o gpurun_out/x python tools/ncu_target.py ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "em-spec_b200"))
sys.path.insert(0, ROOT)
import torch
import emspec
import bench

n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 128
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 600.0
mode = sys.argv[4] if len(sys.argv) > 4 else "points"
S = int(secs * 48000)
pcm = bench.synth_device(S, 0, torch.device("cuda"))
gate = float(os.environ.get("EMS_NCU_GATE", "-65"))
eng = emspec.Engine(n_fft=n_fft, hop=hop, noise_gate_db=gate, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
if mode == "points":
    pts = eng.process_points(pcm)
    print("frames", pts[0].shape[1], "energy sum", float(pts[2].sum()))
else:
    _, idx = eng.process_grid(pcm, want_grid=False)
    print("frames", idx.shape[1], "index sum", int(idx.sum()))
eng.close()
