"""Smallest workload for one compute-sanitizer tool (racecheck on the tile hand-over of the tiled kernel):
n_fft 4096, hop 128, 3 s of audio, store and deposit modes, plus the single-exchange variant."""
import os, sys
sys.path.insert(0, "em-spec_b200"); sys.path.insert(0, "oracle")
import torch, emspec, reassign_oracle as orc
x = torch.from_numpy(orc.synth_signal(3 * 48000, 48000.0, seed=1)).cuda()
for var in ("0", "64"):
    os.environ["EMS_KERNEL_VARIANT"] = var
    eng = emspec.Engine(n_fft=4096, hop=128, flags=3 | 4)
    eng.process_points(x); eng.process_grid(x)
    eng.close()
print("probe done")
