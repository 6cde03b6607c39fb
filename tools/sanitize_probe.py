"""Tiny workload for compute-sanitizer (one tool per gpurun call): tuned 4096 kernel in all
three modes, a generic size, the 32768 kernel, the post-pass variants and a few stream pushes."""
import sys
sys.path.insert(0, "em-spec_b200"); sys.path.insert(0, "oracle")
import numpy as np, torch, emspec, reassign_oracle as orc
x = torch.from_numpy(orc.synth_signal(24000, 48000.0, seed=1)).cuda()
for n_fft, hop in ((4096, 128), (1024, 256), (32768, 8192)):
    xx = x if n_fft < 32768 else torch.from_numpy(orc.synth_signal(3 * 32768, 48000.0, seed=1)).cuda()
    for flags in (3, 1):
        for smoothing, agc in ((0.0, 0.0), (0.5, 0.7)):
            eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=flags | 4, smoothing=smoothing, agc_strength=agc)
            eng.process_points(xx); eng.process_grid(xx)
            eng.close()
eng = emspec.Engine(n_fft=4096, hop=128, display_rows=300)
col = torch.empty((1, 300), dtype=torch.uint8).pin_memory()
xh = x.cpu()
for i in range(60):
    eng.stream_push(xh[i * 128:(i + 1) * 128].contiguous(), col)
eng.close()
print("probe done")
