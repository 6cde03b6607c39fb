"""Timing of ems_process_points at n_fft = 256 ... 4096, hop = n_fft / 4 (the sweep's small end), for A/B runs
of library builds (EMS_LIB_PATH selects the build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200"), ROOT]
import torch, emspec, bench
dev = torch.device("cuda")
S = 600 * 48000
pcm = bench.synth_device(S, 0, dev)
for n_fft, hop in ((256, 64), (512, 128), (1024, 256), (2048, 512), (4096, 1024), (4096, 128), (1024, 100)):
    eng = emspec.Engine(n_fft=n_fft, hop=hop)
    eng.use_torch_stream()
    F = eng.frame_count(S)
    out = tuple(torch.empty((1, F, n_fft // 2 + 1), dtype=torch.float32, device=dev) for _ in range(3))
    ms = bench.time_calls(lambda: eng.process_points(pcm, out=out), 5, 3)
    gbs = bench.b_points(n_fft, hop) * F / (ms * 1e-3) / 1e9
    print(f"n_fft {n_fft} hop {hop}: {F / ms / 1e3:.2f} M frames/s, {gbs:.0f} GB/s = {gbs / 6554.6:.3f} of measured HBM peak", flush=True)
    eng.close(); del out
