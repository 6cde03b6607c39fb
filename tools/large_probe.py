"""Timing of ems_process_points at n_fft = 8192 (hop 2048 and 256) and streaming latency at configs[1]
(8192 / 256 stereo), for A/B runs of the large-n_fft kernels (EMS_LIB_PATH selects the build)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200"), ROOT]
import numpy as np
import torch, emspec, bench
dev = torch.device("cuda")
S = 600 * 48000
pcm = bench.synth_device(S, 0, dev)
for n_fft, hop in ((8192, 2048), (8192, 256), (16384, 4096)):
    eng = emspec.Engine(n_fft=n_fft, hop=hop)
    eng.use_torch_stream()
    F = eng.frame_count(S)
    out = tuple(torch.empty((1, F, n_fft // 2 + 1), dtype=torch.float32, device=dev) for _ in range(3))
    ms = bench.time_calls(lambda: eng.process_points(pcm, out=out), 5, 3)
    gbs = bench.b_points(n_fft, hop) * F / (ms * 1e-3) / 1e9
    print(f"n_fft {n_fft} hop {hop}: {F / ms / 1e3:.2f} M frames/s, {gbs:.0f} GB/s = {gbs / 6554.6:.3f} of measured HBM peak", flush=True)
    eng.close(); del out
seng = emspec.Engine(n_fft=8192, hop=256, channels=2)
col = torch.empty((2, 4097), dtype=torch.uint8, pin_memory=True)
hopbuf = (0.1 * torch.randn(64, 512)).contiguous()
lat = []
for i in range(3200):
    t0 = time.perf_counter(); seng.stream_push(hopbuf[i % 64], col); lat.append(time.perf_counter() - t0)
lat = np.array(lat[200:]) * 1e6
print(f"stream 8192/256 stereo: p50 {np.percentile(lat, 50):.1f} us, p99 {np.percentile(lat, 99):.1f} us")
