"""Streaming push latency (configs[1] geometry) and, under ncu, the per-kernel durations of one push.
python tools/stream_probe.py [n_fft hop channels pushes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "em-spec_b200"))
import numpy as np, torch, emspec
n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pushes = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
eng = emspec.Engine(n_fft=n_fft, hop=hop, channels=ch)
col = torch.empty((ch, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
hb = (0.1 * torch.randn(64, hop * ch)).contiguous()
lat = []
for i in range(200 + pushes):
    t0 = time.perf_counter(); eng.stream_push(hb[i % 64], col); lat.append(time.perf_counter() - t0)
lat = np.array(lat[200:]) * 1e6
print(f"n_fft={n_fft} hop={hop} ch={ch}: p50 {np.percentile(lat,50):.1f} us  p99 {np.percentile(lat,99):.1f} us  min {lat.min():.1f} us")
eng.close()
