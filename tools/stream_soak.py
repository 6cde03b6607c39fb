"""Streaming soak: 200,000 pushes at configs[1] (8192 / 256, stereo, smoothing + AGC) with the hop format switching every
1000 pushes, a colour map set and cleared on the way, pixels checked against the index column, a checkpoint saved and
reloaded every 40,000 pushes.  Checks consecutive column indices and reports the time per push and the device-memory delta
(last run: 33.5 us per push including the Python call, 199,953 columns, 6 MiB)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200")]
import numpy as np, torch, emspec
eng = emspec.Engine(n_fft=8192, hop=256, channels=2, smoothing=0.3, agc_strength=0.5)
col = torch.empty((2, 4097), dtype=torch.uint8).pin_memory()
px = torch.empty((2, 4097), dtype=torch.int32)
rng = np.random.default_rng(0)
hops = torch.from_numpy((0.1 * rng.standard_normal((64, 512))).astype(np.float32))
hops16 = (hops * 32768).round().clamp(-32768, 32767).to(torch.int16)
lut = np.arange(256, dtype=np.uint32) * 0x010101
free0 = torch.cuda.mem_get_info()[0]
t0 = time.time(); last = -1; n = 0
for i in range(200000):
    if i == 50000: eng.stream_set_colormap(lut)
    if i == 150000: eng.stream_set_colormap(None)
    if i % 40000 == 39999:
        blob = eng.stream_save(); eng.stream_load(blob)
    src = hops16[i % 64] if (i // 1000) % 2 else hops[i % 64]
    r, ci = eng.stream_push(src, col)
    if r:
        assert ci == last + 1 or last == -1, (ci, last)
        last = ci; n += 1
        if 50000 <= i < 150000 and i % 997 == 0:
            eng.stream_column_rgba(px)
            assert (px.numpy().view(np.uint32) == lut[col.numpy()]).all()
dt = time.time() - t0
free1 = torch.cuda.mem_get_info()[0]
print(f"200000 pushes in {dt:.1f} s ({dt / 200000 * 1e6:.1f} us each), {n} columns, last index {last}, device memory delta {(free0 - free1) / 2**20:.1f} MiB")
eng.close()
