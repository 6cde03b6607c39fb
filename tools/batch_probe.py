"""Throughput at the configs[3] geometry (many 60 s clips as planar channels, n_fft 4096, hop 256)."""
import os, sys
sys.path.insert(0, "em-spec_b200"); sys.path.insert(0, ".")
import torch, emspec, bench
C, secs = 64, 60
S = secs * 48000
pcm = torch.stack([bench.synth_device(S, c, torch.device("cuda")) for c in range(C)])
for mode in ("points", "grid"):
    eng = emspec.Engine(n_fft=4096, hop=256, channels=C)
    eng.use_torch_stream()
    F = eng.frame_count(S)
    out = None
    ms = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "points":
            out = eng.process_points(pcm, out=out)
        else:
            out = eng.process_grid(pcm, want_grid=False, out=out)
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    best = min(ms[1:])
    print(f"configs[3] geometry, {C} clips x {secs} s as channels, hop 256, {mode}: {C*F/best/1e3:.1f} M frames/s ({best:.2f} ms)")
    eng.close(); out = None; torch.cuda.empty_cache()
