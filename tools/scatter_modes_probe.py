"""ems_scatter_points in its three modes on the points of a 4096/128 stream (the bench's scatter_modes key as a
stand-alone probe, e.g. under ncu):  python tools/scatter_modes_probe.py [seconds] [gate_db] [mode ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200"), ROOT]
import torch, emspec, bench
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gate = float(sys.argv[2]) if len(sys.argv) > 2 else -200.0
modes = sys.argv[3:] or ["u64", "f32", "sorted"]
dev = torch.device("cuda")
S = secs * 48000
pcm = bench.synth_device(S, 0, dev)
src = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=gate)
src.use_torch_stream()
pts = src.process_points(pcm)
F = pts[2].shape[1]
src.close()
idx = torch.empty((1, F, 2049), dtype=torch.uint8, device=dev)
FL = {"u64": emspec.FLAG_DETERMINISTIC, "f32": 0, "sorted": emspec.FLAG_DETERMINISTIC | emspec.FLAG_SORTED_SCATTER}
for m in modes:
    e = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=gate, flags=emspec.FLAG_REASSIGN | FL[m])
    e.use_torch_stream()
    fn = lambda: e.lib.ems_scatter_points(e.h, pts[0].data_ptr(), pts[1].data_ptr(), pts[2].data_ptr(), F, None, idx.data_ptr())
    ms = bench.time_calls(fn, 3, 1)
    print(f"{m}: gate {gate} dB, {F} frames, {int(torch.count_nonzero(pts[2]))} kept points: {ms:.3f} ms per call, "
          f"deposit stage {e.stage_ms(emspec.STAGE_SCATTER):.3f} ms, {F / ms / 1e3:.1f} M frames/s", flush=True)
    e.close()
