"""n_fft sweep throughput probe (configs[4] geometry: hop = n_fft/4, display controls on).
Times ems_process_points (store) and ems_process_grid (fused deposit + post-pass) per n_fft with
CUDA events on torch's stream.  Usage: python tools/sweep_probe.py [seconds_of_audio]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "em-spec_b200"))
import torch
import emspec

SR = 48000


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    S = int(secs * SR)
    g = torch.Generator(device="cuda").manual_seed(0)
    t = torch.arange(S, device="cuda", dtype=torch.float64) / SR
    x = (0.5 * torch.sin(2 * torch.pi * (20.0 * t + (20000.0 - 20.0) / (2 * secs) * t * t))
         + 0.25 * torch.sin(2 * torch.pi * 440.0 * t)).float()
    x += 1e-3 * torch.randn(S, device="cuda", generator=g)
    out = []
    for n_fft in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
        hop = n_fft // 4
        row = {"n_fft": n_fft, "hop": hop}
        for force in ("0", "1"):
            os.environ["EMS_FORCE_GENERIC"] = force
            eng = emspec.Engine(n_fft=n_fft, hop=hop, low_end_boost=3.9, smoothing=0.0, noise_gate_db=-65.0)
            eng.use_torch_stream()
            F = eng.frame_count(S)
            pts = None
            for mode in ("points", "grid"):
                ms = []
                for it in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    if mode == "points":
                        pts = eng.process_points(x, out=pts)
                    else:
                        eng.process_grid(x, want_grid=False)
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                best = min(ms[1:])
                row[f"{mode}_{'generic' if force == '1' else 'tuned'}_Mfps"] = round(F / best / 1e3, 2)
            pts = None
            eng.close()
            torch.cuda.empty_cache()
        row["frames"] = F
        row["points_GBs_tuned"] = round(row["points_tuned_Mfps"] * 1e6 * (4 * hop + 12 * (n_fft // 2 + 1)) / 1e9, 1)
        out.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
