"""Stage times (CUDA events inside the library) of ems_process_grid on the bench stream for the sparse,
dense (gate -200 dB) and broadband cases; deterministic and fast accumulators."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200"), ROOT]
import torch, emspec, bench
dev = torch.device("cuda")
S = int(float(sys.argv[1]) if len(sys.argv) > 1 else 1800) * 48000
pcm = bench.synth_device(S, 0, dev)
bb = bench.synth_broadband_device(S, 0, dev)
F = bench.frame_count(S, 4096, 128)
idx = torch.empty((1, F, 2049), dtype=torch.uint8, device=dev)
for name, x, gate in (("sparse", pcm, -65.0), ("dense", pcm, -200.0), ("broadband", bb, -65.0)):
  for fused in os.environ.get("PROBE_FUSED", "0,1").split(","):
    os.environ["EMS_FUSED_POST"] = fused
    for det in (1, 0):
          fl = emspec.FLAG_REASSIGN | (emspec.FLAG_DETERMINISTIC if det else 0)
          eng = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=gate, flags=fl)
          eng.use_torch_stream()
          for _ in range(3):
              eng.process_grid(x, out=(None, idx))
          torch.cuda.synchronize()
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record(); eng.process_grid(x, out=(None, idx)); e1.record(); torch.cuda.synchronize()
          print(f"{name:10s} fused={fused} ring={os.environ.get('EMS_FUSED_RING','8192')} {'u64' if det else 'f32'}: total {e0.elapsed_time(e1):7.2f} ms = {F / e0.elapsed_time(e1) / 1e3:6.1f} M frames/s | "
                f"stft+deposit {eng.stage_ms(0):7.2f} ms, post {eng.stage_ms(2):7.2f} ms | nonzero {float(torch.count_nonzero(idx)) / idx.numel():.3f}", flush=True)
          eng.close()
