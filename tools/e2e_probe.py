import os, sys, time
sys.path.insert(0, 'em-spec_b200'); sys.path.insert(0, '.')
import torch, emspec, bench
S = 3600 * 48000
pcm = bench.synth_device(S, 0, torch.device('cuda'))
ph = torch.empty((1, S), dtype=torch.float32, pin_memory=True); ph.copy_(pcm[None])
for chunk in (8192, 16384, 32768, 65536, 131072):
    os.environ["EMS_HOST_CHUNK_FRAMES"] = str(chunk)
    eng = emspec.Engine(n_fft=4096, hop=128)
    F = eng.frame_count(S)
    ih = torch.empty((1, F, 2049), dtype=torch.uint8, pin_memory=True)
    for _ in range(2): eng.process_host(ph, index_out=ih)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): eng.process_host(ph, index_out=ih)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(chunk, f"{dt*1e3:.1f} ms  {F/dt/1e6:.1f} M frames/s")
    eng.close()
