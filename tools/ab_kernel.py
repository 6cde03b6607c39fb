"""A/B timing of ems_process_points at n_fft=4096 hop=128 for two builds of libemspec.so.
Usage: python tools/ab_kernel.py libA.so libB.so@ENV=VAL ...   (each timed in its own subprocess, interleaved
twice; `@ENV=VAL` sets an environment variable for that arm, e.g. an experimental dispatch switch)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "em-spec_b200")); sys.path.insert(0, %(root)r)
import torch, emspec, bench
S = 1800 * 48000
pcm = bench.synth_device(S, 0, torch.device("cuda"))
eng = emspec.Engine(n_fft=4096, hop=128)
eng.use_torch_stream()
F = eng.frame_count(S)
pts = eng.process_points(pcm)
ms = []
for it in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.process_points(pcm, out=pts); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms[2:])
idx = torch.empty((1, F, 2049), dtype=torch.uint8, device="cuda")
gms = []
for it in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.process_grid(pcm, out=(None, idx)); e1.record(); torch.cuda.synchronize()
    gms.append(e0.elapsed_time(e1))
gms = sorted(gms[2:])
print("RESULT", "grid", round(F / gms[len(gms) // 2] / 1e3, 2), "points", F / ms[len(ms) // 2] / 1e3, F / ms[0] / 1e3, "checks", float(pts[2].double().sum()), float(pts[0].double().abs().sum()), float(pts[1].double().abs().sum()))
'''


def main():
    libs = sys.argv[1:]
    for rnd in range(2):
        for spec in libs:
            lib, _, kv = spec.partition("@")
            env = dict(os.environ, EMS_LIB_PATH=os.path.abspath(lib))
            if kv:
                k, _, v = kv.partition("=")
                env[k] = v
            out = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
            res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
            print(os.path.basename(spec), res[0] if res else out.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
