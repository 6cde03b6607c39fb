"""Parity + timing of an experimental kernel variant (EMS_KERNEL_VARIANT) at n_fft=4096 against the
default kernel and the oracle.  Usage: python tools/variant_check.py <variant> [hop ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "em-spec_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT]
variant = sys.argv[1]
hops = [int(h) for h in sys.argv[2:]] or [128, 256, 1000, 333]
import numpy as np
import torch
import emspec
import reassign_oracle as orc
from parity_util import check_points

SR = 48000


def run(x, hop, var, mode):
    os.environ["EMS_KERNEL_VARIANT"] = var
    eng = emspec.Engine(n_fft=4096, hop=hop, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    xd = torch.from_numpy(x).cuda()
    if mode == "points":
        out = tuple(p[0].cpu().numpy() for p in eng.process_points(xd))
    else:
        g, i = eng.process_grid(xd)
        out = (g[0].cpu().numpy(), i[0].cpu().numpy())
    eng.close()
    return out


ok = True
for hop in hops:
    for sig in ("sparse", "music"):
        x = orc.synth_signal(SR, SR, seed=1) if sig == "sparse" else orc.synth_music(SR, SR, seed=2)
        prm = orc.Params(n_fft=4096, hop=hop)
        try:
            st = check_points(run(x, hop, variant, "points"), x, prm)
            g0, i0 = run(x, hop, "0", "grid")
            g1, i1 = run(x, hop, variant, "grid")
            d = np.abs(i0.astype(int) - i1.astype(int))
            print(f"hop {hop} {sig}: points OK max_dt {st['max_dt_strong']:.2e} p99_dt {st['p99_dt']:.2e} e {st['e_rel_l2']:.2e}; "
                  f"grid vs default kernel: rel {np.linalg.norm(g0 - g1) / np.linalg.norm(g0):.2e}, index diff>1 {(d > 1).mean():.2e}", flush=True)
        except AssertionError as e:
            ok = False
            print(f"hop {hop} {sig}: FAIL {str(e)[:300]}", flush=True)

import bench
S = 600 * SR
pcm = bench.synth_device(S, 0, torch.device("cuda"))
arms = [("0", None), (variant, None)] + [(variant, l) for l in os.environ.get("EMS_AB_LIBS", "").split(",") if l]
for var, lib in arms + arms:
    os.environ["EMS_KERNEL_VARIANT"] = var
    if lib:
        # a second build of the same C-ABI, loaded side by side (ctypes keeps one handle per path)
        import ctypes
        emspec._lib = None
        emspec.LIB_PATH = os.path.abspath(lib)
    else:
        emspec._lib = None
        emspec.LIB_PATH = os.path.join(ROOT, "em-spec_b200", "emspec", "libemspec.so")
    for gate in (-65.0, -200.0):
        eng = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=gate)
        eng.use_torch_stream()
        F = eng.frame_count(S)
        pts = eng.process_points(pcm)
        ms = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.process_points(pcm, out=pts); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        idx = torch.empty((1, F, 2049), dtype=torch.uint8, device="cuda")
        gms = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.process_grid(pcm, out=(None, idx)); e1.record(); torch.cuda.synchronize()
            gms.append(e0.elapsed_time(e1))
        print(f"variant {var} lib {os.path.basename(lib) if lib else 'default'} gate {gate}: points {F / sorted(ms[2:])[2] / 1e3:.1f} M frames/s, grid {F / sorted(gms[2:])[1] / 1e3:.1f} M frames/s", flush=True)
        eng.close()
        del pts, idx
print("PARITY", "OK" if ok else "FAILED")
