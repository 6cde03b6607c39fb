"""Writes profiles/r02_kernel_profile.json from an .ncu-rep of the dominant kernel (read on the CPU box):
    python tools/ncu_profile_json.py rep.ncu-rep <frames in the captured launch> [out.json]
DRAM bytes and executed fp32 flops per frame (2 x FFMA + FADD + FMUL at thread level, packed forms
count twice; taken from the per-instruction "Predicated-On Thread Instructions Executed" column of
the source page), shared-memory wavefronts per frame, and the sha256 of the kernel's sources (common.cuh, stft_generic.cuh, stft_r16.cuh) at the
time of writing.  bench.py reads the file at run time and marks it stale when the sources changed."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, frames = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_kernel_profile.json")


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


raw = page("raw")
d = dict(zip(raw[0], raw[2]))
u = dict(zip(raw[0], raw[1]))


def val(k):
    v = float(d[k].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u[k], 1.0)


src = page("source")
hdr = src[1]
ix = {h: i for i, h in enumerate(hdr)}
FLOPS = {"FFMA": 2, "FFMA2": 4, "FADD": 1, "FADD2": 2, "FMUL": 1, "FMUL2": 2}
flops = 0
mix = {}
for r in src[2:]:
    if len(r) < len(hdr):
        continue
    toks = r[ix["Source"]].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n = int(r[ix["Predicated-On Thread Instructions Executed"]] or 0)
    mix[op] = mix.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
    flops += FLOPS.get(op, 0) * n
hsh = hashlib.sha256()
csrc = os.path.join(ROOT, "em-spec_b200", "csrc")
for fn in ("common.cuh", "stft_generic.cuh", "stft_r16.cuh"):      # the sources of the profiled kernel
    hsh.update(open(os.path.join(csrc, fn), "rb").read())
res = {
    "kernel": d["Kernel Name"], "frames": frames, "source_report": os.path.basename(rep),
    "gpu_time_us": val("gpu__time_duration.sum") if u["gpu__time_duration.sum"] == "us" else d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"],
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_bytes_per_frame": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / frames,
    "flops_per_frame": flops / frames,
    "smem_wavefronts_per_frame": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", "")) / frames,
    "warp_instructions_per_frame": sum(mix.values()) / frames,
    "registers_per_thread": int(d["launch__registers_per_thread"]),
    "sass_mix_top": dict(sorted(mix.items(), key=lambda kv: -kv[1])[:12]),
    "csrc_sha16": hsh.hexdigest()[:16],
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
