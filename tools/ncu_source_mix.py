"""Instruction mix and stall hot spots from `ncu --page source --csv` output: python tools/ncu_source_mix.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LDS", "STS", "LDG", "STG", "BAR", "MUFU")) else "")
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    mix[op] += n; samp[op] += s; tot += n; tots += s
    for c in stall_cols:
        stalls[c] += int(r[ix[c]] or 0)
    lines.append((s, n, src))
print(f"total warp-instr {tot}, samples {tots}")
for op, n in mix.most_common(28):
    print(f"  {op:22s} {n:12d} {100*n/tot:6.2f}%   samples {100*samp[op]/tots:6.2f}%")
print("stalls:", ", ".join(f"{k[6:]}={100*v/tots:.1f}%" for k, v in stalls.most_common(10)))
print("hottest instructions:")
for s, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"  {100*s/tots:5.2f}%  x{n:9d}  {src}")
