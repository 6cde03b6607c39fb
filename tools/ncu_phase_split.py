"""Stall samples per program segment of a kernel: splits the SASS listing of
`ncu -i rep --page source --csv` at its barriers (BAR / SYNCS / WARPSYNC) and sums samples and
executed instructions per segment.  python tools/ncu_phase_split.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seg = 0; segs = collections.OrderedDict()
def cur():
    return segs.setdefault(seg, {"samples": 0, "inst": 0, "n": 0, "first": None, "st": collections.Counter(), "ops": collections.Counter()})
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip(); toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    s = int(r[ix["# Samples"]] or 0); n = int(r[ix["Instructions Executed"]] or 0)
    c = cur(); c["samples"] += s; c["inst"] += n; c["n"] += 1; tot += s
    c["ops"][op.split(".")[0]] += n
    if c["first"] is None: c["first"] = src
    for k in stall_cols: c["st"][k[6:]] += int(r[ix[k]] or 0)
    if op.startswith("BAR") or (op.startswith("WARPSYNC") and n > 100000):
        seg += 1
for k, c in segs.items():
    if c["samples"] * 200 < tot: continue
    top = ", ".join(f"{a}={100*b/max(c['samples'],1):.0f}%" for a, b in c["st"].most_common(4))
    ops = ", ".join(f"{a}:{b//1000}k" for a, b in c["ops"].most_common(6))
    print(f"seg {k:3d}: {100*c['samples']/tot:5.1f}% samples, {c['inst']/1e6:8.1f} M warp-instr, {c['n']:5d} SASS | {top} | {ops}")
