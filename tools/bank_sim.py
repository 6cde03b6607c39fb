"""Bank-conflict check of the tuned kernel's shared-memory layouts (8-byte elements).
A warp LDS.64/STS.64 is served per half-warp; a half-warp is conflict-free when its 16
lanes touch 16 distinct bank pairs (or identical addresses)."""
import itertools

def wavefronts(addrs):  # addrs: 32 element indices (float2 units)
    tot = 0
    for h in (addrs[:16], addrs[16:]):
        banks = {}
        for a in h:
            banks.setdefault(a % 16, set()).add(a)
        tot += max(len(v) for v in banks.values())
    return tot

padZ = lambda a: a + (a >> 8)
padY = lambda a: a + (a >> 7)

def report(name, fn):
    worst = 0; total = 0; n = 0
    for warp in range(4):
        for slot in fn.slots:
            w = wavefronts([fn(32 * warp + l, slot) for l in range(32)])
            worst = max(worst, w); total += w; n += 1
    print(f"{name:28s} worst {worst} avg {total / n:.2f} (ideal 2)")

class F:
    def __init__(s, f, slots): s.f = f; s.slots = slots
    def __call__(s, p, slot): return s.f(p, slot)

def tA(p): return p if p else 0
def tB(p): return 256 - p if p else 128

for u in (0, 1):
    report(f"Z pass1 st u={u}", F(lambda p, j, u=u: padZ(p + 128 * u + 256 * j), range(16)))
    report(f"Z pass2 ld/st u={u}", F(lambda p, j, u=u: padZ(256 * (p & 15) + ((p >> 4) + 8 * u) + 16 * j), range(16)))
report("Z pass3 ld A", F(lambda p, j: padZ(256 * (tA(p) & 15) + 16 * (tA(p) >> 4) + j), range(16)))
report("Z pass3 ld B", F(lambda p, j: padZ(256 * (tB(p) & 15) + 16 * (tB(p) >> 4) + j), range(16)))
report("Y pass1 st", F(lambda p, j: padY(p + 128 * j), range(16)))
report("Y pass2 ld/st", F(lambda p, j: padY(128 * (p & 15) + (p >> 4) + 8 * j), range(16)))
report("Y pass3 ld A", F(lambda p, j: padY(128 * (tA(p) & 15) + 8 * (tA(p) >> 4) + j), range(8)))
report("Y pass3 ld B", F(lambda p, j: padY(128 * (tB(p) & 15) + 8 * (tB(p) >> 4) + j), range(8)))
report("Ztab ld u=0", F(lambda p, i: 256 * i + p, range(1, 16)))
report("Ztab ld u=1", F(lambda p, i: 256 * i + p + 128, range(1, 16)))
report("Ytab ld", F(lambda p, i: 128 * i + p, range(1, 16)))
report("T2 ld u=0", F(lambda p, i: 16 * i + (p >> 4), range(1, 16)))
report("T2Y ld", F(lambda p, i: 8 * i + (p >> 4), range(1, 16)))
