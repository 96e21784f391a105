"""Timeline of one tile of k_posterior_fast8 from the event trace (library built with EXTRA=-DF8_TRACING=1, run with
OMBO_FAST_PROFILE=2): per K-block, when the MMA issuer waited for / received A and B, when generator warps were in
which phase, when the cache thread copied.  Usage: python scripts/trace_fast8.py gpurun_out/trace_gw16.txt"""
import sys
from collections import defaultdict

ev = defaultdict(list)
for line in open(sys.argv[1]):
    r, t, c = line.split()
    ev[int(r)].append((int(t), int(c)))
t0 = min(t for r in ev for t, _ in ev[r])
for r in ev:
    ev[r] = [(t - t0, c) for t, c in ev[r]]
print("events per role", {r: len(v) for r, v in ev.items()}, "span", max(t for r in ev for t, _ in ev[r]))

# MMA: 0x80|kb wait start, 0xC0|kb a_full got, kb*4+c unit got b, 0x40|kb committed
print("\nMMA issuer: per block  [wait_a_start, a_got (+wait), units(b_got...), commit]")
cur = None
rows = []
for t, c in ev[1]:
    if c & 0xC0 == 0x80:
        cur = {"kb": c & 63, "ws": t, "units": []}
    elif c & 0xC0 == 0xC0:
        cur["ag"] = t
    elif c & 0xC0 == 0x40:
        cur["cm"] = t
        rows.append(cur)
    else:
        cur["units"].append((c & 3, t))
prev_cm = None
for r in rows:
    s = "kb %2d  t=%7d  wait_a %5d | " % (r["kb"], r["ws"], r["ag"] - r["ws"])
    last = r["ag"]
    for c, t in r["units"]:
        s += "c%d +%d " % (c, t - last)
        last = t
    s += "| issue_done +%d  (block %d cycles)" % (r["cm"] - last, r["cm"] - r["ws"])
    print(s)

def gen(role):
    # kb: start, 0x40|kb slice got, 0x80|kb r2 done (wait a_empty), 0xC0|kb a_empty got, 0x20|kb stores done
    out = []
    cur = {}
    for t, c in ev[role]:
        k = c & 0xE0
        if k == 0x00: cur = {"kb": c & 31, "s": t}
        elif k == 0x40: cur["sl"] = t
        elif k == 0x80: cur["r2"] = t
        elif k == 0xC0: cur["ae"] = t
        elif k == 0x20:
            cur["st"] = t
            out.append(cur)
    return out
for role in (5, 6, 7):
    if role not in ev: continue
    print("\ngenerator role %d: kb  t_start  wait_slice  r2_phase  wait_a_empty  kv+store  | total" % role)
    prev = None
    for g in gen(role):
        print("kb %2d  t=%7d  %5d %5d %5d %5d | %5d   gap_from_prev %s" % (g["kb"], g["s"], g["sl"] - g["s"], g["r2"] - g["sl"],
              g["ae"] - g["r2"], g["st"] - g["ae"], g["st"] - g["s"], (g["s"] - prev) if prev else "-"))
        prev = g["st"]
print("\ncache thread (role 3): 0x00|kb a_written got, 0x40|kb reload issued, 0x80|kb a_empty arrive")
print(" ".join("%s%d@%d" % ({0: "W", 0x40: "R", 0x80: "A"}[c & 0xC0], c & 63, t) for t, c in ev[3]))
print("\nepilogue (role 4): c got / done")
print(" ".join("%s%d@%d" % ("D" if c & 0x40 else "G", c & 63, t) for t, c in ev[4]))
print("\nB producer (role 0): issue times kb*8+c")
print(" ".join("k%dc%d@%d" % (c >> 3, c & 7, t) for t, c in ev[0]))
