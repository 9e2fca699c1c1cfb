"""Merge the per-unit events of a trace run (tools/fa_trace.py with the trace build) into one timeline."""
import re, sys
ev = []
for ln in open(sys.argv[1]):
    if ln.startswith("----"):
        ev.append((0, ln.strip())); continue
    m = re.match(r"umma unit +(\d+): P seen +(\d+) +S issue +(\d+) \.\. +(\d+)", ln)
    if m:
        u = int(m.group(1))
        ev += [(int(m.group(2)), f"  umma u{u} p_full seen"), (int(m.group(3)), f"  umma u{u} S issue begin"),
               (int(m.group(4)), f"  umma u{u} S issue end+commit")]
    m = re.match(r"epi warp +(\d+) unit +(\d+): wait from +(\d+) +S seen +(\d+) +arithmetic done +(\d+) +P arrived +(\d+)", ln)
    if m and m.group(1) == "0":
        u = int(m.group(2))
        ev += [(int(m.group(3)), f"epi0 u{u} wait begins"), (int(m.group(4)), f"epi0 u{u} S seen"),
               (int(m.group(5)), f"epi0 u{u} arith done"), (int(m.group(6)), f"epi0 u{u} P arrived")]
groups, cur = [], None
for t, e in ev:
    if t == 0:
        cur = []; groups.append((e, cur))
    else:
        cur.append((t, e))
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 6, int(sys.argv[3]) if len(sys.argv) > 3 else 11
for name, c in groups:
    print(name)
    c.sort()
    t0 = next(t for t, e in c if f"u{lo} " in e)
    for t, e in c:
        if lo <= int(re.search(r"u(\d+) ", e).group(1)) <= hi:
            print(f"{t - t0:7d} {e}")
