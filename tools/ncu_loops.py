#!/usr/bin/env python3
"""Loop-level attribution of one kernel from an ncu report: consecutive SASS lines with (nearly) equal execution counts are merged
into runs; per run: instruction share, sample share, dominant opcodes, top stall reasons.
usage: ncu_loops.py report.ncu-rep kernel_regex [min_share_pct]"""
import csv, subprocess, sys
from collections import Counter
rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        hdr = r; start = i + 1; break
si, ii, sa = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [c for c in hdr if c.startswith("stall_") and "(Not Issued)" not in c]
sidx = [hdr.index(c) for c in stalls]
def f(x):
    try: return float(x or 0)
    except ValueError: return 0.0
def op(src):
    t = src.split()
    return (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
data = [(r[si], f(r[ii]), f(r[sa]), [f(r[j]) for j in sidx]) for r in rows[start:] if len(r) > max(sidx + [sa])]
tot = sum(d[1] for d in data) or 1; ts = sum(d[2] for d in data) or 1
runs = []
for n, (src, i, s, st) in enumerate(data):
    if runs and i > 0 and abs(runs[-1]["c"] - i) <= 0.25 * max(runs[-1]["c"], i):
        R = runs[-1]; R["n"] += 1; R["i"] += i; R["s"] += s; R["end"] = n; R["ops"].append(op(src)); R["st"] = [a + b for a, b in zip(R["st"], st)]
    else:
        runs.append(dict(start=n, end=n, c=i, n=1, i=i, s=s, ops=[op(src)], st=list(st)))
print("total warp-inst %.4g, samples %d" % (tot, ts))
for R in runs:
    if 100 * R["i"] / tot >= min_share or 100 * R["s"] / ts >= min_share:
        c = " ".join("%s:%d" % x for x in Counter(R["ops"]).most_common(5))
        top = " ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(R["s"], 1)) for k, v in sorted(zip(stalls, R["st"]), key=lambda x: -x[1])[:3])
        print("%4d-%4d n=%3d cnt %.3g inst %4.1f%% samp %4.1f%% | %s | %s" % (R["start"], R["end"], R["n"], R["c"], 100 * R["i"] / tot, 100 * R["s"] / ts, c, top))
