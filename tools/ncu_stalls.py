#!/usr/bin/env python3
"""Per-phase stall breakdown of one kernel from the ncu source page (SASS view): phases are split at BAR.SYNC / RET / WARPSYNC.
usage: ncu_stalls.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        hdr = r; start = i + 1; break
si, ii, sa = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [c for c in hdr if c.startswith("stall_") and "(Not Issued)" not in c]
sidx = [hdr.index(c) for c in stalls]
def f(x):
    try: return float(x or 0)
    except ValueError: return 0.0
data = [(r[si], f(r[ii]), f(r[sa]), [f(r[j]) for j in sidx]) for r in rows[start:] if len(r) > max(sidx)]
tot_i = sum(d[1] for d in data) or 1; tot_s = sum(d[2] for d in data) or 1
print("total warp-inst %.4g, samples %d" % (tot_i, tot_s))
acc_i = acc_s = 0; acc_st = [0.0] * len(stalls); first = 0
for n, (src, i, s, st) in enumerate(data):
    acc_i += i; acc_s += s; acc_st = [a + b for a, b in zip(acc_st, st)]
    if "BAR.SYNC" in src or "RET." in src or "WARPSYNC" in src or n == len(data) - 1:
        if acc_s / tot_s > 0.01:
            top = sorted(zip(stalls, acc_st), key=lambda x: -x[1])[:4]
            print("sass %4d-%4d inst %5.1f%% samples %5.1f%%  %s" % (first, n, 100 * acc_i / tot_i, 100 * acc_s / tot_s,
                  "  ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(acc_s, 1)) for k, v in top)))
        acc_i = acc_s = 0; acc_st = [0.0] * len(stalls); first = n + 1
