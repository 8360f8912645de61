#!/usr/bin/env python3
"""Executed warp-instructions and stall samples of one kernel, split at BAR.SYNC / CALL boundaries and per opcode.
usage: ncu_phases.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; data = []
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r; si = r.index("Source"); ii = r.index("Instructions Executed"); sa = r.index("# Samples"); continue
    if hdr is None or len(r) <= ii: continue
    try: n = float(r[ii] or 0); s = float(r[sa] or 0)
    except ValueError: continue
    data.append((r[si], n, s))
tot = sum(d[1] for d in data) or 1; ts = sum(d[2] for d in data) or 1
acc = [0, 0, 0]; start = 0
print("total warp-inst %.4g, samples %d, %d SASS instructions" % (tot, ts, len(data)))
for i, (src, n, s) in enumerate(data):
    acc[0] += n; acc[1] += s; acc[2] += 1
    if "BAR.SYNC" in src or "RET." in src or i == len(data) - 1 or "WARPSYNC" in src:
        if acc[0] / tot > 0.004 or acc[1] / ts > 0.004:
            print("sass %4d-%4d  inst %5.1f%%  samples %5.1f%%  (%d instrs) ends with %s" % (start, i, acc[0] / tot * 100, acc[1] / ts * 100, acc[2], src.split(";")[0][:40]))
        acc = [0, 0, 0]; start = i + 1
