#!/usr/bin/env python3
"""Executed warp-instructions per SASS opcode from an ncu report.  usage: ncu_opcodes.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; agg = collections.Counter(); smp = collections.Counter()
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r; si = r.index("Source"); ii = r.index("Instructions Executed"); sa = r.index("# Samples"); continue
    if hdr is None or len(r) <= ii: continue
    try: n = float(r[ii] or 0); s = float(r[sa] or 0)
    except ValueError: continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    agg[op] += n; smp[op] += s
tot = sum(agg.values()) or 1; ts = sum(smp.values()) or 1
print("total warp-inst %.4g" % tot)
for op, n in agg.most_common(top): print("%6.2f%% inst %6.2f%% samples  %s" % (100 * n / tot, 100 * smp[op] / ts, op))
