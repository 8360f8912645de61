#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.OrderedDict()
hdr = None
cur_file = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ii: continue
    try: inst = float(r[ii] or 0); smp = float(r[sa] or 0)
    except ValueError: continue
    key = (cur_file, r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0.0, 0.0]); a[0] += inst; a[1] += smp
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst %.3g, samples %d" % (ti, ts))
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s:%s  %s" % (100 * i / ti, 100 * s / ts, f, ln, src))
