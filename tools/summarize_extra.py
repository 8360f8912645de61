#!/usr/bin/env python3
"""Per-kernel ncu --set full metrics of the kernels outside the step chain (tools/profile_extra.sh) -> profiles/<tag>_extra_kernels.csv
usage: summarize_extra.py <tag> <extra.ncu-rep> [output suffix, default _extra_kernels.csv]"""
import csv, os, re, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
hdr, units = rr[0], rr[1]
idx = {c: i for i, c in enumerate(hdr)}
rows = [["kernel"] + ["%s [%s]" % (w, units[idx[w]]) for w in want if w in idx]]
seen = {}
for r in rr[2:]:
    n = re.sub(r"<.*$", "", re.sub(r"^void\s+", "", r[idx["Kernel Name"]].split("(")[0].strip()))
    seen[n] = seen.get(n, 0) + 1
    if seen[n] > (4 if n.startswith("k_fast") else 2) and not n.startswith("k_resize_exact"):
        continue
    rows.append(["%s#%d" % (n, seen[n])] + [r[idx[w]] for w in want if w in idx])
csv.writer(open(os.path.join(root, tag + (sys.argv[3] if len(sys.argv) > 3 else "_extra_kernels.csv")), "w")).writerows(rows)
for r in rows[1:]:
    print(r[0].ljust(22), " ".join(x[:10].rjust(11) for x in r[1:10]))
