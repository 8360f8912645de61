#!/usr/bin/env python3
"""Summarise ncu outputs into profiles/ (tracked).  usage: summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [bench.json]
Writes profiles/<tag>_launches.csv (kernel, grid, block, ns per launch), profiles/<tag>_kernels.csv (per-kernel ncu --set full
metrics) and profiles/<tag>_summary.md."""
import csv, json, subprocess, sys, collections, os
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
bench = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else None
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
os.makedirs(root, exist_ok=True)
import re


def kname(full):
    """'void k_fast_cells<80>(LevelMaps, ...)' -> 'k_fast_cells'"""
    n = full.split("(")[0].strip()
    n = re.sub(r"^void\s+", "", n)
    return re.sub(r"<.*$", "", n)


rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
out = [("id", "kernel", "block", "grid", "gpu__time_duration_ns")]
agg = collections.OrderedDict()
for r in rows:
    name = kname(r[4])
    out.append((r[0], name, r[7], r[8], r[-1]))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
csv.writer(open(os.path.join(root, tag + "_launches.csv"), "w")).writerows(out)
total = sum(a[1] for a in agg.values()) or 1
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
hdr, units = rr[0], rr[1]
idx = {c: i for i, c in enumerate(hdr)}
krows = [["kernel"] + ["%s [%s]" % (w, units[idx[w]]) for w in want if w in idx]]
seen = {}
for r in rr[2:]:
    name = kname(r[idx["Kernel Name"]])
    seen[name] = seen.get(name, 0) + 1
    krows.append(["%s#%d" % (name, seen[name])] + [r[idx[w]] for w in want if w in idx])
csv.writer(open(os.path.join(root, tag + "_kernels.csv"), "w")).writerows(krows)
# DRAM traffic per frame of every captured kernel (bench.py reports it as roofline.traffic, scaled to its batch)
FRAMES = 32                                       # tools/profile_round.sh captures bench.py --batch 32
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for r in rr[2:]:
    name = kname(r[idx["Kernel Name"]])
    b = sum(float(r[idx[m]]) * scale.get(units[idx[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    t = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "warp_inst": 0.0})
    t["launches"] += 1; t["dram_bytes"] += b
    if "smsp__inst_executed.sum" in idx:
        t["warp_inst"] += float(r[idx["smsp__inst_executed.sum"]])
for name, t in traffic.items():
    t["dram_bytes_per_frame_per_step"] = t["dram_bytes"] / FRAMES      # summed over the launches of one step
    t["warp_inst_per_frame_per_step"] = t["warp_inst"] / FRAMES        # executed warp instructions (bench.py: issue-rate view)
    t["capture_frames"] = FRAMES
json.dump(traffic, open(os.path.join(root, tag + "_traffic.json"), "w"), indent=1)
with open(os.path.join(root, tag + "_summary.md"), "w") as f:
    f.write("# %s — ncu summary\n\n" % tag)
    f.write("Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over one timed step of `bench.py --kernels-only` "
            "(cold-cache, serialised: compare SHARES).  Full metrics: `ncu --set full --clock-control none --import-source on`.\n\n")
    f.write("| kernel | launches | total us | share of step |\n|---|---|---|---|\n")
    for k, (n, ns) in agg.items():
        f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, n, ns / 1e3, 100 * ns / total))
    if bench:
        f.write("\nCUDA-event shares from the un-profiled bench run of the same build (`bench.py`, %d frames/step):\n\n| kernel | ms/step | share | achieved GB/s | frac of measured HBM peak |\n|---|---|---|---|---|\n" % bench["config"]["frames_per_step_per_gpu"])
        for k, v in bench["kernels"].items():
            f.write("| %s | %.3f | %.1f %% | %s | %s |\n" % (k, v["ms_per_step"], 100 * v["share"], "%.0f" % v["achieved_gbs"] if "achieved_gbs" in v else "-", "%.3f" % v["frac_of_hbm"] if "frac_of_hbm" in v else "-"))
        f.write("\nvalue %.0f frames/s device-resident, e2e %.0f frames/s, roofline: %s\n" % (bench["value"], bench["e2e"]["value"], json.dumps(bench["roofline"])))
    f.write("\nPer-kernel `--set full` metrics are in `%s_kernels.csv` (one row per captured launch).\n" % tag)
print("wrote profiles/%s_{launches.csv,kernels.csv,summary.md}" % tag)
