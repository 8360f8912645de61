#!/usr/bin/env python3
"""Static SASS instruction counts per kernel of the built liborbx.so (cuobjdump -sass): the TMA / mbarrier / packed-integer mnemonics
that show what the kernels are made of.  usage: tools/sass_evidence.py <tag>   -> profiles/<tag>_sass_evidence.md"""
import collections, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = subprocess.run(["cuobjdump", "-sass", os.path.join(root, "dynamic-visual-slam_b200", "lib", "liborbx.so")], capture_output=True, text=True).stdout
cols = ["UTMALDG", "UTMAPF", "SYNCS", "UTCIMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "IMMA", "VABSDIFF4", "VIMNMX3", "IDP.4A", "IDP.2A", "PRMT", "POPC", "ATOMS", "REDUX"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in cols:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
        counts[cur]["_total"] += 1
with open(os.path.join(root, "profiles", tag + "_sass_evidence.md"), "w") as f:
    f.write("# SASS evidence (cuobjdump -sass of liborbx.so, sm_100a)\n\nStatic instruction counts per kernel: `UTMALDG` / `UTMAPF` = TMA tensor load / L2 prefetch "
            "(cp.async.bulk.tensor / cp.async.bulk.prefetch.tensor), `SYNCS` = mbarrier operations, `UTCIMMA` = tcgen05.mma kind::i8, `LDTM` = tcgen05.ld (TMEM -> registers), "
            "`UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM allocation, `IMMA` = mma.sync int8, `VABSDIFF4` / `VIMNMX3` / `IDP` / `PRMT` / `POPC` = "
            "the packed-integer instructions the kernels are built on, `ATOMS` = shared-memory atomics, `REDUX` = warp reductions.\n\n")
    f.write("| kernel | SASS instr | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for k, c in counts.items():
        name = re.sub(r"\(.*", "", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()).replace("void ", "")
        if name.startswith("k_"):
            f.write("| `%s` | %d | %s |\n" % (name, c["_total"], " | ".join(str(c[w]) for w in cols)))
print("wrote profiles/%s_sass_evidence.md" % tag)
