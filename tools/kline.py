#!/usr/bin/env python3
"""one bench.py JSON line on stdin -> 'tag fps kernel ms ...' (used by the A/B loops run on the GPU box)"""
import json, sys
d = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "", "%.0f fps" % d["value"], " ".join("%s %.3f" % (k.replace("k_", ""), v["ms_per_step"]) for k, v in d["kernels"].items()))
