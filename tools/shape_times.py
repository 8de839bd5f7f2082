#!/usr/bin/env python
"""Print the strided / transposed rows (or rows matching argv[2]) of a bench.py JSON line's top_shapes."""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
pat = sys.argv[2:] or ["(2, 2, 2)", "(1, 2, 2)"]
print("%.3f ms/step %s MHz" % (d["ms_per_step"], d["clocks"]["sm_mhz"]))
for k in d["top_shapes"]:
    if any(p in k["shape"] for p in pat):
        print("%-14s %-62s %7.4f ms %7.1f" % (k["kernel"], k["shape"], k["ms_per_step"], k["tflops"] or 0))
