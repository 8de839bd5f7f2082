"""Condensed view of a bench.py JSON line: headline, roofline scalars, per-CUDA-kernel and per-shape tables."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.2f patches/s  %.3f ms/step  e2e %s  steps %s" % (d["value"], d["ms_per_step"], d["e2e"].get("value"),
                                                              d.get("ms_each_step")))
r = d.get("roofline") or {}
for k, v in r.items():
    if not k.startswith("ms_"):
        print("  roofline.%-36s %s" % (k, v))
print("  e2e:", {k: v for k, v in d["e2e"].items()})
print("  clocks:", d.get("clocks"))
print("  cpu_baseline:", d.get("cpu_baseline"))
print("-- CUDA kernels (launches/step, ms/step, TFLOP/s)")
for k, v in (d.get("cuda_kernels") or {}).items():
    print("  %-24s %5.1f %8.3f %s" % (k, v["launches"] / d["steps"], v["ms_per_step"],
                                      "%.0f" % v["tflops"] if v["tflops"] else "-"))
print("-- entry points")
for k, v in d["kernels"].items():
    print("  %-24s %5.1f %8.3f %s" % (k, v["launches"] / d["steps"], v["ms_per_step"],
                                      "%.0f" % v["tflops"] if v["tflops"] else "-"))
if len(sys.argv) > 2:
    print("-- shapes")
    for s in d.get("top_shapes", []):
        print("  %-22s %-62s %3d %8.4f %s" % (s["kernel"], s["shape"], s["launches"] // d["steps"], s["ms_per_step"],
                                              s["tflops"]))
