#!/bin/bash
# Same-box A/B of one environment switch.  usage: tools/ab.sh <tag> <ENV_NAME> <valA> <valB> [bench args...]
TAG=$1; VAR=$2; A=$3; B=$4; shift 4
mkdir -p gpurun_out
for v in $A $B $A $B; do
  env $VAR=$v timeout 600 python bench.py --steps 8 --warmup 4 --no-gpu-reference "$@" > gpurun_out/ab_${TAG}_${v}.json 2> gpurun_out/ab_${TAG}_${v}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_${TAG}_${v}.json"))
r = d["roofline"]
print("$VAR=$v ms_per_step %.2f sm %s" % (d["ms_per_step"], d["clocks"].get("sm_mhz")), {k: round(x, 2) for k, x in r.items() if k.startswith("ms_")})
PY
done
