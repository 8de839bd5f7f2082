#!/bin/bash
# as ab_env.sh with 30 timed steps and four alternating rounds (separates a 0.2 ms effect from the power-cap noise)
mkdir -p gpurun_out
for round in 1 2 3 4; do
  for setting in "$@"; do
    out=$(env $setting timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-infer --no-e2e --no-profile 2>gpurun_out/ab_env.err | tail -1)
    echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); import statistics as st; print('round $round [$setting] %.3f ms/step (median of steps %.3f) %s MHz' % (d['ms_per_step'], st.median(d['ms_each_step']), d['clocks']['sm_mhz']))" 2>/dev/null || echo "FAILED"
  done
done
