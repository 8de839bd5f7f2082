#!/bin/bash
# A/B of environment switches on the training step (same box, back to back): tools/ab_env.sh "VAR=a VAR2=b" "VAR=c" ...
# prints ms/step of `bench.py` (value pass only) per setting, two rounds so that drift shows.
mkdir -p gpurun_out
for round in 1 2; do
  for setting in "$@"; do
    out=$(env $setting timeout 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-infer --no-e2e --no-profile 2>gpurun_out/ab_env.err | tail -1)
    ms=$(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms/step  %s MHz' % (d['ms_per_step'], d['clocks']['sm_mhz']))" 2>/dev/null || echo "FAILED: $(tail -2 gpurun_out/ab_env.err)")
    echo "round $round  [$setting]  $ms"
  done
done
