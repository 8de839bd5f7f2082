#!/bin/bash
# Run on the GPU box (under gpurun).  usage: tools/gpu_run.sh <tag> [tests|notests] [bench args...]
TAG=${1:-run}; shift
TESTS=${1:-tests}; shift
mkdir -p gpurun_out
if [ "$TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log
  tail -15 gpurun_out/pytest_${TAG}.log
fi
timeout 1500 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"
tail -5 gpurun_out/bench_${TAG}.err
python tools/show_bench.py gpurun_out/bench_${TAG}.json 2>&1 | head -80
