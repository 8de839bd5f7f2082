#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench, ncu launch list, full ncu captures of the top kernels.
# usage: tools/gpu_round.sh <tag> [kernel-regex for the full capture] [launches to capture] [skip-tests]
TAG=${1:-run}
KRE=${2:-wgrad_line_umma|conv_line_umma}
CNT=${3:-14}
mkdir -p gpurun_out
if [ -z "$4" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_${TAG}.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-e2e --no-infer > /dev/null 2> gpurun_out/ncu_launch_${TAG}.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KRE}" -c ${CNT} -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-e2e --no-infer > /dev/null 2> gpurun_out/ncu_full_${TAG}.err
[ -f gpurun_out/pytest_${TAG}.log ] && tail -2 gpurun_out/pytest_${TAG}.log
cat gpurun_out/bench_${TAG}.json | cut -c1-1800
