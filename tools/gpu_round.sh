#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench, ncu launch list, one full ncu capture of the top kernels.
# usage: tools/gpu_round.sh <tag> [kernel-regex for the full capture]
TAG=${1:-run}
KRE=${2:-conv_taps_umma}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_launch_${TAG}.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 40 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_full_${TAG}.err
tail -2 gpurun_out/pytest_${TAG}.log
cat gpurun_out/bench_${TAG}.json | cut -c1-2500
