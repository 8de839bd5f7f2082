#!/bin/bash
# Round profile on the GPU box (under gpurun): full bench line, ncu launch list, ncu --set full of the top kernels.
# usage: tools/gpu_profile.sh <tag> [kernel regex] [count]
TAG=${1:-r2}
KRE=${2:-conv_line_umma|wgrad_line_umma}
CNT=${3:-18}
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 4 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --settle-steps 0 --no-cpu-baseline --no-profile --no-e2e --no-infer --no-gpu-reference > /dev/null 2> gpurun_out/ncu_launch_${TAG}.err
python tools/ncu_summary.py launches gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.txt 2>&1
head -30 gpurun_out/launches_${TAG}.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KRE}" -c ${CNT} -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --settle-steps 0 --no-cpu-baseline --no-profile --no-e2e --no-infer --no-gpu-reference > /dev/null 2> gpurun_out/ncu_full_${TAG}.err
python tools/ncu_summary.py full gpurun_out/prof_${TAG}.ncu-rep > gpurun_out/ncu_full_${TAG}.txt 2>&1
cat gpurun_out/ncu_full_${TAG}.txt
rm -f gpurun_out/prof_${TAG}.ncu-rep   # the summaries are what gets committed; the report would overflow the 64 MiB return path
python tools/show_bench.py gpurun_out/bench_${TAG}.json | head -70
