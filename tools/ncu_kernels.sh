#!/bin/bash
# usage: tools/ncu_kernels.sh <tag> <kernel regex> <count> [bench args]   -- full ncu capture of selected kernels of one bench step
TAG=$1; KRE=$2; CNT=$3; shift 3
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KRE}" -c ${CNT} -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-e2e --no-infer --no-gpu-reference "$@" > /dev/null 2> gpurun_out/ncu_${TAG}.err
tail -3 gpurun_out/ncu_${TAG}.err
python tools/ncu_summary.py full gpurun_out/prof_${TAG}.ncu-rep 2>/dev/null | head -80
