#!/bin/bash
# ncu visit: launch list + one full capture of the decode kernel. Usage: bash tools/gpu_ncu.sh tag [kernel-regex]
TAG=${1:-n}; KRE=${2:-alac_decode}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload ${WORKLOAD:-c2}"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 1 -f -o gpurun_out/prof_${KRE}_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
