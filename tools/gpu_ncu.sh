#!/bin/bash
# ncu visit: per workload a launch list (gpu__time_duration) and ONE full capture of alac_decode_kernel.
# Usage (under gpurun): bash tools/gpu_ncu.sh TAG "c2 lib16 c4"
TAG=$1; shift
mkdir -p gpurun_out
for w in $1; do
  CMD="python bench.py --only-main --workload $w --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_${TAG}_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_$w.csv $CMD > gpurun_out/ncu_launches_${TAG}_$w.log 2>&1
  echo "ncu launches $w rc=$?"
  $CMD > gpurun_out/plain2_${TAG}_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:alac_decode -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$w $CMD > gpurun_out/ncu_full_${TAG}_$w.log 2>&1
  echo "ncu full $w rc=$?"; tail -2 gpurun_out/ncu_full_${TAG}_$w.log
done
