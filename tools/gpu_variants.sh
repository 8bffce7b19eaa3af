#!/bin/bash
# One GPU-box visit comparing experiment builds (libalacb200_<name>.so, `make variant`): parity tests on the default build,
# then device-resident + e2e bench lines per (variant, workload). Usage: bash tools/gpu_variants.sh TAG "v1 v2" "c2 c3"
TAG=${1:-v}; VARIANTS=${2:-default}; WORKLOADS=${3:-c2}; PYTEST=${4:-yes}
mkdir -p gpurun_out
if [ "$PYTEST" = yes ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
for v in $VARIANTS; do
  if [ "$v" = default ]; then unset ALACB200_LIB; else export ALACB200_LIB=$PWD/saprobe-alac_b200/libalacb200_$v.so; fi
  for w in $WORKLOADS; do
    timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --only-main > gpurun_out/bench_${TAG}_${v}_$w.json 2> gpurun_out/bench_${TAG}_${v}_$w.err
    echo "== $v $w rc=$?"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}_${v}_$w.json'))
    print('   ms/step %.3f  value %.3e  frac %.4f  e2e ms %.2f  clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['clocks'].get('sm_mhz')))
except Exception as e: print('   no bench json', e)
PY
  done
done
