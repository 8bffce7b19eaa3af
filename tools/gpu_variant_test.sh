#!/bin/bash
# Parity tests + bench lines for ONE experiment build. Usage: bash tools/gpu_variant_test.sh TAG VARIANT "c2 lib" [pytest-k-expr]
TAG=$1; V=$2; WORKLOADS=${3:-c2}; KEXPR=${4:-}
mkdir -p gpurun_out
if [ "$V" != default ]; then export ALACB200_LIB=$PWD/saprobe-alac_b200/libalacb200_$V.so; fi
export ALACB200_VERBOSE=1
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
fi
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log
for w in $WORKLOADS; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --only-main > gpurun_out/bench_${TAG}_${V}_$w.json 2> gpurun_out/bench_${TAG}_${V}_$w.err
  echo "== $V $w rc=$?"; grep -h "alacb200:" gpurun_out/bench_${TAG}_${V}_$w.err | head -1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}_${V}_$w.json'))
    print('   ms/step %.3f  value %.3e  frac %.4f  e2e ms %.2f  clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['clocks'].get('sm_mhz')))
except Exception as e:
    print('   no bench json', e)
    import subprocess; print(subprocess.run(['tail','-5','gpurun_out/bench_${TAG}_${V}_$w.err'],capture_output=True,text=True).stdout)
PY
done
