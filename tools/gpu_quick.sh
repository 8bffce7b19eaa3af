#!/bin/bash
# Quick GPU visit: parity tests + bench (no profiler). Usage: bash tools/gpu_quick.sh [tag] [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$TAG.json'))
    print('value %.3e samples/s  ms/step %.3f  xRT %.0f' % (d['value'], d['ms_per_step'], d['x_realtime']))
    print('kernels', d['kernels_ms_per_step'], 'roofline frac %.4f' % d['roofline']['frac'])
    print('e2e %.3e  ms %.2f' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'cpu', d.get('cpu_baseline',{}).get('value'), 'cores', d['host_cores'], d['clocks'])
except Exception as e: print('no bench json', e)
PY
