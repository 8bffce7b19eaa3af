#!/bin/bash
# One GPU-box visit: the whole GPU test-suite, the full default bench line (all sub-records) and the reference arm.
TAG=${1:-full}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$TAG.json'))
    print('main %s: ms/step %.3f value %.3e frac %.4f e2e ms %.2f cpu %.3e' % (d['config']['workload'][:3], d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('cpu_baseline',{}).get('value',0)))
    for k,v in d.get('workloads',{}).items():
        print('  %-10s ms %.3f value %.3e frac %.4f e2e ms %.2f e2e/cpu %s' % (k, v['ms_per_step'], v['value'], v.get('roofline_frac',0), v['e2e']['ms_per_step'], v.get('e2e_over_cpu')))
    print('  api', json.dumps(d.get('e2e_api'))[:600])
except Exception as e: print('no bench json', e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
