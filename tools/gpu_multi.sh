#!/bin/bash
# Multi-GPU visit (run with gpurun --gpus N): library test over all devices, concurrent PCIe floor, bench at N ranks.
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_$TAG.txt; nproc >> gpurun_out/topo_$TAG.txt; free -g | head -2 >> gpurun_out/topo_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_library.py -m gpu -x -q -k "over_all_devices" > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_$TAG.log
H=$((2672 / N)); D=$((3955 / N))  # MiB of c3 per rank: compressed in, PCM out
timeout 300 $TR tools/pcie_probe.py $H $D 2>&1 | grep PCIE_FLOOR | tee gpurun_out/pcie_$TAG.txt
timeout 300 python tools/pcie_probe.py $H $D 2>&1 | grep PCIE_FLOOR | tee -a gpurun_out/pcie_$TAG.txt
timeout 1500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}_${N}gpu.json'))
    print('N=%d c3 strong: ms/step %.3f value %.3e e2e ms %.2f e2e value %.3e' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value']))
    l=d['lib_weak']; print('   lib weak: ms %.3f value %.3e e2e ms %.2f e2e value %.3e' % (l['ms_per_step'], l['value'], l['e2e']['ms_per_step'], l['e2e']['value']))
except Exception as e: print('no bench json', e)
PY
