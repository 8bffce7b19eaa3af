#!/bin/bash
# developer probe: role cycles on the dev build for the given workloads
mkdir -p gpurun_out
for w in "$@"; do
  echo "=== role cycles $w"; ALACB200_LIB=$PWD/saprobe-alac_b200/libalacb200_dev.so timeout 600 python tools/role_cycles.py $w 2>&1 | tail -30
done
