for c in 12 24 48 96; do
  ALACB200_CHUNKS=$c python bench.py --only-main --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c e2e ms %.2f' % d['e2e']['ms_per_step'])"
done
