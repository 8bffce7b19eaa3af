"""bench.py's reference arm on the CPU (no GPU needed): the JSON line keeps the driver's contract, other ranks stay quiet."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, ALAC_B200_CACHE=os.path.join('/tmp', 'alac_b200_cache_test'), **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'smoke', '--steps', '1',
                           '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'decoded PCM samples/s' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['dtype'] == 'int32' and d['data'] == 'synthetic'
    assert d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1 and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['config']['workload'].startswith('smoke')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_do_nothing():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
