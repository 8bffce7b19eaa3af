"""The C++ host mirror of the Go API (saprobe-alac_b200/host/alac.hpp): compiles against the C ABI, behaves like the
reference on the CPU-only paths, and on a GPU decodes an M4A through NewDecoder/Read/Seek bit-exactly."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from m4a_writer import build_m4a
from signals import make_signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, 'tests', 'cpp', 'build', 'host_api_test')


def build_binary():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    lib_dir = os.path.join(ROOT, 'saprobe-alac_b200')
    src = os.path.join(ROOT, 'tests', 'cpp', 'host_api_test.cpp')
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib_dir, 'host', 'alac.hpp'))):
        subprocess.run(['g++', '-std=c++17', '-O1', '-Wall', '-o', BIN, src, '-L' + lib_dir, '-lalacb200', '-Wl,-rpath,' + lib_dir],
                       check=True)
    return BIN


def test_cpp_host_cpu_paths():
    exe = build_binary()
    r = subprocess.run([exe, 'cpu'], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert 'cpu ok' in r.stdout


@pytest.mark.gpu
def test_cpp_decoder_read_seek_on_gpu(tmp_path):
    exe = build_binary()
    cfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
    x = make_signal('silence_lsb', 2, 4096 * 23 + 1500, 24, 96000, seed=77)
    packets = ol.encode_stream(cfg, x)
    data, _ = build_m4a(ol.make_cookie(cfg), packets, channels=2, bits=24, rate=96000, samples_per_chunk=5, last_frames=1500)
    m4a, want = tmp_path / 'a.m4a', tmp_path / 'want.pcm'
    m4a.write_bytes(data)
    want.write_bytes(ol.int_to_pcm_bytes(x, 24))
    r = subprocess.run([exe, 'gpu', str(m4a), str(want)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'gpu ok' in r.stdout
