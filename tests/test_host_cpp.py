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


def _mutated_images(count=6000, seed=7):
    from test_host_abi import _packets
    rng = np.random.default_rng(seed)
    cookie = ol.make_cookie(ol.Config.make(), wrappers=1)
    bases = [build_m4a(cookie, _packets(), **kw)[0] for kw in (dict(), dict(samples_per_chunk=3, co64=True), dict(qt_v1=True, moov_first=True),
                                                                 dict(extra_trak=True), dict(mdat_large=True), dict(constant_stsz=False, samples_per_chunk=1))]
    for it in range(count):
        d = bytearray(bases[it % len(bases)])
        kind = it % 5
        if kind == 0:
            for _ in range(int(rng.integers(1, 8))):
                d[int(rng.integers(0, len(d)))] ^= int(rng.integers(1, 256))
        elif kind == 1:
            i = int(rng.integers(0, len(d) - 4))
            d[i:i + 4] = [b'\xff\xff\xff\xff', b'\x00\x00\x00\x00', b'\x7f\xff\xff\xff', b'\x00\x00\x00\x01', b'\x80\x00\x00\x00'][int(rng.integers(0, 5))]
        elif kind == 2:
            d = d[:int(rng.integers(0, len(d)))]
        elif kind == 3:
            a, b = sorted(int(x) for x in rng.integers(0, len(d), size=2))
            c = int(rng.integers(0, len(d)))
            piece = d[a:b][:64]
            d[c:c + len(piece)] = piece
        else:  # an 8-byte field (co64 offsets, 64-bit box sizes) gets an extreme value
            i = int(rng.integers(0, len(d) - 8))
            d[i:i + 8] = [b'\xff' * 8, b'\x7f' + b'\xff' * 7, b'\x00' * 7 + b'\x01'][int(rng.integers(0, 3))]
        yield bytes(d)


def test_mp4_parser_under_address_sanitizer(tmp_path):
    """The container parser compiled with AddressSanitizer + UBSan, fed 6000 mutated M4A images held in exact-size heap
    blocks: any read past an image, overflow or misuse aborts the driver."""
    probe = tmp_path / 'probe.cpp'
    probe.write_text('int main(){return 0;}')
    if subprocess.run(['g++', '-fsanitize=address,undefined', '-o', str(tmp_path / 'probe'), str(probe)], capture_output=True).returncode != 0:
        pytest.skip('no sanitizer runtime in this image')
    exe = str(tmp_path / 'mp4_fuzz')
    subprocess.run(['g++', '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-o', exe,
                    os.path.join(ROOT, 'tests', 'cpp', 'mp4_fuzz_driver.cpp'), os.path.join(ROOT, 'saprobe-alac_b200', 'host', 'mp4.cpp')], check=True)
    blob = tmp_path / 'images.bin'
    with open(blob, 'wb') as f:
        for img in _mutated_images():
            f.write(len(img).to_bytes(4, 'little'))
            f.write(img)
    r = subprocess.run([exe, str(blob)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert 'fuzz ok=' in r.stdout
    ok, bad = (int(x.split('=')[1]) for x in r.stdout.split()[1:3])
    assert ok > 400 and bad > 400, r.stdout
