"""CPU-side checks of the product library: it loads, exports every symbol include/alac_b200.h declares,
its host logic (cookie, error text, MP4 sample tables) matches the reference's behaviour, and it fails
loudly -- never falls back -- when there is no CUDA device. No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
from golden_io import load_fixtures
from m4a_writer import build_m4a

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def pkg():
    from alac_b200_loader import load_package
    return load_package()


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, 'include', 'alac_b200.h')).read()
    declared = set(re.findall(r'\b(alacb200_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 22
    lib = C.CDLL(pkg.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(pkg.ABI_SYMBOLS) == declared


def test_product_does_not_touch_the_oracle():
    """Parity claims are void if the product path routes through oracle/: check the sources."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'saprobe-alac_b200')):
        for f in files:
            if f.endswith(('.cu', '.cuh', '.cpp', '.h', '.py', '.go')):
                src = open(os.path.join(dirpath, f), errors='ignore').read()
                assert 'oracle_lib' not in src and 'alac_oracle' not in src and 'libalac_oracle' not in src, f


def test_parse_magic_cookie_matches_oracle(pkg):
    for fx in load_fixtures().values():
        cfg = pkg.ParseMagicCookie(fx['cookie'])
        st, ocfg = ol.parse_cookie(fx['cookie'])
        assert st == 0
        assert (cfg.FrameLength, cfg.BitDepth, cfg.NumChannels, cfg.PB, cfg.MB, cfg.KB, cfg.MaxRun, cfg.MaxFrameBytes,
                cfg.AvgBitRate, cfg.SampleRate) == (ocfg.frame_length, ocfg.bit_depth, ocfg.num_channels, ocfg.pb,
                                                    ocfg.mb, ocfg.kb, ocfg.max_run, ocfg.max_frame_bytes,
                                                    ocfg.avg_bit_rate, ocfg.sample_rate)


def test_cookie_wrappers_and_errors(pkg):
    """error_test.go:81-122: too short / nil / bad version -> ErrConfig; frma+alac wrappers are stripped."""
    ocfg = ol.Config.make(bit_depth=24, num_channels=6, sample_rate=48000, max_frame_bytes=1234, avg_bit_rate=99)
    for wrappers in (0, 1, 2, 3):
        cfg = pkg.ParseMagicCookie(ol.make_cookie(ocfg, wrappers))
        assert (cfg.BitDepth, cfg.NumChannels, cfg.SampleRate, cfg.MaxFrameBytes) == (24, 6, 48000, 1234)
    for bad in (b'', None, b'\0' * 23, ol.make_cookie(ocfg, 1)[:30]):
        with pytest.raises(pkg.ErrConfig) as ei:
            pkg.ParseMagicCookie(bad)
        assert 'invalid configuration: alac: invalid magic cookie' in str(ei.value)
    ck = bytearray(ol.make_cookie(ocfg, 0))
    ck[4] = 1
    with pytest.raises(pkg.ErrConfig) as ei:
        pkg.ParseMagicCookie(bytes(ck))
    assert 'unsupported compatible version' in str(ei.value)


def test_bad_bit_depth_is_config_error_before_any_device_use(pkg):
    """error_test.go:126-142 / decoder.go:91-93."""
    cfg = pkg.ParseMagicCookie(ol.make_cookie(ol.Config.make(bit_depth=16)))
    cfg.BitDepth = 17
    with pytest.raises(pkg.ErrConfig) as ei:
        pkg.NewPacketDecoder(cfg)
    assert 'unsupported bit depth: 17' in str(ei.value)
    cfg.BitDepth, cfg.NumChannels = 16, 9
    with pytest.raises(pkg.ErrConfig):
        pkg.NewPacketDecoder(cfg)


def test_no_device_fails_loudly_no_cpu_fallback(pkg):
    if pkg.lib.alacb200_device_count() > 0:
        pytest.skip('a CUDA device is present')
    cfg = pkg.ParseMagicCookie(ol.make_cookie(ol.Config.make()))
    with pytest.raises(pkg.CudaError) as ei:
        pkg.NewPacketDecoder(cfg)
    assert 'no CPU fallback' in str(ei.value)


def test_library_decoder_without_device_fails_loudly(pkg):
    """The multi-track entry point has no CPU fallback either; its argument checks run without a device."""
    assert pkg.lib.alacb200_library_create(None, 0, None) == pkg.E_ARG
    if pkg.lib.alacb200_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(pkg.CudaError) as ei:
        pkg.NewLibraryDecoder((0,))
    assert 'no CPU fallback' in str(ei.value)


def test_error_text_matches_reference_wrapping(pkg):
    assert pkg.format_error(6) == 'decode failed: alac: bitstream overrun'                       # decoder.go:144
    assert pkg.format_error(6 | (2 << 8) | (2 << 12)) == 'decode failed: CPE: entropy decode U: alac: bitstream overrun'
    assert pkg.format_error(7 | (1 << 8) | (1 << 12)) == 'decode failed: SCE/LFE: entropy decode: alac: sample count exceeds buffer'
    assert pkg.format_error(4 | (2 << 8)) == 'decode failed: CPE: alac: invalid frame header'
    assert pkg.format_error(3) == 'decode failed: alac: unsupported element type (CCE/PCE)'
    assert pkg.format_error(6 | (4 << 8)) == 'decode failed: FIL: alac: bitstream overrun'
    assert pkg.format_error(8) == 'invalid configuration: alac: unsupported bit depth'
    assert isinstance(pkg.error_from_status(5 | (1 << 8)), pkg.ErrDecode)
    assert isinstance(pkg.error_from_status(1), pkg.ErrConfig)
    assert pkg.format_error(pkg.ST_IO_TRUNCATED) == 'unexpected EOF'  # io.ReadFull's error, decode.go:172-174


def _packets(n=11, size=None):
    rng = np.random.default_rng(3)
    return [bytes(rng.integers(0, 256, size=size or int(rng.integers(20, 400)), dtype=np.uint8)) for _ in range(n)]


@pytest.mark.parametrize('kw', [dict(), dict(samples_per_chunk=4), dict(samples_per_chunk=1, co64=True), dict(qt_v1=True),
                                dict(moov_first=True), dict(extra_trak=True, samples_per_chunk=3), dict(mdat_large=True),
                                dict(constant_stsz=True)])
def test_mp4_sample_table(pkg, kw):
    """buildSampleTable, mp4.go:382-420: stco|co64 x stsc x stsz (+ QT v1 entries, 64-bit boxes)."""
    pk = _packets(size=100) if kw.get('constant_stsz') else _packets()
    cookie = ol.make_cookie(ol.Config.make(), wrappers=1)
    data, want = build_m4a(cookie, pk, **kw)
    got_cookie, samples = pkg.FindALACTrack(data)
    assert got_cookie == cookie
    assert samples == want
    for (off, size), p in zip(samples, pk):
        assert data[off:off + size] == p


def test_mp4_error_paths(pkg):
    """error_test.go:146-366: empty / garbage / truncated containers -> ErrNoTrack; zeroed stsz count -> empty table."""
    cookie = ol.make_cookie(ol.Config.make(), wrappers=1)
    data, _ = build_m4a(cookie, _packets())
    for bad in (b'', b'\x12\x34' * 50, data[:20], data[:data.index(b'moov') + 20]):
        with pytest.raises(pkg.ErrNoTrack):
            pkg.FindALACTrack(bad)
    corrupt = data.replace(b'alac', b'xxxx')  # no ALAC sample entry any more
    with pytest.raises(pkg.ErrNoTrack) as ei:
        pkg.FindALACTrack(corrupt)
    assert 'no ALAC track' in str(ei.value)
    i = data.index(b'stsz')
    zeroed = data[:i + 12] + b'\0\0\0\0' + data[i + 16:]
    _, samples = pkg.FindALACTrack(zeroed)
    assert samples == []
    j = data.index(b'stco')
    with pytest.raises(pkg.ErrNoTrack):
        pkg.FindALACTrack(data[:j] + b'xtco' + data[j + 4:])


def test_mp4_parser_survives_mutations(pkg):
    """The container parser is host code that reads untrusted files: bit flips, overwritten size fields, truncations and
    splices of valid containers must end in a table or in ErrNoTrack, never in a crash, and a table it does return must
    be bounded by what the file declares (entries are validated again, against the file length, when they are read)."""
    rng = np.random.default_rng(2024)
    cookie = ol.make_cookie(ol.Config.make(), wrappers=1)
    bases = [build_m4a(cookie, _packets(), **kw)[0] for kw in (dict(), dict(samples_per_chunk=3, co64=True), dict(qt_v1=True, moov_first=True),
                                                                 dict(extra_trak=True), dict(mdat_large=True))]
    ok = bad = 0
    for it in range(600):
        d = bytearray(bases[it % len(bases)])
        kind = it % 4
        if kind == 0:  # random byte flips
            for _ in range(int(rng.integers(1, 8))):
                d[int(rng.integers(0, len(d)))] ^= int(rng.integers(1, 256))
        elif kind == 1:  # a 32-bit field (most of the file is box sizes / counts / offsets) gets an extreme value
            i = int(rng.integers(0, len(d) - 4))
            d[i:i + 4] = [b'\xff\xff\xff\xff', b'\x00\x00\x00\x00', b'\x7f\xff\xff\xff', b'\x00\x00\x00\x01'][int(rng.integers(0, 4))]
        elif kind == 2:  # truncation
            d = d[:int(rng.integers(0, len(d)))]
        else:  # splice a piece of the file over another place
            a, b2 = sorted(int(x) for x in rng.integers(0, len(d), size=2))
            c = int(rng.integers(0, len(d)))
            piece = d[a:b2][:64]
            d[c:c + len(piece)] = piece
        try:
            got_cookie, samples = pkg.FindALACTrack(bytes(d))
            ok += 1
            assert len(got_cookie) <= len(d)
            assert len(samples) <= 1 << 24
        except pkg.ErrNoTrack:
            bad += 1
    assert ok > 50 and bad > 50, (ok, bad)


def test_missing_library_fails_loudly(pkg, monkeypatch):
    """No silent fallback: without libalacb200.so the package refuses to import (and says how to build it)."""
    monkeypatch.setattr(pkg, 'LIB_PATH', os.path.join(ROOT, 'saprobe-alac_b200', 'does_not_exist.so'))
    with pytest.raises(ImportError) as ei:
        pkg._load()
    assert 'no CPU fallback' in str(ei.value)
