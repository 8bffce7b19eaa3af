"""Parity of the CUDA path (through the C ABI / Go-named host API) against the CPU oracle.

Bit-exact: every packet's status word and every PCM byte must equal the oracle's. Integer path, so
the tolerance is zero.
"""
import hashlib

import numpy as np
import pytest

import oracle_lib as ol
import synth_cases
from golden_io import load_fixtures
from signals import make_signal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pkg():
    from alac_b200_loader import load_package
    p = load_package()
    assert p.lib.alacb200_device_count() >= 1, 'no CUDA device: the product path has no CPU fallback'
    return p


def to_pkg_cfg(pkg, ocfg):
    return pkg.ParseMagicCookie(ol.make_cookie(ocfg))


def assert_parity(pkg, ocfg, packets, name=''):
    dec = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
    try:
        packed, offs, sizes = pkg.pack_packets(packets)
        out, nb, st = dec.decode_packed(packed, offs, sizes)
        want_out, want_nb, want_st = ol.decode_batch(ocfg, packed, offs, sizes, nthreads=4)
        bad = np.nonzero(st != want_st)[0]
        assert len(bad) == 0, f'{name}: status differs at {bad[:5]}: gpu {[hex(int(s)) for s in st[bad[:5]]]} oracle {[hex(int(s)) for s in want_st[bad[:5]]]}'
        assert np.array_equal(nb, want_nb), f'{name}: out_bytes differ'
        for i in range(len(packets)):
            if st[i] == 0 and not np.array_equal(out[i, :nb[i]], want_out[i, :nb[i]]):
                d = np.nonzero(out[i, :nb[i]] != want_out[i, :nb[i]])[0]
                raise AssertionError(f'{name}: packet {i} PCM differs at byte {d[0]} ({len(d)} bytes differ)')
        return out, nb, st
    finally:
        dec.close()


def test_golden_fixtures(pkg):
    """FFmpeg-encoded golden vectors: GPU == oracle == recorded sha256 of the source PCM."""
    for name, fx in load_fixtures().items():
        st, ocfg = ol.parse_cookie(fx['cookie'])
        assert st == 0
        dec = pkg.NewPacketDecoder(pkg.ParseMagicCookie(fx['cookie']), 0)
        pcm, errs = dec.DecodePackets(fx['packets'])
        dec.close()
        assert all(e is None for e in errs), (name, errs)
        assert hashlib.sha256(b''.join(pcm)).hexdigest() == fx['meta']['pcm_sha256'], name
        for p, g in zip(fx['packets'], pcm):
            st, want = ol.decode_packet(ocfg, p)
            assert g == want, name


def test_format_and_single_packet_api(pkg):
    fx = load_fixtures()['s24_stereo_96k']
    cfg = pkg.ParseMagicCookie(fx['cookie'])
    dec = pkg.NewPacketDecoder(cfg, 0)
    assert dec.Format() == pkg.PCMFormat(96000, 24, 2)
    _, ocfg = ol.parse_cookie(fx['cookie'])
    out = dec.DecodePacket(fx['packets'][-1])  # partial last packet -> shorter slice (decoder.go:127)
    assert len(out) == (9000 - 2 * 4096) * 2 * 3
    assert out == ol.decode_packet(ocfg, fx['packets'][-1])[1]
    with pytest.raises(pkg.ErrDecode) as ei:
        dec.DecodePacket(b'')
    assert 'bitstream overrun' in str(ei.value)
    dec.close()


def test_exotic_shapes(pkg):
    """20/32-bit, every order and mode, shift 0/1/2, sweeps, DSE/FIL/tag 3, partial+escape, odd element orders."""
    for name, ocfg, packets in synth_cases.exotic_cases():
        assert_parity(pkg, ocfg, packets, name)


def test_hostile_packets(pkg):
    """Truncated / bit-flipped / garbage packets: same status word as the oracle (incl. where the Go
    reference would panic), same PCM whenever the packet still decodes."""
    seen = set()
    for name, ocfg, packets in synth_cases.hostile_cases():
        _, _, st = assert_parity(pkg, ocfg, packets, name)
        seen |= {int(s) & 0xff for s in st}
    assert {0, 3, 4, 6, 7, 9} <= seen


def test_large_mixed_batch(pkg):
    """A few thousand packets in one call (several pipeline chunks, ragged tail group)."""
    ocfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
    x = make_signal('bench', 2, 4096 * 40 + 777, 24, 96000, seed=5)
    base = ol.encode_stream(ocfg, x)
    packets = (base * 60)[:2311]
    out, nb, st = assert_parity(pkg, ocfg, packets, 'large')
    assert (st == 0).all()
    # round trip against the source signal for the first pass
    got = b''.join(bytes(out[i, :nb[i]]) for i in range(len(base)))
    assert np.array_equal(ol.pcm_bytes_to_int(got, 24, 2), x)


def test_unaligned_offsets_and_wide_stride(pkg):
    """Packets at arbitrary byte offsets (the kernel aligns its 128-bit loads itself) and out_stride > frame bytes."""
    ocfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
    x = make_signal('silence_lsb', 2, 4096 * 5, 16, 44100, seed=9)
    packets = ol.encode_stream(ocfg, x)
    blob = bytearray(b'\xee' * 3)
    offs, sizes = [], []
    for i, p in enumerate(packets):
        offs.append(len(blob))
        sizes.append(len(p))
        blob += p + b'\xdd' * (i % 7)
    blob += b'\0' * 64
    dec = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
    stride = dec.frame_bytes + 36
    out, nb, st = dec.decode_packed(np.frombuffer(bytes(blob), dtype=np.uint8), np.array(offs, dtype=np.uint64),
                                    np.array(sizes, dtype=np.uint32), out_stride=stride)
    dec.close()
    assert (st == 0).all()
    got = b''.join(bytes(out[i, :nb[i]]) for i in range(len(packets)))
    assert np.array_equal(ol.pcm_bytes_to_int(got, 16, 2), x)
    assert (out[:, dec.frame_bytes:] == 0).all()
