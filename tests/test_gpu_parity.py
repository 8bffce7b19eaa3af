"""Parity of the CUDA path (through the C ABI / Go-named host API) against the CPU oracle.

Bit-exact: every packet's status word and every PCM byte must equal the oracle's. Integer path, so
the tolerance is zero.
"""
import hashlib
import os

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


def test_frame_length_sweep(pkg):
    """Frame lengths at and around the kernel's own granularities (16 / 32 samples, 8 / 16 frames), 1 and 65536."""
    for name, ocfg, packets in synth_cases.frame_length_cases():
        _, _, st = assert_parity(pkg, ocfg, packets, name)
        assert (st == 0).all(), name


def test_entropy_edge_sweep(pkg):
    """Cookie parameters at their edges, silence / impulses / long zero runs at every internal boundary, +-LSB and
    clipping noise: the entropy stage's batches, freezes and flavour switches against the oracle."""
    for name, ocfg, packets in synth_cases.entropy_edge_cases():
        assert_parity(pkg, ocfg, packets, name)


def test_wide_ffmpeg_matrix_gpu(pkg):
    """The reference's conformance matrix ({16, 24} x 11 rates x 1-8 channels, conformance_test.go:573-628), FFmpeg-encoded
    on this box: the CUDA path must return the source PCM bit for bit (FFmpeg's own decoder is checked in build())."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import wide_matrix
    if not wide_matrix.ffmpeg_available():
        pytest.skip('FFmpeg libraries not importable on this box; the committed ffmpeg_fixtures.npz remains the pin')
    npk = 0
    for case in wide_matrix.wide_cases():
        cookie, packets, x = wide_matrix.build(case)
        cfg = pkg.ParseMagicCookie(cookie)
        dec = pkg.NewPacketDecoder(cfg, 0)
        try:
            assert dec.Format() == pkg.PCMFormat(case['rate'], case['bits'], case['channels'])
            packed, offs, sizes = pkg.pack_packets(packets)
            out, nb, st = dec.decode_packed(packed, offs, sizes)
        finally:
            dec.close()
        assert (st == 0).all(), case['name']
        got = b''.join(bytes(out[i, :nb[i]]) for i in range(len(packets)))
        assert got == ol.int_to_pcm_bytes(x, case['bits']), case['name']
        npk += len(packets)
    assert npk > 3000


def test_exotic_ffmpeg_matrix_gpu(pkg):
    """20- and 32-bit, 0 / 1 / 2 shifted bytes, the order-31 pre-pass, orders 0-31 (tests/golden/exotic_matrix.py; the
    packets come from the test-side encoder and FFmpeg's decoder has returned the source for them in build()): the CUDA
    path must return the source PCM bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import exotic_matrix
    if not exotic_matrix.ffmpeg_available():
        pytest.skip('FFmpeg libraries not importable on this box')
    confirmed = 0
    decs = {}
    try:
        for case in exotic_matrix.exotic_cases():
            built = exotic_matrix.build(case)
            if built is None:
                continue
            ocfg, packets, x = built
            key = (case['bits'], case['channels'], ocfg.frame_length)
            if key not in decs:
                decs[key] = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
            packed, offs, sizes = pkg.pack_packets(packets)
            out, nb, st = decs[key].decode_packed(packed, offs, sizes)
            assert (st == 0).all(), case['name']
            got = b''.join(bytes(out[i, :nb[i]]) for i in range(len(packets)))
            assert got == ol.int_to_pcm_bytes(x, case['bits']), case['name']
            confirmed += 1
    finally:
        for d in decs.values():
            d.close()
    assert confirmed >= 300


@pytest.mark.parametrize('build', ['by batch size', 'throughput build only'])
def test_synth_hashes_pin_the_gpu(pkg, build, monkeypatch):
    """The committed drift pin (tests/golden/synth_hashes.json: status word, byte count and PCM of every synthetic case
    as recorded from the cross-checked oracle) against the CUDA path directly -- not via today's oracle build.
    The library carries two builds of the one kernel (register budgets of six and of eight CTAs per SM, chosen by batch
    size); small batches normally run the first, so the second pass forces them through the other one."""
    import synth_pin
    if build == 'throughput build only':
        monkeypatch.setenv('ALACB200_NO_LAT_BUILD', '1')
    pin = synth_pin.load_pin()['cases']
    other_input, bad, n = [], [], 0
    decs = {}
    try:
        for name, ocfg, packets in synth_pin.all_cases():
            if synth_pin.packets_digest(ocfg, packets) != pin[name]['packets_sha256']:
                other_input.append(name)
                continue
            key = bytes(ol.make_cookie(ocfg))
            if key not in decs:
                decs[key] = pkg.NewPacketDecoder(pkg.ParseMagicCookie(key), 0)
            packed, offs, sizes = pkg.pack_packets(packets)
            out, nb, st = decs[key].decode_packed(packed, offs, sizes)
            n += 1
            if synth_pin.result_digest(st, nb, out) != pin[name]['result_sha256']:
                bad.append(name)
    finally:
        for d in decs.values():
            d.close()
    assert not bad, f'GPU output differs from the committed pin on {len(bad)} cases: {bad[:5]}'
    assert n > 400 and len(other_input) <= len(pin) // 20, (n, other_input[:5])


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


def test_strides_and_offsets_for_every_emit_path(pkg):
    """out_stride = frame bytes + 0 / 4 / 8 / 12 / 20 (so packet slots are 16-, 8- or only 4-byte aligned) and packets at odd
    byte offsets, for the shapes that take the live path (stereo), the direct tail (mono), the row tail (6 / 8 channels)
    and the tile tail (3 channels, under-filled packets): the 128-bit store paths must fall back to narrower stores and
    never put PCM outside a packet's frame bytes."""
    for ch, bits, shift in ((2, 24, 1), (2, 16, 0), (1, 24, 1), (6, 16, 0), (8, 24, 1), (3, 24, 0), (2, 20, 0), (2, 32, 2)):
        ocfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000)
        x = make_signal('silence_lsb', ch, 4096 * 3 + 999, bits, 48000, seed=60 + ch + bits)
        packets = ol.encode_stream(ocfg, x, ol.PacketOpts.make(bytes_shifted=shift))
        packets.append(ol.Writer(ocfg).element(0, x[:4096, 0].copy(), order=4, coefs=[60, -30, 10, 5]).end().bytes())  # under-filled
        blob = bytearray(b'\xee' * 5)
        offs, sizes = [], []
        for i, p in enumerate(packets):
            offs.append(len(blob))
            sizes.append(len(p))
            blob += p + b'\xdd' * (1 + i % 5)
        blob += b'\0' * 64
        packed = np.frombuffer(bytes(blob), dtype=np.uint8)
        offs = np.array(offs, dtype=np.uint64)
        sizes = np.array(sizes, dtype=np.uint32)
        want, want_nb, want_st = ol.decode_batch(ocfg, packed, offs, sizes, nthreads=2)
        dec = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
        try:
            for extra in (0, 4, 8, 12, 20):
                stride = (dec.frame_bytes + 3) // 4 * 4 + extra
                canary = np.full((len(packets), stride), 0xA5, dtype=np.uint8)
                out, nb, st = dec.decode_packed(packed, offs, sizes, out=canary, out_stride=stride)
                assert np.array_equal(st, want_st) and np.array_equal(nb, want_nb), (ch, bits, extra)
                for i in range(len(packets)):
                    assert np.array_equal(out[i, :nb[i]], want[i, :nb[i]]), (ch, bits, extra, i)
                    assert (out[i, nb[i]:dec.frame_bytes] == 0).all(), (ch, bits, extra, i)  # rest of the frame: zeros
                    gap = out[i, dec.frame_bytes:]  # the gap carries no PCM: zeros (copied with the slot) or untouched
                    assert ((gap == 0) | (gap == 0xA5)).all(), (ch, bits, extra, i, gap[:8].tolist())
        finally:
            dec.close()


def test_packed_row_tail_shapes(pkg):
    """The packed row tail (emit_rows_packed: canonical multi-channel packets of full-length elements, 16-bit, or 24-bit with
    0 / 1 shifted byte) for every channel count that has elements to interleave, with packets at odd byte offsets (the
    aligned shift-byte windows then start at every phase), 40 full packets (a whole group and a part of one), slots
    16-byte aligned and not (the latter must fall back to the byte-granular row path), and a frame length of 1024."""
    for ch, bits, shift, fl in ((3, 24, 1, 4096), (4, 24, 1, 4096), (5, 16, 0, 4096), (5, 24, 1, 4096), (6, 24, 0, 4096),
                                (6, 24, 1, 1024), (7, 24, 1, 4096), (8, 16, 0, 4096), (8, 24, 1, 4096), (2, 24, 1, 4096)):
        ocfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000, frame_length=fl)
        x = make_signal('silence_lsb', ch, fl * 40, bits, 48000, seed=900 + ch + bits + shift)
        packets = ol.encode_stream(ocfg, x, ol.PacketOpts.make(bytes_shifted=shift))
        if ch == 2:  # two mono elements in a 2-channel stream: canonical, not a pair
            packets = [ol.Writer(ocfg).element(0, x[i * fl:(i + 1) * fl, 0].copy(), order=4, coefs=[60, -30, 10, 5], bytes_shifted=shift)
                       .element(0, x[i * fl:(i + 1) * fl, 1].copy(), order=5, coefs=[50, -20, 10, 5, 1], bytes_shifted=shift, instance=1)
                       .end().bytes() for i in range(40)]
        blob = bytearray(b'\xee' * 3)
        offs, sizes = [], []
        for i, p in enumerate(packets):
            offs.append(len(blob))
            sizes.append(len(p))
            blob += p + b'\xdd' * (i % 7)
        blob += b'\0' * 64
        packed = np.frombuffer(bytes(blob), dtype=np.uint8)
        offs = np.array(offs, dtype=np.uint64)
        sizes = np.array(sizes, dtype=np.uint32)
        want, want_nb, want_st = ol.decode_batch(ocfg, packed, offs, sizes, nthreads=4)
        assert (want_st == 0).all() and (want_nb == ocfg.frame_bytes()).all()
        dec = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
        try:
            for extra in (0, 16, 4):
                stride = (dec.frame_bytes + 15) // 16 * 16 + extra
                canary = np.full((len(packets), stride), 0xA5, dtype=np.uint8)
                out, nb, st = dec.decode_packed(packed, offs, sizes, out=canary, out_stride=stride)
                assert np.array_equal(st, want_st) and np.array_equal(nb, want_nb), (ch, bits, shift, extra)
                assert np.array_equal(out[:, :dec.frame_bytes], want), (ch, bits, shift, extra)
                gap = out[:, dec.frame_bytes:]
                assert ((gap == 0) | (gap == 0xA5)).all(), (ch, bits, shift, extra)
        finally:
            dec.close()


def test_decoder_read_seek_m4a(pkg):
    """NewDecoder/Read/Seek over an M4A (BASELINE configs[0] shape, shortened): conformance_test.go:282-292 (bit-for-bit
    vs source) and :343-421 (seek at 0/25/50/75 % equals the tail of the full decode)."""
    from m4a_writer import build_m4a
    ocfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
    frames = 44100 * 6 + 321
    x = make_signal('bench', 2, frames, 16, 44100, seed=1)
    packets = ol.encode_stream(ocfg, x)
    data, _ = build_m4a(ol.make_cookie(ocfg), packets, samples_per_chunk=7, last_frames=frames % 4096)
    want = ol.int_to_pcm_bytes(x, 16)
    dec = pkg.NewDecoder(data, window=16)
    assert dec.Format() == pkg.PCMFormat(44100, 16, 2)
    assert dec.Duration() == len(packets) * 4096 * 10**9 // 44100  # over-counts the partial last packet (decode.go:82-88)
    got = b''
    while True:
        b = dec.Read(70001)
        if not b:
            break
        got += b
    assert got == want
    assert dec.Read(10) == b''  # io.EOF stays EOF
    for frac in (0.0, 0.25, 0.5, 0.75):
        at = dec.Seek(int(dec.Duration() * frac))
        assert at == dec.Position()
        frame = dec.sampleIdx * 4096  # packet-aligned (decode.go:109-121); `at` is that frame in truncated nanoseconds
        assert at == frame * 10**9 // 44100
        assert dec.ReadAll() == want[frame * 4:]
    assert dec.Seek(10 * dec.Duration()) == dec.Duration() and dec.Read(1) == b''
    assert dec.Seek(-5) == 0


def test_container_error_paths_on_gpu(pkg):
    """error_test.go:368-442: garbage mdat -> success or ErrDecode, truncated packet -> any outcome but no crash."""
    from m4a_writer import build_m4a
    ocfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
    x = make_signal('music', 2, 4096 * 3, 16, 44100, seed=4)
    packets = ol.encode_stream(ocfg, x)
    data, samples = build_m4a(ol.make_cookie(ocfg), packets)
    off, size = samples[1]
    garbage = bytearray(data)
    garbage[off:off + size] = np.random.default_rng(1).integers(0, 256, size, dtype=np.uint8).tobytes()
    dec = pkg.NewDecoder(bytes(garbage))
    try:
        out = dec.ReadAll()
        assert len(out) <= len(x) * 4
    except pkg.ErrDecode as e:
        assert 'decoding packet 1' in str(e)
    # same statuses as the oracle for the damaged packet
    st_o, _ = ol.decode_packet(ocfg, bytes(garbage[off:off + size]))
    d = pkg.NewPacketDecoder(pkg.ParseMagicCookie(ol.make_cookie(ocfg)))
    _, errs = d.DecodePackets([bytes(garbage[off:off + size])])
    assert (errs[0].status if errs[0] else 0) == st_o
    truncated = pkg.NewDecoder(data[:off + size // 2] + b'\0' * 16 + data[off + size // 2 + 16:])
    try:
        truncated.ReadAll()
    except pkg.AlacError:
        pass


def test_full_size_c2_matches_oracle(pkg):
    """BASELINE configs[1] at full size (14 063 packets, 345.6 MB of PCM): GPU PCM == oracle PCM, digest of the whole
    stream, every packet OK, sizes add up."""
    import bench
    wl = bench.build_workload('c2', seed=2, threads=bench.host_cores())
    n = len(wl['sizes'])
    assert n == 14063
    dec = pkg.NewPacketDecoder(pkg.ParseMagicCookie(wl['cookie']))
    out, nb, st = dec.decode_packed(wl['packed'], wl['offsets'], wl['sizes'])
    dec.close()
    assert (st == 0).all()
    assert int(nb.sum()) == wl['frames'] * 2 * 3
    want, wnb, wst = ol.decode_batch(wl['cfg'], wl['packed'], wl['offsets'], wl['sizes'], nthreads=bench.host_cores())
    assert (wst == 0).all() and np.array_equal(nb, wnb)
    fb = wl['cfg'].frame_bytes()
    assert np.array_equal(out[:, :fb], want)
    h = hashlib.sha256()
    for i in range(n):
        h.update(out[i, :nb[i]].tobytes())
    h2 = hashlib.sha256()
    for i in range(n):
        h2.update(want[i, :wnb[i]].tobytes())
    assert h.hexdigest() == h2.hexdigest()


def test_mixed_group_live_and_fallback(pkg):
    """2-channel batch whose 32-packet groups mix pairs the emit warp writes live (orders 4-6/8) with packets that must fall
    back (general orders, escape pairs, two SCEs, under-filled, mode != 0, partial frames, broken packets): every group
    shape -- all live, none live, a lane failing mid-stream next to live lanes -- must give the oracle's bytes."""
    ocfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
    rng = np.random.default_rng(31)
    x = make_signal('silence_lsb', 2, 4096 * 12, 24, 96000, seed=12)
    kinds = [ol.PacketOpts.make(), ol.PacketOpts.make(min_order=8, max_order=8), ol.PacketOpts.make(min_order=12, max_order=12),
             ol.PacketOpts.make(force_escape=1), ol.PacketOpts.make(mode=2), ol.PacketOpts.make(bytes_shifted=0),
             ol.PacketOpts.make(bytes_shifted=2, min_order=31, max_order=31), ol.PacketOpts.make(mix_bits=2, mix_res=-1)]
    variants = [ol.encode_stream(ocfg, x, o) for o in kinds]
    a = x[:4096, 0]
    two_sce = ol.Writer(ocfg).element(0, a, order=4, coefs=[60, -30, 10, 5], bytes_shifted=1).element(0, a[::-1].copy(), order=6, coefs=[90, -45, 22, -11, 5, -2]).end().bytes()
    under = ol.Writer(ocfg).element(0, a, order=5, coefs=[60, -30, 10, 5, 1]).end().bytes()
    partial = ol.encode_packet(ocfg, x[:1234])
    def batch(pick):
        out = []
        for i in range(32 * 9 + 5):
            k = pick(i)
            if k == 'two': out.append(two_sce)
            elif k == 'under': out.append(under)
            elif k == 'partial': out.append(partial)
            elif k == 'broken':
                p = bytearray(variants[0][i % 12]); p[len(p) // 2] ^= 0x5a; p = p[:len(p) - 7]; out.append(bytes(p))
            else: out.append(variants[k][i % 12])
        return out
    # groups: 0 all live; 1 all order 8; 2 random mix of everything; 3 live + one broken lane; 4 live + partial; ...
    def pick(i):
        g, l = divmod(i, 32)
        if g == 0: return 0
        if g == 1: return 1
        if g == 3: return 'broken' if l == 17 else 0
        if g == 4: return 'partial' if l % 5 == 0 else 7
        if g == 5: return 3
        if g == 6: return 'two' if l % 2 else 'under'
        return [0, 1, 2, 3, 4, 5, 6, 7, 'two', 'under', 'partial', 'broken'][int(rng.integers(0, 12))]
    packets = batch(pick)
    _, _, st = assert_parity(pkg, ocfg, packets, 'mixed groups')
    assert (st != 0).any() and (st == 0).sum() > 200


def test_mixed_groups_multichannel_tail(pkg):
    """Multi-channel batches whose 32-packet groups mix canonical packets (row-path tail) with packets that force the
    transpose-tile tail (under-filled, pairs first / spilling, mixed per-element counts, END-only, broken): whole groups
    of one kind, and groups where a single odd packet sits among canonical ones."""
    for ch, bits in ((6, 16), (8, 24), (3, 24), (5, 20)):
        ocfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000)
        x = make_signal('silence_lsb', ch, 4096 * 6, bits, 48000, seed=40 + ch)
        canon = ol.encode_stream(ocfg, x)
        shifted = ol.encode_stream(ocfg, x, ol.PacketOpts.make(bytes_shifted=1)) if bits >= 24 else canon
        a = x[:4096, 0].copy()
        b = (x[:4096, 1] // 2).copy()
        under = ol.Writer(ocfg).element(0, a, order=4, coefs=[60, -30, 10, 5]).end().bytes()
        w = ol.Writer(ocfg)
        idx = 0
        while idx + 2 <= ch:
            w.element(1, a, b, order=4, coefs=[80, -40, 20, -10])
            idx += 2
        if idx < ch:
            w.element(0, b, order=6, coefs=[90, -45, 22, -11, 5, -2])
        pairs_first = w.end().bytes()
        w = ol.Writer(ocfg)
        for k in range(ch):
            nk = [4096, 100, 4096, 7, 2000, 1, 4096, 333][k]
            w.element(0, a[:nk], order=[4, 0, 31, 2, 5, 6, 8, 9][k], coefs=[40, -20, 10, -5, 2, -1, 1, 0, 0], partial=1)
        mixed_counts = w.end().bytes()
        end_only = ol.Writer(ocfg).end().bytes()
        partial = ol.encode_packet(ocfg, x[:1500])
        broken = bytes(bytearray(canon[0])[:len(canon[0]) // 2])
        odd = [under, pairs_first, mixed_counts, end_only, partial, broken]
        packets = []
        for i in range(32 * 6 + 11):
            g, l = divmod(i, 32)
            if g == 0: p = canon[i % 6]                      # all canonical
            elif g == 1: p = shifted[i % 6]                  # all canonical, shift buffer
            elif g == 2: p = odd[l % 6]                      # none canonical
            elif g == 3: p = odd[(l // 5) % 6] if l % 5 == 0 else canon[i % 6]
            elif g == 4: p = end_only if l == 9 else partial  # same short count + an empty packet: still canonical
            else: p = canon[i % 6] if l != 21 else odd[g % 6]
            packets.append(p)
        _, _, st = assert_parity(pkg, ocfg, packets, f'multichannel mixed c{ch} d{bits}')
        assert (st == 0).sum() > 150


def test_full_size_c3_shift_buffer_stream(pkg):
    """BASELINE configs[2] at full size: 1 h of 24-bit stereo 192 kHz, 168 750 packets, every element header carries
    bytesShifted=1. Size-independent checks: every packet OK, byte count, and the digest of the whole PCM stream equals
    the oracle's (decoded on all host cores)."""
    import bench
    wl = bench.build_workload('c3', seed=3, threads=bench.host_cores())
    n = len(wl['sizes'])
    assert n == 168750
    # element header: tag(3) instance(4) unused(12) partial(1) shift(2) escape(1): shift = bits 20-21 of the packet
    first = wl['packed'][wl['offsets'][::997].astype(np.int64) + 2]
    assert (((first >> 2) & 3) == 1).all()
    dec = pkg.NewPacketDecoder(pkg.ParseMagicCookie(wl['cookie']))
    fb = wl['cfg'].frame_bytes()
    out, nb, st = dec.decode_packed(wl['packed'], wl['offsets'], wl['sizes'], out=np.empty((n, fb), dtype=np.uint8), out_stride=fb)
    dec.close()
    assert (st == 0).all()
    assert int(nb.astype(np.int64).sum()) == wl['frames'] * 2 * 3
    want, wnb, wst = ol.decode_batch(wl['cfg'], wl['packed'], wl['offsets'], wl['sizes'], nthreads=bench.host_cores(),
                                     out=np.empty((n, fb), dtype=np.uint8))
    assert (wst == 0).all() and np.array_equal(nb, wnb)
    assert hashlib.sha256(out[:n - 1].tobytes()).hexdigest() == hashlib.sha256(want[:n - 1].tobytes()).hexdigest()
    assert np.array_equal(out[n - 1, :nb[n - 1]], want[n - 1, :nb[n - 1]])


def test_cli_twin_wav_and_pcm(pkg, tmp_path):
    """tools/alac_decode.py mirrors cmd/alac-example-decoder (example_decoder_test.go:35-131): PCM output equals the
    source, WAV output is a 44-byte header + the same PCM."""
    import subprocess
    import sys
    from m4a_writer import build_m4a
    ocfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
    x = make_signal('music', 2, 4096 * 4 + 99, 16, 44100, seed=8)
    data, _ = build_m4a(ol.make_cookie(ocfg), ol.encode_stream(ocfg, x), last_frames=99)
    f = tmp_path / 't.m4a'
    f.write_bytes(data)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, 'tools', 'alac_decode.py')
    pcm = subprocess.run([sys.executable, cli, '-format', 'pcm', str(f)], capture_output=True, check=True).stdout
    want = ol.int_to_pcm_bytes(x, 16)
    assert pcm == want
    wav = subprocess.run([sys.executable, cli, '-'], input=data, capture_output=True, check=True)
    assert wav.stdout[:4] == b'RIFF' and wav.stdout[8:16] == b'WAVEfmt ' and wav.stdout[44:] == want
    assert b'44100 Hz, 16-bit, 2 channel' in wav.stderr


def test_huge_batch_runs_as_several_launches(pkg):
    """More packet groups than the scratch cap allows in one launch (8 GB / (channels * frame_length * 128 B)): the device
    path splits into stream-ordered launches reusing the scratch; results must not depend on where the cuts fall."""
    import torch
    ocfg = ol.Config.make(bit_depth=16, num_channels=8, sample_rate=48000, frame_length=65536)
    x = make_signal('music', 8, 65536 * 2, 16, 48000, seed=21)
    base = ol.encode_stream(ocfg, x)
    n = 32 * 128 * 2 + 45  # cap = 8 GB / (8 * 65536 * 128 B) = 128 groups -> 3 launches
    packets = [base[i % 2] for i in range(n)]
    dec = pkg.NewPacketDecoder(to_pkg_cfg(pkg, ocfg), 0)
    packed, offs, sizes = pkg.pack_packets(packets)
    stride = dec.frame_bytes
    d_packed = torch.from_numpy(packed).cuda()
    d_off = torch.from_numpy(offs.view(np.int64)).cuda()
    d_sz = torch.from_numpy(sizes.view(np.int32)).cuda()
    d_pcm = torch.empty(n * stride, dtype=torch.uint8, device='cuda')
    d_nb = torch.zeros(n, dtype=torch.int32, device='cuda')
    d_st = torch.full((n,), -1, dtype=torch.int32, device='cuda')
    rc = pkg.lib.alacb200_decode_packets_device(dec._h, d_packed.data_ptr(), d_packed.numel(), d_off.data_ptr(), d_sz.data_ptr(), n,
                                                d_pcm.data_ptr(), stride, d_nb.data_ptr(), d_st.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int((d_st != 0).sum()) == 0 and int((d_nb != stride).sum()) == 0
    want = [torch.frombuffer(bytearray(ol.decode_packet(ocfg, base[k])[1]), dtype=torch.uint8).cuda() for k in range(2)]
    rows = d_pcm.view(n, stride)
    for i in list(range(0, n, 211)) + [n - 1, 4095, 4096, 4097, 8191, 8192]:
        assert torch.equal(rows[i], want[i % 2]), i
    dec.close()
