"""BASELINE configs[4] (library batch of mixed 16/24-bit tracks), configs[3] at full size and configs[0] through
NewDecoder/Read, plus the host-path features they rest on: tracks read in place from file images, bounds of the sample
table against the image, decoder-owned pinned arenas. Bit-exact against the CPU oracle / the source PCM."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol
from m4a_writer import build_m4a
from signals import make_signal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pkg():
    from alac_b200_loader import load_package
    p = load_package()
    assert p.lib.alacb200_device_count() >= 1, 'no CUDA device: the product path has no CPU fallback'
    return p


def _library(pkg, ntracks=72, seed=500):
    """-> (tracks for DecodeTracks, [(ocfg, packets, source or None)]): distinct tracks, 16- and 24-bit stereo alternating
    (the configs[4] mix), a few other shapes in between, some as whole M4A images with their sample tables."""
    tracks, ref = [], []
    rng = np.random.default_rng(seed)
    for t in range(ntracks):
        if t % 12 == 5:
            bits, ch, rate = 20, 2, 48000
        elif t % 12 == 9:
            bits, ch, rate = 24, 6, 48000
        elif t % 12 == 11:
            bits, ch, rate = 16, 1, 22050
        else:
            bits, ch, rate = (16, 2, 44100) if t % 2 == 0 else (24, 2, 96000)
        ocfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=rate)
        frames = 4096 * int(rng.integers(3, 40)) + int(rng.integers(0, 4096))
        x = make_signal(['bench', 'silence_lsb', 'music', 'loud'][t % 4], ch, frames, bits, rate, seed=seed + t)
        packets = ol.encode_stream(ocfg, x)
        cookie = ol.make_cookie(ocfg, wrappers=t % 3)
        if t % 4 == 1:  # a whole file image, read in place
            data, samples = build_m4a(ol.make_cookie(ocfg, wrappers=1), packets, channels=ch, bits=bits, rate=rate,
                                      samples_per_chunk=1 + t % 5, last_frames=frames % 4096)
            fc, fs = pkg.FindALACTrack(data)
            assert fs == samples
            tracks.append(pkg.Track(fc, np.frombuffer(data, dtype=np.uint8), np.array([o for o, _ in fs], dtype=np.uint64),
                                    np.array([z for _, z in fs], dtype=np.uint32)))
        else:
            packed, offs, sizes = pkg.pack_packets(packets)
            tracks.append(pkg.Track(cookie, packed, offs, sizes))
        ref.append((ocfg, packets, x))
    return tracks, ref


def _check_track(res, ocfg, packets, x, name):
    assert res.err is None, (name, res.err)
    packed, offs, sizes = ol.pack(packets)
    want, wnb, wst = ol.decode_batch(ocfg, packed, offs, sizes, nthreads=4)
    assert np.array_equal(res.status, wst) and np.array_equal(res.out_bytes, wnb), name
    for i in range(len(packets)):
        assert np.array_equal(res.pcm[i, :wnb[i]], want[i, :wnb[i]]), (name, i)
    if x is not None and ocfg.bit_depth != 20:
        assert np.array_equal(ol.pcm_bytes_to_int(res.pcm_bytes(), ocfg.bit_depth, ocfg.num_channels), x), name


def test_library_mixed_depth_tracks(pkg):
    """72 distinct tracks, 16- and 24-bit mixed (plus 20-bit, 5.1 and mono ones), packed buffers and whole M4A images, in ONE
    DecodeTracks call: every track equals the oracle packet for packet and (lossless) its source; a second call on the
    same library reuses everything."""
    tracks, ref = _library(pkg)
    lib = pkg.NewLibraryDecoder((0,))
    try:
        for rnd in range(2):
            results = lib.DecodeTracks(tracks)
            assert len(results) == len(tracks)
            for t, (res, (ocfg, packets, x)) in enumerate(zip(results, ref)):
                assert res.device == 0
                assert (res.config.BitDepth, res.config.NumChannels, res.config.SampleRate) == (ocfg.bit_depth, ocfg.num_channels, ocfg.sample_rate)
                _check_track(res, ocfg, packets, x if rnd == 0 else None, f'track {t} round {rnd}')
    finally:
        lib.close()


def test_library_bad_tracks_do_not_disturb_good_ones(pkg):
    """A bad cookie / unsupported depth gives that track NewPacketDecoder's ErrConfig; a sample table that points outside
    the image gives those packets a READ error (decode.go:172-174), not a decode error; an empty track is fine; the
    other tracks of the call decode as usual."""
    tracks, ref = _library(pkg, ntracks=6, seed=900)
    good = ol.make_cookie(ref[0][0])
    bad_depth = bytearray(ol.make_cookie(ref[0][0], wrappers=0))
    bad_depth[5] = 12
    tracks.insert(1, pkg.Track(b'\x00' * 10, tracks[0].data, tracks[0].offsets, tracks[0].sizes))
    tracks.insert(3, pkg.Track(bytes(bad_depth), tracks[0].data, tracks[0].offsets, tracks[0].sizes))
    tracks.append(pkg.Track(good, tracks[0].data, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint32)))
    offs = tracks[0].offsets.copy()
    sizes = tracks[0].sizes.copy()
    offs[1] = len(tracks[0].data) + 5          # starts outside
    sizes[2] = len(tracks[0].data)             # runs outside
    tracks.append(pkg.Track(good, tracks[0].data, offs, sizes))
    lib = pkg.NewLibraryDecoder((0,))
    try:
        res = lib.DecodeTracks(tracks)
    finally:
        lib.close()
    assert isinstance(res[1].err, pkg.ErrConfig) and 'invalid magic cookie' in str(res[1].err)
    assert isinstance(res[3].err, pkg.ErrConfig) and 'unsupported bit depth: 12' in str(res[3].err)
    assert res[-2].err is None and len(res[-2].status) == 0
    k = 0
    for t, r in enumerate(res[:-2]):
        if t in (1, 3):
            continue
        _check_track(r, *ref[k], f'track {t}')
        k += 1
    last = res[-1]
    assert last.err is None
    assert last.status[1] == pkg.ST_IO_TRUNCATED and last.status[2] == pkg.ST_IO_TRUNCATED and last.out_bytes[1] == 0
    assert (np.delete(last.status, [1, 2]) == 0).all()
    ocfg, packets, _ = ref[0]
    for i in (0, 3, len(packets) - 1):
        assert last.packet(i) == ol.decode_packet(ocfg, packets[i])[1]


def test_library_over_all_devices(pkg):
    """The same library sharded by contiguous track ranges over every device of the box (one submitting thread each, no
    collective): same bytes, and every device gets work."""
    ndev = pkg.lib.alacb200_device_count()
    if ndev < 2:
        pytest.skip('one device on this box')
    tracks, ref = _library(pkg, ntracks=40, seed=700)
    lib = pkg.NewLibraryDecoder(tuple(range(ndev)))
    try:
        results = lib.DecodeTracks(tracks)
    finally:
        lib.close()
    assert {r.device for r in results} == set(range(min(ndev, len(tracks))))
    devs = [r.device for r in results]
    assert devs == sorted(devs)  # contiguous ranges
    for t, (res, (ocfg, packets, x)) in enumerate(zip(results, ref)):
        _check_track(res, ocfg, packets, x, f'track {t}')


def test_arena_is_reused_and_in_place_bounds(pkg):
    """DecodePackets packs into the decoder's own pinned arena (no pinned allocation per call once it has grown), and the
    raw call checks every (offset, size) against the bytes handed over."""
    ocfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
    x = make_signal('music', 2, 4096 * 9, 16, 44100, seed=3)
    packets = ol.encode_stream(ocfg, x)
    dec = pkg.NewPacketDecoder(pkg.ParseMagicCookie(ol.make_cookie(ocfg)), 0)
    try:
        a0, b0 = dec._arena(1 << 20, 1 << 20)
        ptrs = (a0.ctypes.data, b0.ctypes.data)
        for _ in range(3):
            pcm, errs = dec.DecodePackets(packets)
            assert all(e is None for e in errs)
            assert np.array_equal(ol.pcm_bytes_to_int(b''.join(pcm), 16, 2), x)
            a, b = dec._arena(1 << 20, 1 << 20)
            assert (a.ctypes.data, b.ctypes.data) == ptrs
        assert dec.DecodePacket(packets[4]) == ol.decode_packet(ocfg, packets[4])[1]
        packed, offs, sizes = pkg.pack_packets(packets)
        offs2 = offs.copy()
        offs2[3] = packed.nbytes - 10
        out, nb, st = dec.decode_packed(packed, offs2, sizes)
        assert st[3] == pkg.ST_IO_TRUNCATED and nb[3] == 0 and (np.delete(st, 3) == 0).all()
        assert 'unexpected EOF' in pkg.format_error(pkg.ST_IO_TRUNCATED)
    finally:
        dec.close()


def test_c1_full_size_through_read_and_seek(pkg):
    """BASELINE configs[0] at full size: 16-bit stereo 44.1 kHz, 60 s (646 packets, the last one short) in an M4A, decoded
    through NewDecoder/Read in io.Reader-sized pieces: bit for bit the source (conformance_test.go:282-292), then Seek to
    0 / 25 / 50 / 75 % and the tail of the stream (conformance_test.go:343-421). The file image is read in place."""
    import bench
    wl = bench.build_workload('c1', seed=1, threads=bench.host_cores())
    n = len(wl['sizes'])
    assert n == 646
    packets = [bytes(wl['packed'][int(o):int(o) + int(z)]) for o, z in zip(wl['offsets'], wl['sizes'])]
    data, _ = build_m4a(wl['cookie'], packets, samples_per_chunk=9, last_frames=wl['frames'] % 4096)
    want, wnb, wst = ol.decode_batch(wl['cfg'], wl['packed'], wl['offsets'], wl['sizes'], nthreads=bench.host_cores())
    assert (wst == 0).all()
    want = b''.join(bytes(want[i, :wnb[i]]) for i in range(n))
    assert len(want) == wl['frames'] * 4
    dec = pkg.NewDecoder(data)
    try:
        assert dec.Format() == pkg.PCMFormat(44100, 16, 2)
        got = bytearray()
        while True:
            b = dec.Read(32 * 1024)  # io.Copy's buffer size
            if not b:
                break
            got += b
        assert bytes(got) == want
        for frac in (0.0, 0.25, 0.5, 0.75):
            at = dec.Seek(int(dec.Duration() * frac))
            assert at == dec.Position()
            frame = dec.sampleIdx * 4096
            assert hashlib.sha256(dec.ReadAll()).hexdigest() == hashlib.sha256(want[frame * 4:]).hexdigest()
    finally:
        dec.close()
    # a truncated file (moov first, the tail of mdat missing): the packets that are still whole decode, the first missing
    # one is a READ error with the reference's text (decode.go:172-174), not a decode error
    data2, samples2 = build_m4a(wl['cookie'], packets, moov_first=True, last_frames=wl['frames'] % 4096)
    cut_len = len(data2) - 200000
    first_bad = next(i for i, (o, z) in enumerate(samples2) if o + z > cut_len)
    cut = pkg.NewDecoder(data2[:cut_len])
    try:
        got = bytearray()
        with pytest.raises(IOError) as ei:
            while True:
                b = cut.Read(1 << 20)
                if not b:
                    break
                got += b
        assert f'reading sample {first_bad}: unexpected EOF' in str(ei.value)
        assert bytes(got) == want[:first_bad * 4096 * 4]
    finally:
        cut.close()


def test_full_size_c4_matches_oracle(pkg):
    """BASELINE configs[3] at full size: 7.1 (SCE + 3 x CPE + LFE) 24-bit 48 kHz, 30 min, 21 094 packets, 2.07 GB of PCM.
    Every packet OK, byte count, and the digest of the whole PCM stream equals the oracle's."""
    import bench
    wl = bench.build_workload('c4', seed=4, threads=bench.host_cores())
    n = len(wl['sizes'])
    assert n == 21094
    dec = pkg.NewPacketDecoder(pkg.ParseMagicCookie(wl['cookie']))
    fb = wl['cfg'].frame_bytes()
    out, nb, st = dec.decode_packed(wl['packed'], wl['offsets'], wl['sizes'], out=np.empty((n, fb), dtype=np.uint8), out_stride=fb)
    dec.close()
    assert (st == 0).all()
    assert int(nb.astype(np.int64).sum()) == wl['frames'] * 8 * 3
    want, wnb, wst = ol.decode_batch(wl['cfg'], wl['packed'], wl['offsets'], wl['sizes'], nthreads=bench.host_cores(),
                                     out=np.empty((n, fb), dtype=np.uint8))
    assert (wst == 0).all() and np.array_equal(nb, wnb)
    assert hashlib.sha256(out[:n - 1].tobytes()).hexdigest() == hashlib.sha256(want[:n - 1].tobytes()).hexdigest()
    assert np.array_equal(out[n - 1, :nb[n - 1]], want[n - 1, :nb[n - 1]])
