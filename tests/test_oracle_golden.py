"""The CPU oracle against the committed golden vectors.

Every fixture was produced by FFmpeg's independent ALAC encoder and verified at generation time to
decode to the source PCM with FFmpeg's decoder; here the oracle must reproduce the recorded sha256
of that source PCM -- the same bit-for-bit check as tests/conformance_test.go:282-292.
"""
import hashlib

import numpy as np
import pytest

import oracle_lib as ol
from golden_io import load_fixtures
from signals import make_signal

FIX = load_fixtures()


@pytest.mark.parametrize('name', sorted(FIX))
def test_oracle_matches_ffmpeg_golden(name):
    fx = FIX[name]
    m = fx['meta']
    st, cfg = ol.parse_cookie(fx['cookie'])
    assert st == ol.OK
    assert (cfg.bit_depth, cfg.num_channels, cfg.sample_rate) == (m['bits'], m['channels'], m['sample_rate'])
    pcm = b''
    for p in fx['packets']:
        st, out = ol.decode_packet(cfg, p)
        assert st == ol.OK
        pcm += out
    assert len(pcm) == m['frames'] * m['channels'] * cfg.bps()
    assert hashlib.sha256(pcm).hexdigest() == m['pcm_sha256']


@pytest.mark.parametrize('name', ['s16_stereo_44k', 's24_71_48k', 's16_stereo_silence_lsb'])
def test_golden_source_is_reproducible(name):
    """The recorded hash really is the seeded source signal (format-metadata check, conformance_test.go:267-279)."""
    m = FIX[name]['meta']
    x = make_signal(m['kind'], m['channels'], m['frames'], m['bits'], m['sample_rate'], seed=m['seed'])
    assert hashlib.sha256(ol.int_to_pcm_bytes(x, m['bits'])).hexdigest() == m['pcm_sha256']


def test_batch_driver_equals_single(fixtures=None):
    fx = FIX['s24_stereo_96k']
    st, cfg = ol.parse_cookie(fx['cookie'])
    packed, offs, sizes = ol.pack(fx['packets'])
    out, nb, status = ol.decode_batch(cfg, packed, offs, sizes, nthreads=3)
    assert (status == 0).all()
    for i, p in enumerate(fx['packets']):
        st, pcm = ol.decode_packet(cfg, p)
        assert bytes(out[i, :nb[i]]) == pcm


def test_encoder_roundtrip_all_depths():
    """test-side encoder -> oracle == source for the depths FFmpeg cannot make (restatement only)."""
    for bits, shift in ((20, 0), (32, 2), (24, 0), (24, 2), (16, 0)):
        for ch in (1, 2, 6):
            cfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000)
            x = make_signal('silence_lsb', ch, 5000, bits, 48000, seed=bits + ch)
            pk = ol.encode_stream(cfg, x, ol.PacketOpts.make(bytes_shifted=shift))
            got = []
            for p in pk:
                st, pcm = ol.decode_packet(cfg, p)
                assert st == ol.OK
                got.append(ol.pcm_bytes_to_int(pcm, bits, ch))
            y = np.concatenate(got)
            if bits == 20:
                y = y >> 4
            assert np.array_equal(y, x), (bits, shift, ch)


def test_synthetic_cases_are_well_formed_for_the_oracle():
    """The synthetic suites the GPU parity tests use: exotic cases decode (or hit a reference panic), hostile packets end in one of the
    reference's sentinels (or where it would panic), and the threaded batch driver equals packet-by-packet decoding."""
    import synth_cases
    n_exotic = n_hostile = 0
    for name, cfg, packets in synth_cases.exotic_cases():
        for p in packets:
            st, pcm = ol.decode_packet(cfg, p)
            assert (pcm is None) == (st != ol.OK), name
            # the only exotic shapes that do not decode are the ones on which the reference panics (appendix B7)
            assert st == ol.OK or ol.code(st) == 9, (name, st)
            n_exotic += st == ol.OK
    seen = set()
    for name, cfg, packets in synth_cases.hostile_cases(max_per_seed=12):
        packed, offs, sizes = ol.pack(packets)
        out, nb, status = ol.decode_batch(cfg, packed, offs, sizes, nthreads=3)
        for i, p in enumerate(packets):
            n_hostile += 1
            st, pcm = ol.decode_packet(cfg, p)
            assert st == status[i], name
            assert (pcm is None) == (st != ol.OK), name
            if pcm is not None:
                assert bytes(out[i, :nb[i]]) == pcm, name
            seen.add(ol.code(st))
    assert n_exotic > 200 and n_hostile > 300
    assert seen <= {0, 3, 4, 5, 6, 7, 9}, seen
