"""The reference's conformance matrix, reproduced with the FFmpeg ALAC encoder + decoder bundled in this image.

/root/reference/tests/conformance_test.go:573-628 runs {16, 24 bit} x 11 sample rates x 1-8 channels: encode with
FFmpeg, decode with every decoder, compare bit for bit with the source and between decoders. Here:

    wide_cases()   {16, 24} x 1-8 channels x 3 rates (8 k, 44.1 k, 192 k) with 64 full packets + a partial one each, and
                   the other 8 rates of the reference's list with 2 packets + a partial one each; signals alternate
                   between music, music with silence / +-2 LSB passages (zero-run and k == 1 branches) and loud clipping
    build(case)    -> (cookie, packets, source [frames, ch]) after asserting FFmpeg-decode(packets) == source

At ~65 packets per case the packets would be ~50 MB, too much to commit, so the matrix is REGENERATED from its seeds by
the tests that use it (tests/test_oracle_golden.py on the CPU, tests/test_gpu_parity.py on the GPU box -- the same image,
so the same FFmpeg build) and every decoder is compared with the source PCM itself. When the FFmpeg libraries are not
importable those tests skip, and the committed subset tests/golden/ffmpeg_fixtures.npz remains the pin.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from signals import make_signal  # noqa: E402

RATES_ALL = (8000, 11025, 16000, 22050, 32000, 44100, 48000, 88200, 96000, 176400, 192000)  # conformance_test.go:573-575
RATES_LONG = (8000, 44100, 192000)
KINDS = ('silence_lsb', 'music', 'loud')


def wide_cases():
    """-> list of dict(name, bits, channels, rate, frames, kind, seed)"""
    out = []
    for bits in (16, 24):
        for ch in range(1, 9):
            for ri, rate in enumerate(RATES_ALL):
                long_case = rate in RATES_LONG
                frames = 64 * 4096 + 1000 + 17 * ch if long_case else 2 * 4096 + 500 + ch
                kind = KINDS[(RATES_LONG.index(rate) + ch) % 3] if long_case else KINDS[(ri + ch + bits) % 3]
                out.append(dict(name=f'w{bits}_c{ch}_{rate}', bits=bits, channels=ch, rate=rate, frames=frames, kind=kind,
                                seed=bits * 1000 + ch * 37 + ri))
    return out


def ffmpeg_available():
    try:
        sys.path.insert(0, HERE)
        import ffmpeg_alac  # noqa: F401
        return True
    except Exception:
        return False


def build(case):
    """-> (cookie, [packets], source int64 [frames, ch]); asserts the FFmpeg decoder returns the source."""
    sys.path.insert(0, HERE)
    import ffmpeg_alac as ff
    x = make_signal(case['kind'], case['channels'], case['frames'], case['bits'], case['rate'], seed=case['seed'])
    cookie, packets = ff.alac_encode(np.ascontiguousarray(x.T), case['bits'], case['rate'], {})
    y = ff.alac_decode(cookie, packets, case['bits'], case['channels'], case['rate']).T
    assert np.array_equal(y, x), f"FFmpeg round trip failed for {case['name']}"
    return cookie, packets, x
