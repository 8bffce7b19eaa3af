"""Synthesised ALAC packets for shapes FFmpeg never emits (SURVEY.md section 4 "what FFmpeg fixtures do
not cover"): 20/32-bit, mode != 0, orders 0..31, pbFactor/denShift/mixRes sweeps, bytesShifted 0/1/2,
tag 3, DSE/FIL/END placement, escape + partial, tiny partial frames, odd frame lengths, cookie parameter
sweeps, non-canonical element orders, and hostile mutations.

Parity on these is "restatement only": the CUDA path is compared with the CPU oracle (status word and
PCM), the committed hashes in tests/golden/synth_hashes.json keep the oracle itself from drifting.
"""
import numpy as np

import oracle_lib as ol
from signals import make_signal


def _sig(ch, n, bits, seed, kind='silence_lsb'):
    return make_signal(kind, ch, n, bits, 48000, seed=seed)


def exotic_cases():
    """-> list of (name, ol.Config, [packet bytes])"""
    out = []
    seed = 1000

    def add(name, cfg, packets):
        out.append((name, cfg, list(packets)))

    # -- every depth x shift, mono / stereo / 5.1, LPC orders 4..6 --------------------------------
    for bits in (16, 20, 24, 32):
        for shift in (0, 1, 2):
            for ch in (1, 2, 6):
                seed += 1
                cfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000)
                x = _sig(ch, 4096 + 700, bits, seed)
                add(f'd{bits}_s{shift}_c{ch}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(bytes_shifted=shift)))
    # -- every predictor order, both coefficient-width paths, with and without the delta pre-pass --
    for order in list(range(0, 32)):
        for mode in (0, 3):
            seed += 1
            cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
            x = _sig(2, 2500, 16, seed, 'music')
            mn, mx = (0, 0) if order == 0 else (order, order)
            add(f'order{order}_mode{mode}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(min_order=mn, max_order=mx, mode=mode)))
    for order in (4, 6, 8, 13, 31):
        seed += 1
        cfg = ol.Config.make(bit_depth=24, num_channels=1, sample_rate=96000)
        x = _sig(1, 4096, 24, seed, 'music')
        add(f'd24_mono_order{order}_mode1', cfg,
            ol.encode_stream(cfg, x, ol.PacketOpts.make(min_order=order, max_order=order, mode=1, bytes_shifted=1)))
    # -- denShift / pbFactor / mix sweeps -------------------------------------------------------------
    for den in (0, 1, 4, 9, 12, 15):
        for pbf in (0, 1, 4, 7):
            seed += 1
            cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
            x = _sig(2, 3000, 16, seed, 'music')
            add(f'den{den}_pbf{pbf}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(den_shift=den, pb_factor=pbf)))
    for (mb_, mr) in ((0, 0), (1, 1), (2, 1), (2, 3), (5, -3), (8, 100), (8, -128), (31, 1), (32, 1), (40, -7), (255, 127), (0, 5)):
        seed += 1
        cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100)
        x = _sig(2, 3000, 16, seed, 'music') // 4
        add(f'mix{mb_}_{mr}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(mix_bits=mb_, mix_res=mr)))
        seed += 1
        cfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
        x = _sig(2, 3000, 24, seed, 'music') // 4
        add(f'd24_mix{mb_}_{mr}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(mix_bits=mb_, mix_res=mr)))
    # -- cookie parameter sweeps ----------------------------------------------------------------------
    for (pb, mb, kb) in ((40, 10, 14), (0, 10, 14), (255, 255, 31), (40, 0, 0), (40, 10, 1), (40, 10, 8), (20, 100, 20), (40, 10, 255)):
        seed += 1
        cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100, pb=pb, mb=mb, kb=kb)
        x = _sig(2, 3000, 16, seed)
        add(f'cookie_pb{pb}_mb{mb}_kb{kb}', cfg, ol.encode_stream(cfg, x))
    for fl in (1, 2, 7, 33, 256, 1024, 4095, 8192, 16384):
        seed += 1
        cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100, frame_length=fl)
        x = _sig(2, min(3 * fl + fl // 2, 20000), 16, seed)
        add(f'framelen{fl}', cfg, ol.encode_stream(cfg, x))
    # -- partial frames of every awkward length (incl. 0), escape + partial ------------------------------
    for n in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 31, 32, 33, 63, 64, 65, 4095):
        for esc in (0, 1):
            seed += 1
            cfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
            x = _sig(2, n, 24, seed, 'music')
            add(f'partial{n}_esc{esc}', cfg, [ol.encode_packet(cfg, x, ol.PacketOpts.make(force_escape=esc, min_order=8, max_order=8))])
    for bits in (16, 20, 24, 32):
        for ch in (1, 2, 3, 8):
            seed += 1
            cfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000)
            x = _sig(ch, 4096 + 123, bits, seed, 'white')
            add(f'escape_d{bits}_c{ch}', cfg, ol.encode_stream(cfg, x, ol.PacketOpts.make(force_escape=1)))
    # -- DSE / FIL / tag 3 / END placement -------------------------------------------------------------------
    for ch in (1, 2, 6, 8):
        for (dse, fil, lfe3, no_end, ap) in ((7, 0, 0, 0, 0), (0, 9, 0, 0, 0), (300, 200, 1, 0, 1), (255, 15, 1, 1, 0), (1, 14, 0, 1, 1)):
            seed += 1
            cfg = ol.Config.make(bit_depth=16, num_channels=ch, sample_rate=48000)
            x = _sig(ch, 4096 + 50, 16, seed)
            add(f'extras_c{ch}_dse{dse}_fil{fil}_lfe{lfe3}_noend{no_end}', cfg,
                ol.encode_stream(cfg, x, ol.PacketOpts.make(dse_bytes=dse, fil_bytes=fil, lfe_tag3=lfe3, no_end=no_end, always_partial=ap)))
    # -- element-level shapes: under-filled packets, odd element orders, unaligned DSE --------------------
    rng = np.random.default_rng(77)
    for ch in (2, 3, 4, 6, 7, 8):
        cfg = ol.Config.make(bit_depth=16, num_channels=ch, sample_rate=48000)
        a = _sig(1, 4096, 16, 2000 + ch, 'music')[:, 0]
        b = _sig(1, 4096, 16, 2100 + ch, 'music')[:, 0] // 2
        # END before all channels are present: the rest of the frame stays zero (appendix B6)
        add(f'underfilled_c{ch}', cfg, [ol.Writer(cfg).element(0, a, order=4, coefs=[60, -30, 10, 5]).end().bytes()])
        # a pair first (non-canonical order): for 3/6/7/8 channels some pairs land on the last channel
        w = ol.Writer(cfg)
        idx = 0
        while idx + 2 <= ch:
            w.element(1, a[:4000], b[:4000], order=5, coefs=[80, -40, 20, -10, 5], mix_bits=2, mix_res=1)
            idx += 2
        if idx < ch:
            w.element(3, b[:4000], order=6, coefs=[90, -45, 22, -11, 5, -2])
        add(f'pairs_first_partial_c{ch}', cfg, [w.end().bytes()])
        w = ol.Writer(cfg)
        idx = 0
        while idx + 2 <= ch:
            w.element(1, a, b, order=4, coefs=[80, -40, 20, -10])
            idx += 2
        if idx < ch:
            w.element(0, b, order=8, coefs=[90, -45, 22, -11, 5, -2, 1, 0])
        add(f'pairs_first_full_c{ch}', cfg, [w.end().bytes()])
        # mixed per-element sample counts
        w = ol.Writer(cfg).dse(5, align=0)
        for k in range(ch):
            nk = [4096, 100, 4096, 7, 2000, 1, 4096, 333][k]
            w.element(0 if k % 3 else 3, a[:nk], order=[4, 0, 31, 2, 5, 6, 8, 9][k], coefs=[40, -20, 10, -5, 2, -1, 1, 0, 0], partial=1, mode=k & 1)
        add(f'mixed_counts_c{ch}', cfg, [w.fil(3).bytes()])
    # stereo element with different params per channel
    cfg = ol.Config.make(bit_depth=24, num_channels=2, sample_rate=96000)
    a = _sig(1, 4096, 24, 3001, 'music')[:, 0] // 2
    b = _sig(1, 4096, 24, 3002, 'music')[:, 0] // 2
    for (ou, ov, mu, mv) in ((4, 8, 0, 1), (6, 5, 1, 0), (12, 0, 0, 0), (31, 4, 0, 0), (0, 31, 2, 2), (8, 8, 0, 0), (5, 20, 0, 0)):
        cu = [100, -50, 25, -12, 6, -3, 2, -1] + [0] * 24
        w = ol.Writer(cfg).element(1, a, b, order=ou, order_v=ov, mode=mu, mode_v=mv, coefs=cu, coefs_v=cu[::-1][-32:],
                                   den_shift=9, den_shift_v=7, pb_factor=4, pb_factor_v=3, bytes_shifted=1, mix_bits=2, mix_res=1)
        add(f'cpe_u{ou}_v{ov}_m{mu}{mv}', cfg, [w.end().bytes()])
    _ = rng
    return out


def frame_length_cases():
    """Frame lengths around every granularity of the kernel (16-sample entropy batches, 32-sample ring slots, 8- and
    16-frame emit rows) and the extremes, for mono / stereo / 3 and 6 channels: two full packets and a short last one
    each, from a signal with silence and +-LSB passages (zero runs that end exactly at, or run across, those edges)."""
    out = []
    seed = 7000
    for fl in (1, 2, 3, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 4095, 4097, 65536):
        for ch, bits in ((1, 16), (2, 24), (3, 24), (6, 16), (2, 16)):
            if fl == 65536 and ch > 2:
                continue
            seed += 1
            cfg = ol.Config.make(bit_depth=bits, num_channels=ch, frame_length=fl, sample_rate=48000)
            n = 2 * fl + max(1, fl // 3)
            x = _sig(ch, n, bits, seed, 'silence_lsb' if seed % 2 else 'music')
            out.append((f'fl{fl}_c{ch}_d{bits}', cfg, ol.encode_stream(cfg, x)))
    return out


def entropy_edge_cases():
    """Shapes that stress the entropy stage's bookkeeping: cookie parameters at the edges (kb 0 / 1 / 22 / 23 / 31 /
    255, mb and pb 0 / 255), digital silence of every length around the batch / ring-slot / stream sizes (one long zero
    run per stream, escape-coded when it is long), single impulses at those edges, +-LSB noise (a run-length code after
    nearly every sample) and loud clipping noise (escape codes), as full and as partial packets."""
    out = []
    seed = 9000
    # cookie parameters x signal kinds
    for pb, mb, kb in ((40, 10, 14), (40, 10, 0), (40, 10, 1), (40, 10, 2), (40, 10, 22), (40, 10, 23), (40, 10, 31), (40, 10, 255),
                       (0, 10, 14), (255, 10, 14), (40, 0, 14), (40, 255, 14), (1, 1, 9), (255, 255, 255)):
        for kind, bits in (('lsb', 24), ('silence_lsb', 16), ('loud', 16), ('white', 24)):
            seed += 1
            cfg = ol.Config.make(bit_depth=bits, num_channels=2, sample_rate=48000, pb=pb, mb=mb, kb=kb)
            x = _sig(2, 4096 + 1234, bits, seed, kind)
            out.append((f'ag_pb{pb}_mb{mb}_kb{kb}_{kind}{bits}', cfg, ol.encode_stream(cfg, x)))
    # silence and impulses at the edges
    for fl in (16, 17, 32, 33, 48, 4096, 65536):
        cfg = ol.Config.make(bit_depth=16, num_channels=2, frame_length=fl, sample_rate=48000)
        z = np.zeros((fl, 2), dtype=np.int64)
        packets = [ol.encode_packet(cfg, z)]
        for n in {1, min(fl, 15), min(fl, 16), min(fl, 17), fl - 1}:
            if n > 0:
                packets.append(ol.encode_packet(cfg, z[:n]))  # partial silent packets
        for pos in {0, 1, 14, 15, 16, 17, 31, 32, fl // 2, fl - 2, fl - 1}:
            if 0 <= pos < fl:
                y = z.copy()
                y[pos, 0] = 1000
                y[pos, 1] = -3
                packets.append(ol.encode_packet(cfg, y))
        # long runs of zeros between two non-zero stretches: run lengths around 16 / 32 / 255 / 256 / 4095
        for run in (15, 16, 17, 31, 32, 33, 255, 256, 257, 4000):
            if run + 8 <= fl:
                y = z.copy()
                y[:4] = [[5, -5], [7, 1], [-2, 2], [1, 1]]
                y[4 + run:4 + run + 4] = [[3, 3], [-1, -8], [2, 0], [9, 9]]
                packets.append(ol.encode_packet(cfg, y))
        out.append((f'silence_edges_fl{fl}', cfg, packets))
    return out


def hostile_cases(max_per_seed=40):
    """Mutated packets: (name, cfg, [packets]). Statuses (incl. ST_REF_PANIC) must match the oracle."""
    out = []
    rng = np.random.default_rng(4242)
    bases = []
    for bits, ch in ((16, 2), (24, 2), (16, 6), (24, 1), (32, 2), (20, 3)):
        cfg = ol.Config.make(bit_depth=bits, num_channels=ch, sample_rate=48000, frame_length=1024)
        x = make_signal('silence_lsb', ch, 2048, bits, 48000, seed=bits * 10 + ch)
        for opts in (ol.PacketOpts.make(), ol.PacketOpts.make(force_escape=1), ol.PacketOpts.make(min_order=12, max_order=12, dse_bytes=3, fil_bytes=2)):
            bases.append((cfg, ol.encode_stream(cfg, x, opts)))
    for bi, (cfg, pkts) in enumerate(bases):
        muts = []
        for p in pkts:
            p = bytearray(p)
            for _ in range(max_per_seed):
                q = bytearray(p)
                kind = rng.integers(0, 6)
                if kind == 0:  # truncate
                    q = q[:int(rng.integers(0, len(q)))]
                elif kind == 1:  # flip bits in the header region
                    for _k in range(int(rng.integers(1, 4))):
                        pos = int(rng.integers(0, min(len(q), 40)))
                        q[pos] ^= 1 << int(rng.integers(0, 8))
                elif kind == 2:  # flip bits anywhere
                    for _k in range(int(rng.integers(1, 6))):
                        pos = int(rng.integers(0, len(q)))
                        q[pos] ^= 1 << int(rng.integers(0, 8))
                elif kind == 3:  # random garbage
                    q = bytearray(rng.integers(0, 256, size=int(rng.integers(1, 300)), dtype=np.uint8).tobytes())
                elif kind == 4:  # truncate to a few bytes past the headers
                    q = q[:int(rng.integers(1, min(len(q), 64)))]
                else:  # overwrite the tail with ones (long unary prefixes / escapes at the end)
                    k = int(rng.integers(1, 12))
                    q[-k:] = b'\xff' * k
                muts.append(bytes(q))
        muts.append(b'')
        out.append((f'hostile{bi}_d{cfg.bit_depth}_c{cfg.num_channels}', cfg, muts))
    return out
