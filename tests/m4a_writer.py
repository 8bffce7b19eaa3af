"""Minimal M4A/MP4 muxer for tests: ftyp, [free], mdat, moov/trak/mdia/minf/stbl(stsd,stts,stsc,stsz,stco|co64).

Knobs cover what internal/mp4 handles (mp4.go:233-420): chunking via stsc, constant-size stsz, co64,
QuickTime v1 sample entries, 64-bit and zero ("to end of file") box sizes, extra non-ALAC traks.
"""
import struct


def box(fourcc, payload, large=False):
    if large:
        return struct.pack('>I4sQ', 1, fourcc, 16 + len(payload)) + payload
    return struct.pack('>I4s', 8 + len(payload), fourcc) + payload


def full(fourcc, payload, version=0, flags=0):
    return box(fourcc, struct.pack('>I', (version << 24) | flags) + payload)


def sample_entry(fourcc, cookie, channels, bits, rate, qt_v1=False):
    body = b'\0' * 6 + struct.pack('>H', 1)                      # reserved + data_reference_index
    body += struct.pack('>HHI', 1 if qt_v1 else 0, 0, 0)         # version, revision, vendor
    body += struct.pack('>HHHHI', channels, bits, 0, 0, (rate & 0xffff) << 16)
    if qt_v1:
        body += struct.pack('>IIII', 4096, 0, 0, 0)
    return box(fourcc, body + cookie)


def build_m4a(cookie, packets, channels=2, bits=16, rate=44100, frame_length=4096, samples_per_chunk=None,
              co64=False, constant_stsz=False, qt_v1=False, moov_first=False, extra_trak=False, mdat_large=False,
              last_frames=None):
    """-> (file bytes, [(offset, size)]) ; cookie = the raw bytes stored after the sample entry (e.g. 36-byte alac atom)."""
    n = len(packets)
    spc = samples_per_chunk or max(1, n)
    sizes = [len(p) for p in packets]
    if constant_stsz:
        assert len(set(sizes)) <= 1
    ftyp = box(b'ftyp', b'M4A \0\0\0\0M4A mp42isom')
    free = box(b'free', b'\0' * 8)
    mdat_payload = b''.join(packets)
    mdat = box(b'mdat', mdat_payload, large=mdat_large)
    mdat_hdr = 16 if mdat_large else 8

    def make_moov(mdat_off):
        # chunk offsets
        offs, pos, sample_offsets = [], mdat_off + mdat_hdr, []
        for i in range(0, n, spc):
            offs.append(pos)
            for s in sizes[i:i + spc]:
                sample_offsets.append((pos, s))
                pos += s
        stsd = full(b'stsd', struct.pack('>I', 1) + sample_entry(b'alac', cookie, channels, bits, rate, qt_v1))
        last = last_frames if last_frames is not None else frame_length
        if n > 1 and last != frame_length:
            stts = full(b'stts', struct.pack('>IIIII', 2, n - 1, frame_length, 1, last))
        else:
            stts = full(b'stts', struct.pack('>III', 1, n, frame_length))
        nchunks = len(offs)
        entries = [(1, spc, 1)]
        if n % spc and nchunks > 1:
            entries.append((nchunks, n % spc, 1))
        elif n % spc:
            entries = [(1, n % spc, 1)]
        stsc = full(b'stsc', struct.pack('>I', len(entries)) + b''.join(struct.pack('>III', *e) for e in entries))
        if constant_stsz and n:
            stsz = full(b'stsz', struct.pack('>II', sizes[0], n))
        else:
            stsz = full(b'stsz', struct.pack('>II', 0, n) + b''.join(struct.pack('>I', s) for s in sizes))
        if co64:
            stco = full(b'co64', struct.pack('>I', nchunks) + b''.join(struct.pack('>Q', o) for o in offs))
        else:
            stco = full(b'stco', struct.pack('>I', nchunks) + b''.join(struct.pack('>I', o) for o in offs))
        stbl = box(b'stbl', stsd + stts + stsc + stsz + stco)
        minf = box(b'minf', full(b'smhd', b'\0' * 4) + stbl)
        mdia = box(b'mdia', full(b'mdhd', struct.pack('>IIIIHH', 0, 0, rate, n * frame_length, 0x55c4, 0)) + minf)
        traks = b''
        if extra_trak:  # a non-ALAC trak first: FindALACTrack must skip it (mp4.go:279-281)
            stsd2 = full(b'stsd', struct.pack('>I', 1) + sample_entry(b'mp4a', b'\0' * 8, 2, 16, 44100))
            stbl2 = box(b'stbl', stsd2)
            traks += box(b'trak', box(b'mdia', box(b'minf', stbl2)))
        traks += box(b'trak', full(b'tkhd', b'\0' * 80) + mdia)
        return box(b'moov', full(b'mvhd', b'\0' * 96) + traks), sample_offsets

    if moov_first:
        moov, _ = make_moov(0)
        mdat_off = len(ftyp) + len(moov)
        moov, so = make_moov(mdat_off)
        return ftyp + moov + mdat, so
    mdat_off = len(ftyp) + len(free)
    moov, so = make_moov(mdat_off)
    return ftyp + free + mdat + moov, so
