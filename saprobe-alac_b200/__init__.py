"""saprobe-alac_b200 -- Python host binding over the C ABI (include/alac_b200.h).

Mirrors the reference's exported Go API name for name so the parity tests read like the reference's
own tests (/root/reference/README.md:38-54):

    ParseMagicCookie(cookie) -> PacketConfig                  config.go:47
    NewPacketDecoder(config) -> PacketDecoder                 decoder.go:90
    PacketDecoder.DecodePacket(packet) -> bytes               decoder.go:117
    PacketDecoder.DecodePackets(packets) -> (pcm list, errs)  NEW (north star): one batched GPU call
    PacketDecoder.Format() -> PCMFormat                       decoder.go:112
    NewDecoder(file bytes | file object) -> Decoder           decode.go:50
    Decoder.Read(n) / Seek(ns) / Format() / Duration() / Position()   decode.go:78-190

Errors mirror errors.go:22-34: ErrConfig, ErrNoTrack, ErrDecode (exception classes; `.status` is the
status word of the C ABI, `str()` the reference's wrapped message).

All decoding happens in libalacb200.so (hand-written CUDA, no CPU fallback): importing this module
fails loudly if the library is missing, and creating a decoder fails if no CUDA device is usable.
The directory name contains a hyphen, so load it with `load_package()` from the repo-root helper
`alac_b200_loader.py` (registered in sys.modules as `saprobe_alac_b200`).
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ALACB200_LIB: developer override (an experiment or `make dev` build of the same ABI, e.g. libalacb200_dev.so)
LIB_PATH = os.environ.get('ALACB200_LIB') or os.path.join(_HERE, 'libalacb200.so')

# ---- API results / status words (include/alac_b200.h) ----------------------------------------------
OK = 0
E_ARG, E_CUDA, E_NO_DEVICE, E_NOMEM, E_CONFIG, E_IO, E_NO_TRACK = -1, -2, -3, -4, -5, -6, -7
ST_OK = 0
ST_INVALID_COOKIE, ST_UNSUPPORTED_VERSION, ST_UNSUPPORTED_ELEMENT, ST_INVALID_HEADER = 1, 2, 3, 4
ST_INVALID_SHIFT, ST_BITSTREAM_OVERRUN, ST_SAMPLE_OVERRUN, ST_BIT_DEPTH = 5, 6, 7, 8
ST_REF_PANIC, ST_UNSUPPORTED_CONFIG = 9, 10


class PacketConfig(C.Structure):
    """PacketConfig, config.go:27-38 == alacb200_config."""
    _fields_ = [('FrameLength', C.c_uint32), ('BitDepth', C.c_uint8), ('NumChannels', C.c_uint8),
                ('PB', C.c_uint8), ('MB', C.c_uint8), ('KB', C.c_uint8), ('_reserved', C.c_uint8),
                ('MaxRun', C.c_uint16), ('MaxFrameBytes', C.c_uint32), ('AvgBitRate', C.c_uint32),
                ('SampleRate', C.c_uint32)]

    def __repr__(self):
        return ('PacketConfig(' + ', '.join(f'{n}={getattr(self, n)}' for n, _ in self._fields_ if n[0] != '_') + ')')


@dataclass(frozen=True)
class PCMFormat:
    """PCMFormat, format.go:20-24."""
    SampleRate: int
    BitDepth: int
    Channels: int


class _PcmFormatC(C.Structure):
    _fields_ = [('sample_rate', C.c_int32), ('bit_depth', C.c_int32), ('channels', C.c_int32)]


class Profile(C.Structure):
    _fields_ = [('launches_decode', C.c_uint64), ('launches_emit', C.c_uint64), ('ms_decode', C.c_double),
                ('ms_emit', C.c_double)]


class SampleInfo(C.Structure):
    """SampleInfo, internal/mp4/mp4.go:28-31."""
    _fields_ = [('Offset', C.c_uint64), ('Size', C.c_uint32), ('_reserved', C.c_uint32)]


# ---- errors (errors.go:22-34) -------------------------------------------------------------------------
class AlacError(Exception):
    status = 0


class ErrConfig(AlacError):
    """invalid configuration"""


class ErrNoTrack(AlacError):
    """no track found"""


class ErrDecode(AlacError):
    """decode failed"""


class CudaError(RuntimeError):
    """The CUDA path is unusable (no device, driver failure). There is no CPU fallback."""


_CONFIG_CODES = (ST_INVALID_COOKIE, ST_UNSUPPORTED_VERSION, ST_BIT_DEPTH, ST_UNSUPPORTED_CONFIG)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(nvcc, sm_100a). saprobe-alac_b200 has no CPU fallback.')
    L = C.CDLL(LIB_PATH)
    vp, u8p = C.c_void_p, C.c_char_p
    sig = {
        'alacb200_parse_cookie': (C.c_int32, [u8p, C.c_size_t, C.POINTER(PacketConfig)]),
        'alacb200_bytes_per_sample': (C.c_int32, [C.c_uint8]),
        'alacb200_create': (C.c_int32, [C.POINTER(PacketConfig), C.c_int, C.POINTER(vp), C.POINTER(C.c_int32)]),
        'alacb200_destroy': (None, [vp]),
        'alacb200_format': (C.c_int32, [vp, C.POINTER(_PcmFormatC)]),
        'alacb200_get_config': (C.c_int32, [vp, C.POINTER(PacketConfig)]),
        'alacb200_max_packet_pcm_bytes': (C.c_uint64, [vp]),
        'alacb200_decode_packets': (C.c_int32, [vp, vp, vp, vp, C.c_uint32, vp, C.c_uint64, vp, vp]),
        'alacb200_decode_packets_device': (C.c_int32, [vp, vp, C.c_uint64, vp, vp, C.c_uint32, vp, C.c_uint64, vp, vp, vp]),
        'alacb200_pinned_alloc': (vp, [C.c_size_t]),
        'alacb200_pinned_free': (None, [vp]),
        'alacb200_strerror': (C.c_char_p, [C.c_int32]),
        'alacb200_format_error': (C.c_size_t, [C.c_int32, C.c_char_p, C.c_size_t]),
        'alacb200_last_error': (C.c_char_p, []),
        'alacb200_device_count': (C.c_int32, []),
        'alacb200_set_profiling': (C.c_int32, [vp, C.c_int]),
        'alacb200_get_profile': (C.c_int32, [vp, C.POINTER(Profile)]),
        'alacb200_mp4_find_alac_track': (C.c_int32, [vp, C.c_uint64, C.POINTER(vp)]),
        'alacb200_mp4_free_track': (None, [vp]),
        'alacb200_mp4_cookie': (vp, [vp, C.POINTER(C.c_size_t)]),
        'alacb200_mp4_samples': (C.POINTER(SampleInfo), [vp, C.POINTER(C.c_uint64)]),
        'alacb200_mp4_error': (C.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


lib = _load()
ABI_SYMBOLS = ('alacb200_parse_cookie alacb200_bytes_per_sample alacb200_create alacb200_destroy alacb200_format '
               'alacb200_get_config alacb200_max_packet_pcm_bytes alacb200_decode_packets '
               'alacb200_decode_packets_device alacb200_pinned_alloc alacb200_pinned_free alacb200_strerror '
               'alacb200_format_error alacb200_last_error alacb200_device_count alacb200_set_profiling '
               'alacb200_get_profile alacb200_mp4_find_alac_track alacb200_mp4_free_track alacb200_mp4_cookie '
               'alacb200_mp4_samples alacb200_mp4_error').split()


def format_error(status: int) -> str:
    buf = C.create_string_buffer(256)
    lib.alacb200_format_error(int(status), buf, 256)
    return buf.value.decode()


def error_from_status(status: int, prefix: str = '') -> AlacError:
    """Rebuild the reference's error chain from a status word (what the cgo shim does in Go)."""
    cls = ErrConfig if (int(status) & 0xff) in _CONFIG_CODES else ErrDecode
    err = cls(prefix + format_error(status))
    err.status = int(status)
    return err


def _check(rc: int, what: str):
    if rc != OK:
        raise CudaError(f'{what} failed (rc={rc}): {lib.alacb200_last_error().decode()}')


# ---- config.go ----------------------------------------------------------------------------------------
def ParseMagicCookie(cookie: bytes) -> PacketConfig:
    cfg = PacketConfig()
    cookie = bytes(cookie) if cookie is not None else b''
    st = lib.alacb200_parse_cookie(cookie, len(cookie), C.byref(cfg))
    if st != ST_OK:
        raise error_from_status(st)
    return cfg


def BytesPerSample(depth: int) -> int:
    n = lib.alacb200_bytes_per_sample(depth)
    if n == 0:
        raise ValueError(f'alac: BytesPerSample called with unsupported bit depth {depth}')  # format.go:32 panics
    return n


def pack_packets(packets, align=16):
    """Host packer: list of packet bytes -> (packed u8 [+64 B pad], offsets u64, sizes u32), 16-byte aligned."""
    sizes = np.fromiter((len(p) for p in packets), dtype=np.uint32, count=len(packets))
    padded = (sizes.astype(np.uint64) + (align - 1)) // align * align
    offsets = np.zeros(len(packets), dtype=np.uint64)
    if len(packets) > 1:
        offsets[1:] = np.cumsum(padded)[:-1]
    total = int(padded.sum()) if len(packets) else 0
    packed = np.zeros(total + 64, dtype=np.uint8)
    for p, o in zip(packets, offsets):
        packed[int(o):int(o) + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return packed, offsets, sizes


def shard_ranges(weights, world: int):
    """Contiguous [lo, hi) ranges, one per rank, balanced by `weights` (compressed bytes per packet or per
    track). Packets/tracks are independent (decoder.go:79-87), so multi-GPU decode is a plain partition:
    no collective on the data path, output order = input order."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    world = max(1, int(world))
    if n == 0:
        return [(0, 0)] * world
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side='left'))
        k = max(bounds[-1], min(k, n))
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


# ---- decoder.go ---------------------------------------------------------------------------------------
class PacketDecoder:
    """PacketDecoder, decoder.go:79-128, on one CUDA device."""

    def __init__(self, config: PacketConfig, device: int = 0):
        self._h = C.c_void_p()
        st = C.c_int32(0)
        rc = lib.alacb200_create(C.byref(config), device, C.byref(self._h), C.byref(st))
        if rc == E_CONFIG:
            err = error_from_status(st.value)
            if st.value == ST_BIT_DEPTH:
                err.args = (f'{err.args[0]}: {config.BitDepth}',)  # decoder.go:92
            raise err
        _check(rc, 'alacb200_create')
        self.config = config
        self.device = device
        self.frame_bytes = int(lib.alacb200_max_packet_pcm_bytes(self._h))

    def close(self):
        if getattr(self, '_h', None):
            lib.alacb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the module globals may already be gone
            pass

    def Format(self) -> PCMFormat:
        f = _PcmFormatC()
        _check(lib.alacb200_format(self._h, C.byref(f)), 'alacb200_format')
        return PCMFormat(f.sample_rate, f.bit_depth, f.channels)

    # -- batched entry points ------------------------------------------------------------------------
    def decode_packed(self, packed, offsets, sizes, out=None, out_stride=None):
        """Raw batched call on host arrays -> (pcm [n, out_stride] u8, out_bytes u32 [n], status i32 [n])."""
        n = len(sizes)
        out_stride = out_stride or (self.frame_bytes + 3) // 4 * 4
        if out is None:
            out = np.zeros((n, out_stride), dtype=np.uint8)
        nb = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int32)
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        rc = lib.alacb200_decode_packets(self._h, packed.ctypes.data, offsets.ctypes.data, sizes.ctypes.data, n,
                                         out.ctypes.data, out_stride, nb.ctypes.data, st.ctypes.data)
        _check(rc, 'alacb200_decode_packets')
        return out, nb, st

    def DecodePackets(self, packets):
        """Batched DecodePacket: -> (list of PCM bytes or None, list of error or None)."""
        if len(packets) == 0:
            return [], []
        packed, offsets, sizes = pack_packets(packets)
        out, nb, st = self.decode_packed(packed, offsets, sizes)
        pcm, errs = [], []
        for i in range(len(packets)):
            if st[i] == ST_OK:
                pcm.append(out[i, :nb[i]].tobytes())
                errs.append(None)
            else:
                pcm.append(None)
                errs.append(error_from_status(int(st[i])))
        return pcm, errs

    def DecodePacket(self, packet: bytes) -> bytes:
        pcm, errs = self.DecodePackets([bytes(packet)])
        if errs[0] is not None:
            raise errs[0]
        return pcm[0]

    # -- profiling -----------------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        _check(lib.alacb200_set_profiling(self._h, int(on)), 'alacb200_set_profiling')

    def get_profile(self) -> Profile:
        p = Profile()
        _check(lib.alacb200_get_profile(self._h, C.byref(p)), 'alacb200_get_profile')
        return p


def NewPacketDecoder(config: PacketConfig, device: int = 0) -> PacketDecoder:
    return PacketDecoder(config, device)


# ---- internal/mp4 ------------------------------------------------------------------------------------
def FindALACTrack(data: bytes):
    """FindALACTrack, internal/mp4/mp4.go:233-300 -> (cookie bytes, [(offset, size)])."""
    data = bytes(data)
    t = C.c_void_p()
    rc = lib.alacb200_mp4_find_alac_track(data, len(data), C.byref(t))
    try:
        if rc != OK:
            e = ErrNoTrack('no track found: ' + lib.alacb200_mp4_error(t).decode())  # decode.go:53
            raise e
        n = C.c_size_t()
        p = lib.alacb200_mp4_cookie(t, C.byref(n))
        cookie = C.string_at(p, n.value) if p else b''
        cnt = C.c_uint64()
        sp = lib.alacb200_mp4_samples(t, C.byref(cnt))
        samples = [(sp[i].Offset, sp[i].Size) for i in range(cnt.value)]
        return cookie, samples
    finally:
        lib.alacb200_mp4_free_track(t)


# ---- decode.go ----------------------------------------------------------------------------------------
_NS = 1_000_000_000


class Decoder:
    """Streaming Decoder, decode.go:32-190, with a GPU read-ahead window behind Read/Seek.

    The reference decodes one packet per Read loop iteration (decode.go:159-186); here Read pulls a
    window of `window` packets through one batched GPU call and serves bytes from it. Packet order,
    io.Reader semantics (short reads, EOF as b''), Seek's packet alignment and Duration's over-count
    of a partial last packet (decode.go:82-88) are unchanged.
    """

    def __init__(self, rs, device: int = 0, window: int = 2048):
        data = rs if isinstance(rs, (bytes, bytearray, memoryview)) else rs.read()
        self._data = bytes(data)
        cookie, self.samples = FindALACTrack(self._data)
        try:
            config = ParseMagicCookie(cookie)
        except AlacError as e:
            e.args = ('parsing ALAC config: ' + e.args[0],)  # decode.go:58
            raise
        self.dec = NewPacketDecoder(config, device)
        self.sampleIdx = 0
        self._window = max(1, int(window))
        self._ready = []      # decoded (pcm bytes | error) of packets [self._ready_base, ...)
        self._ready_base = 0
        self._buf = b''
        self._bufOff = 0
        self._eof = False

    def Format(self) -> PCMFormat:
        return self.dec.Format()

    def Duration(self) -> int:
        """nanoseconds, decode.go:82-88"""
        c = self.dec.config
        return len(self.samples) * c.FrameLength * _NS // c.SampleRate

    def Position(self) -> int:
        c = self.dec.config
        return self.sampleIdx * c.FrameLength * _NS // c.SampleRate

    def Seek(self, t_ns: int) -> int:
        """decode.go:103-124: packet-granular, clamped to [0, len(samples)]; returns the aligned time."""
        c = self.dec.config
        target_frame = int((t_ns / _NS) * float(c.SampleRate))
        q = abs(target_frame) // c.FrameLength
        target = -q if target_frame < 0 else q  # Go integer division truncates toward zero
        target = max(0, min(target, len(self.samples)))
        self.sampleIdx = target
        self._buf, self._bufOff = b'', 0
        self._eof = target >= len(self.samples)
        return self.sampleIdx * c.FrameLength * _NS // c.SampleRate

    def _fill(self):
        idx = self.sampleIdx
        if not (self._ready_base <= idx < self._ready_base + len(self._ready)):
            hi = min(len(self.samples), idx + self._window)
            packets = []
            for k in range(idx, hi):
                off, size = self.samples[k]
                pkt = self._data[off:off + size]
                if len(pkt) != size:  # io.ReadFull failure, decode.go:172-174
                    packets.append(None)
                    break
                packets.append(pkt)
            good = [p for p in packets if p is not None]
            pcm, errs = self.dec.DecodePackets(good)
            self._ready = [e if e is not None else p for p, e in zip(pcm, errs)]
            if len(good) < len(packets):
                self._ready.append(IOError(f'reading sample {idx + len(good)}: unexpected EOF'))
            self._ready_base = idx
        item = self._ready[idx - self._ready_base]
        if isinstance(item, AlacError):
            err = type(item)(f'decoding packet {idx}: {item.args[0]}')  # decode.go:181
            err.status = item.status
            raise err
        if isinstance(item, Exception):
            raise item
        self._buf, self._bufOff = item, 0
        self.sampleIdx += 1

    def Read(self, n: int) -> bytes:
        """io.Reader: up to n bytes; b'' means io.EOF (decode.go:127-190)."""
        out = bytearray()
        while len(out) < n:
            if self._bufOff < len(self._buf):
                take = min(n - len(out), len(self._buf) - self._bufOff)
                out += self._buf[self._bufOff:self._bufOff + take]
                self._bufOff += take
                continue
            if self._eof or self.sampleIdx >= len(self.samples):
                self._eof = True
                break
            try:
                self._fill()
            except Exception:
                if out:  # Go returns (total, err): hand out the bytes now, the error on the next call
                    break
                raise
        return bytes(out)

    def ReadAll(self) -> bytes:
        chunks = []
        while True:
            b = self.Read(1 << 24)
            if not b:
                return b''.join(chunks)
            chunks.append(b)


def NewDecoder(rs, device: int = 0, window: int = 2048) -> Decoder:
    return Decoder(rs, device, window)
