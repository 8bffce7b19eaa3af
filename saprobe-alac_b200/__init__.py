"""saprobe-alac_b200 -- Python host binding over the C ABI (include/alac_b200.h).

Mirrors the reference's exported Go API name for name so the parity tests read like the reference's
own tests (/root/reference/README.md:38-54):

    ParseMagicCookie(cookie) -> PacketConfig                  config.go:47
    NewPacketDecoder(config) -> PacketDecoder                 decoder.go:90
    PacketDecoder.DecodePacket(packet) -> bytes               decoder.go:117
    PacketDecoder.DecodePackets(packets) -> (pcm list, errs)  NEW (north star): one batched GPU call
    PacketDecoder.Format() -> PCMFormat                       decoder.go:112
    NewLibraryDecoder(devices).DecodeTracks(tracks)           NEW: many tracks of mixed cookies, sharded over devices
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
ST_REF_PANIC, ST_UNSUPPORTED_CONFIG, ST_IO_TRUNCATED = 9, 10, 11


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
    _fields_ = [('launches_decode', C.c_uint64), ('ms_decode', C.c_double)]


class TrackDesc(C.Structure):
    """alacb200_track_desc (include/alac_b200.h): one track of a library batch."""
    _fields_ = [('cookie', C.c_void_p), ('cookie_len', C.c_size_t), ('data', C.c_void_p), ('data_len', C.c_uint64),
                ('offsets', C.c_void_p), ('sizes', C.c_void_p), ('n', C.c_uint32), ('reserved', C.c_uint32),
                ('pcm_out', C.c_void_p), ('out_stride', C.c_uint64), ('out_bytes', C.c_void_p), ('status', C.c_void_p),
                ('config', PacketConfig), ('track_status', C.c_int32), ('result', C.c_int32), ('device', C.c_int32),
                ('reserved2', C.c_int32)]


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
        'alacb200_decode_packets': (C.c_int32, [vp, vp, C.c_uint64, vp, vp, C.c_uint32, vp, C.c_uint64, vp, vp]),
        'alacb200_arena': (C.c_int32, [vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]),
        'alacb200_library_create': (C.c_int32, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        'alacb200_library_destroy': (None, [vp]),
        'alacb200_library_devices': (C.c_int32, [vp]),
        'alacb200_library_decode_tracks': (C.c_int32, [vp, C.POINTER(TrackDesc), C.c_uint32]),
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
               'alacb200_get_config alacb200_max_packet_pcm_bytes alacb200_decode_packets alacb200_arena '
               'alacb200_library_create alacb200_library_destroy alacb200_library_devices '
               'alacb200_library_decode_tracks alacb200_decode_packets_device alacb200_pinned_alloc alacb200_pinned_free alacb200_strerror '
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
        """Raw batched call on host arrays -> (pcm [n, out_stride] u8, out_bytes u32 [n], status i32 [n]). `packed` may be
        any byte buffer the packets live in (a packed array or a whole file image)."""
        n = len(sizes)
        out_stride = out_stride or (self.frame_bytes + 3) // 4 * 4
        if out is None:
            out = np.zeros((n, out_stride), dtype=np.uint8)
        nb = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int32)
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        rc = lib.alacb200_decode_packets(self._h, packed.ctypes.data, packed.nbytes, offsets.ctypes.data, sizes.ctypes.data, n,
                                         out.ctypes.data, out_stride, nb.ctypes.data, st.ctypes.data)
        _check(rc, 'alacb200_decode_packets')
        return out, nb, st

    def _arena(self, in_bytes, out_bytes):
        """The decoder's own pinned staging (grow-only): -> (in u8 [in_bytes], out u8 [out_bytes]) views."""
        pi, po = C.c_void_p(), C.c_void_p()
        _check(lib.alacb200_arena(self._h, in_bytes, out_bytes, C.byref(pi), C.byref(po)), 'alacb200_arena')
        a = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint8)), shape=(max(in_bytes, 1),))
        b = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint8)), shape=(max(out_bytes, 1),))
        return a, b

    def decode_in_place(self, data_ptr, data_len, offsets, sizes):
        """Packets read in place from `data_ptr` (a pinned file image or packed buffer) -> (pcm rows, out_bytes, status);
        the PCM lands in the decoder's pinned arena (valid until the next call)."""
        n = len(sizes)
        stride = (self.frame_bytes + 3) // 4 * 4
        _, out = self._arena(0, n * stride)
        out = out[:n * stride].reshape(n, stride)
        nb = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int32)
        rc = lib.alacb200_decode_packets(self._h, data_ptr, data_len, offsets.ctypes.data, sizes.ctypes.data, n,
                                         out.ctypes.data, stride, nb.ctypes.data, st.ctypes.data)
        _check(rc, 'alacb200_decode_packets')
        return out, nb, st

    def DecodePackets(self, packets):
        """Batched DecodePacket: -> (list of PCM bytes or None, list of error or None). The packets are packed into the
        decoder's pinned arena and the PCM comes back through it: no allocation of pinned memory per call."""
        n = len(packets)
        if n == 0:
            return [], []
        sizes = np.fromiter((len(p) for p in packets), dtype=np.uint32, count=n)
        padded = (sizes.astype(np.uint64) + 15) // 16 * 16
        offsets = np.zeros(n, dtype=np.uint64)
        if n > 1:
            offsets[1:] = np.cumsum(padded)[:-1]
        total = int(padded.sum())
        stride = (self.frame_bytes + 3) // 4 * 4
        a_in, a_out = self._arena(total, n * stride)
        for p, o in zip(packets, offsets):
            a_in[int(o):int(o) + len(p)] = np.frombuffer(p, dtype=np.uint8)
        out = a_out[:n * stride].reshape(n, stride)
        nb = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int32)
        rc = lib.alacb200_decode_packets(self._h, a_in.ctypes.data, total, offsets.ctypes.data, sizes.ctypes.data, n,
                                         out.ctypes.data, stride, nb.ctypes.data, st.ctypes.data)
        _check(rc, 'alacb200_decode_packets')
        pcm, errs = [], []
        for i in range(n):
            if st[i] == ST_OK:
                pcm.append(out[i, :nb[i]].tobytes())  # a fresh buffer per packet, like decoder.go:120
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


# ---- library batches: many tracks, mixed cookies, several devices ------------------------------------
@dataclass
class Track:
    """One track of a library batch: its cookie and the bytes + sample table its packets live in (a packed buffer from
    pack_packets, or a whole M4A image with FindALACTrack's table -- read in place, never re-packed)."""
    cookie: bytes
    data: np.ndarray     # uint8
    offsets: np.ndarray  # uint64 [n]
    sizes: np.ndarray    # uint32 [n]


@dataclass
class TrackResult:
    config: PacketConfig = None
    err: Exception = None        # ErrConfig for a bad cookie / unsupported depth: the track has no decoder
    pcm: np.ndarray = None       # uint8 [n, stride]; packet i is pcm[i, :out_bytes[i]]
    out_bytes: np.ndarray = None
    status: np.ndarray = None    # status word per packet (error_from_status rebuilds the reference's error)
    device: int = -1

    def packet(self, i) -> bytes:
        if self.status[i] != ST_OK:
            raise error_from_status(int(self.status[i]))
        return self.pcm[i, :self.out_bytes[i]].tobytes()

    def pcm_bytes(self) -> bytes:
        return b''.join(self.packet(i) for i in range(len(self.status)))


class LibraryDecoder:
    """Decodes whole libraries: any mix of cookies (a PacketDecoder holds one, decoder.go:79-87), contiguous track ranges
    per device balanced by compressed bytes, one submitting host thread per device, no collective."""

    def __init__(self, devices=(0,)):
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        _check(lib.alacb200_library_create(devs, len(devices), C.byref(self._h)), 'alacb200_library_create')
        self.devices = tuple(devices)

    def close(self):
        if getattr(self, '_h', None):
            lib.alacb200_library_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def DecodeTracks(self, tracks, out=None):
        """tracks: [Track] -> [TrackResult]. `out`: optional list of preallocated uint8 [n, stride] arrays (e.g. pinned)."""
        nt = len(tracks)
        descs = (TrackDesc * max(nt, 1))()
        keep, results = [], []
        for t, tr in enumerate(tracks):
            d = descs[t]
            cookie = np.frombuffer(bytes(tr.cookie) if tr.cookie is not None else b'', dtype=np.uint8)
            data = np.ascontiguousarray(tr.data, dtype=np.uint8)
            offs = np.ascontiguousarray(tr.offsets, dtype=np.uint64)
            sizes = np.ascontiguousarray(tr.sizes, dtype=np.uint32)
            n = len(sizes)
            cfg = PacketConfig()
            st = lib.alacb200_parse_cookie(bytes(tr.cookie) if tr.cookie is not None else b'', len(cookie), C.byref(cfg))
            bps = lib.alacb200_bytes_per_sample(cfg.BitDepth) if st == ST_OK else 0
            stride = (cfg.FrameLength * cfg.NumChannels * bps + 3) // 4 * 4 if bps else 4
            ok_shape = bps and 1 <= cfg.NumChannels <= 8 and 1 <= cfg.FrameLength <= 65536
            pcm = out[t] if out is not None else (np.zeros((n, stride), dtype=np.uint8) if ok_shape else np.zeros((0, 4), dtype=np.uint8))
            nb = np.zeros(n, dtype=np.uint32)
            stt = np.zeros(n, dtype=np.int32)
            d.cookie, d.cookie_len = cookie.ctypes.data if len(cookie) else None, len(cookie)
            d.data, d.data_len = data.ctypes.data, data.nbytes
            d.offsets, d.sizes, d.n = offs.ctypes.data, sizes.ctypes.data, n
            d.pcm_out, d.out_stride = pcm.ctypes.data, (pcm.shape[1] if pcm.ndim == 2 and pcm.shape[0] else stride)
            d.out_bytes, d.status = nb.ctypes.data, stt.ctypes.data
            keep.append((cookie, data, offs, sizes))
            results.append(TrackResult(pcm=pcm, out_bytes=nb, status=stt))
        rc = lib.alacb200_library_decode_tracks(self._h, descs, nt)
        _check(rc, 'alacb200_library_decode_tracks')
        for t, r in enumerate(results):
            d = descs[t]
            r.config = PacketConfig.from_buffer_copy(d.config)
            r.device = d.device
            if d.result == E_CONFIG:
                r.err = error_from_status(d.track_status)
                if d.track_status == ST_BIT_DEPTH:
                    r.err.args = (f'{r.err.args[0]}: {r.config.BitDepth}',)  # decoder.go:92
            elif d.result != OK:
                r.err = CudaError(f'track {t}: rc={d.result}')
        del keep
        return results


def NewLibraryDecoder(devices=(0,)) -> LibraryDecoder:
    return LibraryDecoder(devices)


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
        data = bytes(data)
        cookie, samples = FindALACTrack(data)
        try:
            config = ParseMagicCookie(cookie)
        except AlacError as e:
            e.args = ('parsing ALAC config: ' + e.args[0],)  # decode.go:58
            raise
        self.dec = NewPacketDecoder(config, device)
        # the file image lives in pinned memory for the life of the decoder: every window is read in place from it
        # (image + sample-table offsets straight to the C ABI, asynchronous H2D), nothing is re-packed
        self._img_len = len(data)
        self._img_ptr = lib.alacb200_pinned_alloc(max(1, self._img_len))
        if not self._img_ptr:
            raise CudaError('alacb200_pinned_alloc failed: ' + lib.alacb200_last_error().decode())
        C.memmove(self._img_ptr, data, self._img_len)
        self.samples = samples
        self._offs = np.array([o for o, _ in samples], dtype=np.uint64)
        self._sizes = np.array([z for _, z in samples], dtype=np.uint32)
        self.sampleIdx = 0
        self._window = max(1, int(window))
        self._ready = None    # (pcm rows, out_bytes, status) of packets [self._ready_base, ...)
        self._ready_n = 0
        self._ready_base = 0
        self._buf = b''
        self._bufOff = 0
        self._eof = False

    def close(self):
        if getattr(self, '_img_ptr', None):
            self.dec.close()
            lib.alacb200_pinned_free(self._img_ptr)
            self._img_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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
        if not (self._ready is not None and self._ready_base <= idx < self._ready_base + self._ready_n):
            hi = min(len(self.samples), idx + self._window)
            out, nb, st = self.dec.decode_in_place(self._img_ptr, self._img_len, self._offs[idx:hi], self._sizes[idx:hi])
            # the arena is reused by the next window: keep this window's PCM
            self._ready = (out.copy(), nb, st)
            self._ready_n = hi - idx
            self._ready_base = idx
        out, nb, st = self._ready
        k = idx - self._ready_base
        if st[k] == ST_IO_TRUNCATED:  # io.ReadFull failure, decode.go:172-174
            raise IOError(f'reading sample {idx}: unexpected EOF')
        if st[k] != ST_OK:
            item = error_from_status(int(st[k]))
            err = type(item)(f'decoding packet {idx}: {item.args[0]}')  # decode.go:181
            err.status = item.status
            raise err
        self._buf, self._bufOff = out[k, :nb[k]].tobytes(), 0
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
