"""ctypes binding of the CPU oracle and the test-side encoder (oracle/build/libalac_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, 'oracle')
LIB_PATH = os.path.join(ORACLE_DIR, 'build', 'libalac_oracle.so')

OK = 0
ERR_INVALID_COOKIE, ERR_UNSUPPORTED_VERSION, ERR_UNSUPPORTED_ELEMENT, ERR_INVALID_HEADER = 1, 2, 3, 4
ERR_INVALID_SHIFT, ERR_BITSTREAM_OVERRUN, ERR_SAMPLE_OVERRUN, ERR_BIT_DEPTH = 5, 6, 7, 8
ERR_REF_PANIC, ERR_UNSUPPORTED_CONFIG = 9, 10
CTX_SCE, CTX_CPE, CTX_DSE, CTX_FIL = 1, 2, 3, 4


def code(status):
    return int(status) & 0xff


class Config(C.Structure):
    """PacketConfig, /root/reference/config.go:27-38 (same layout as ao_config / alacb200_config)."""
    _fields_ = [('frame_length', C.c_uint32), ('bit_depth', C.c_uint8), ('num_channels', C.c_uint8),
                ('pb', C.c_uint8), ('mb', C.c_uint8), ('kb', C.c_uint8), ('pad_', C.c_uint8),
                ('max_run', C.c_uint16), ('max_frame_bytes', C.c_uint32), ('avg_bit_rate', C.c_uint32),
                ('sample_rate', C.c_uint32)]

    @classmethod
    def make(cls, bit_depth=16, num_channels=2, frame_length=4096, sample_rate=44100, pb=40, mb=10, kb=14,
             max_run=255, max_frame_bytes=0, avg_bit_rate=0):
        return cls(frame_length, bit_depth, num_channels, pb, mb, kb, 0, max_run, max_frame_bytes, avg_bit_rate,
                   sample_rate)

    def bps(self):
        return {16: 2, 20: 3, 24: 3, 32: 4}[self.bit_depth]

    def frame_bytes(self):
        return self.frame_length * self.num_channels * self.bps()


class ChanParams(C.Structure):
    _fields_ = [('mode', C.c_uint8), ('den_shift', C.c_uint8), ('pb_factor', C.c_uint8), ('order', C.c_uint8),
                ('coefs', C.c_int16 * 32)]


class Element(C.Structure):
    _fields_ = [('tag', C.c_uint8), ('instance', C.c_uint8), ('escape', C.c_uint8), ('bytes_shifted', C.c_uint8),
                ('partial', C.c_uint8), ('mix_bits', C.c_uint8), ('mix_res', C.c_int8), ('pad_', C.c_uint8),
                ('ch', ChanParams * 2)]


class PacketOpts(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('min_order', 'max_order', 'den_shift', 'mode', 'pb_factor', 'mix_bits',
                                         'mix_res', 'bytes_shifted', 'force_escape', 'lfe_tag3', 'fil_bytes',
                                         'dse_bytes', 'no_end', 'always_partial')]

    @classmethod
    def make(cls, min_order=4, max_order=6, den_shift=9, mode=0, pb_factor=4, mix_bits=0, mix_res=-128,
             bytes_shifted=-1, force_escape=0, lfe_tag3=0, fil_bytes=0, dse_bytes=0, no_end=0, always_partial=0):
        return cls(min_order, max_order, den_shift, mode, pb_factor, mix_bits, mix_res, bytes_shifted, force_escape,
                   lfe_tag3, fil_bytes, dse_bytes, no_end, always_partial)


class _Writer(C.Structure):
    _fields_ = [('buf', C.c_void_p), ('cap', C.c_size_t), ('bitpos', C.c_uint64), ('overflow', C.c_int)]


_lib = None


def build():
    subprocess.run(['make', '-s', '-C', ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u8p, u32p, i32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_uint64))
        L.ao_parse_cookie.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Config)]
        L.ao_parse_cookie.restype = C.c_int32
        L.ao_check_config.argtypes = [C.POINTER(Config)]
        L.ao_check_config.restype = C.c_int32
        L.ao_decode_packet.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t, C.c_void_p, u32p]
        L.ao_decode_packet.restype = C.c_int32
        L.ao_decode_batch.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.ao_decode_batch.restype = None
        L.ae_encode_packet.argtypes = [C.POINTER(Config), C.POINTER(PacketOpts), C.c_void_p, C.c_uint32, C.c_void_p,
                                       C.c_size_t]
        L.ae_encode_packet.restype = C.c_int64
        L.ae_write_cookie.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p]
        L.ae_write_cookie.restype = C.c_size_t
        L.ae_writer_init.argtypes = [C.POINTER(_Writer), C.c_void_p, C.c_size_t]
        L.ae_put_bits.argtypes = [C.POINTER(_Writer), C.c_uint32, C.c_uint32]
        L.ae_byte_align.argtypes = [C.POINTER(_Writer)]
        L.ae_encode_element.argtypes = [C.POINTER(_Writer), C.POINTER(Config), C.POINTER(Element), C.c_void_p,
                                        C.c_void_p, C.c_uint32]
        L.ae_encode_element.restype = C.c_int
        L.ae_write_dse.argtypes = [C.POINTER(_Writer), C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        L.ae_write_fil.argtypes = [C.POINTER(_Writer), C.c_uint32]
        L.ae_write_end.argtypes = [C.POINTER(_Writer)]
        _lib = L
    return _lib


# ---- decoder side -------------------------------------------------------------------------------
def parse_cookie(cookie: bytes):
    cfg = Config()
    st = lib().ao_parse_cookie(cookie, len(cookie), C.byref(cfg))
    return st, cfg


def decode_packet(cfg: Config, packet: bytes):
    """-> (status, pcm bytes or None). DecodePacket on a fresh decoder."""
    st = lib().ao_check_config(C.byref(cfg))
    if st != OK:
        return st, None
    out = np.zeros(max(cfg.frame_bytes(), 1), dtype=np.uint8)
    nb = C.c_uint32(0)
    st = lib().ao_decode_packet(C.byref(cfg), bytes(packet), len(packet), out.ctypes.data, C.byref(nb))
    if st != OK:
        return st, None
    return st, out[:nb.value].tobytes()


def pack(packets, align=16):
    """list of bytes -> (packed uint8 array (+64 B tail pad), offsets u64, sizes u32)."""
    sizes = np.array([len(p) for p in packets], dtype=np.uint32)
    offs = np.zeros(len(packets), dtype=np.uint64)
    pos = 0
    for i, p in enumerate(packets):
        offs[i] = pos
        pos += (len(p) + align - 1) // align * align
    packed = np.zeros(pos + 64, dtype=np.uint8)
    for i, p in enumerate(packets):
        packed[int(offs[i]):int(offs[i]) + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return packed, offs, sizes


def decode_batch(cfg: Config, packed, offsets, sizes, nthreads=1, out=None):
    """-> (out [n, frame_bytes] uint8, out_bytes u32 [n], status i32 [n])."""
    n = len(sizes)
    stride = cfg.frame_bytes()
    if out is None:
        out = np.zeros((n, stride), dtype=np.uint8)
    nb = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
    lib().ao_decode_batch(C.byref(cfg), packed.ctypes.data, offsets.ctypes.data, sizes.ctypes.data, n,
                          out.ctypes.data, stride, nb.ctypes.data, st.ctypes.data, nthreads)
    return out, nb, st


# ---- encoder side -------------------------------------------------------------------------------
def make_cookie(cfg: Config, wrappers=1) -> bytes:
    buf = C.create_string_buffer(64)
    n = lib().ae_write_cookie(C.byref(cfg), wrappers, buf)
    return buf.raw[:n]


def encode_packet(cfg: Config, pcm, opts: PacketOpts = None) -> bytes:
    """pcm: int array [n, num_channels] in OUTPUT channel order -> one packet."""
    opts = opts or PacketOpts.make()
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    assert pcm.ndim == 2 and pcm.shape[1] == cfg.num_channels
    n = pcm.shape[0]
    cap = n * cfg.num_channels * 5 + 1024 + opts.fil_bytes + opts.dse_bytes
    buf = np.zeros(cap, dtype=np.uint8)
    sz = lib().ae_encode_packet(C.byref(cfg), C.byref(opts), pcm.ctypes.data, n, buf.ctypes.data, cap)
    if sz < 0:
        raise ValueError('ae_encode_packet failed')
    return buf[:sz].tobytes()


def encode_stream(cfg: Config, pcm, opts: PacketOpts = None):
    """pcm [n, ch] -> list of packets (frame_length samples each, partial last)."""
    fl = cfg.frame_length
    return [encode_packet(cfg, pcm[s:s + fl], opts) for s in range(0, len(pcm), fl)]


class Writer:
    """Element-level bitstream synthesiser (for shapes the packet encoder does not produce)."""

    def __init__(self, cfg: Config, cap=1 << 20):
        self.cfg = cfg
        self._buf = np.zeros(cap, dtype=np.uint8)
        self._w = _Writer()
        lib().ae_writer_init(C.byref(self._w), self._buf.ctypes.data, cap)

    def bits(self, value, nbits):
        lib().ae_put_bits(C.byref(self._w), int(value) & 0xffffffff, nbits)
        return self

    def align(self):
        lib().ae_byte_align(C.byref(self._w))
        return self

    def element(self, tag, c0, c1=None, *, order=4, coefs=None, den_shift=9, mode=0, pb_factor=4, mix_bits=0,
                mix_res=0, bytes_shifted=0, escape=0, partial=None, instance=0, order_v=None, coefs_v=None,
                mode_v=None, den_shift_v=None, pb_factor_v=None):
        e = Element()
        e.tag, e.instance, e.escape, e.bytes_shifted = tag, instance, escape, bytes_shifted
        c0 = np.ascontiguousarray(c0, dtype=np.int32)
        n = len(c0)
        e.partial = int(n != self.cfg.frame_length) if partial is None else int(partial)
        e.mix_bits, e.mix_res = mix_bits, mix_res
        for ci, (o, cf, md, ds, pf) in enumerate(((order, coefs, mode, den_shift, pb_factor),
                                                   (order if order_v is None else order_v,
                                                    coefs if coefs_v is None else coefs_v,
                                                    mode if mode_v is None else mode_v,
                                                    den_shift if den_shift_v is None else den_shift_v,
                                                    pb_factor if pb_factor_v is None else pb_factor_v))):
            p = e.ch[ci]
            p.mode, p.den_shift, p.pb_factor, p.order = md, ds, pf, o
            if cf is not None:
                for j, v in enumerate(cf[:32]):
                    p.coefs[j] = int(v)
        p1 = None
        if c1 is not None:
            c1 = np.ascontiguousarray(c1, dtype=np.int32)
            p1 = c1.ctypes.data
        rc = lib().ae_encode_element(C.byref(self._w), C.byref(self.cfg), C.byref(e), c0.ctypes.data, p1, n)
        if rc != 0:
            raise ValueError(f'ae_encode_element rc={rc}')
        return self

    def dse(self, count, align=1, instance=0):
        lib().ae_write_dse(C.byref(self._w), instance, align, None, count)
        return self

    def fil(self, count):
        lib().ae_write_fil(C.byref(self._w), count)
        return self

    def end(self):
        lib().ae_write_end(C.byref(self._w))
        return self

    def bytes(self) -> bytes:
        assert not self._w.overflow
        return self._buf[:(self._w.bitpos + 7) // 8].tobytes()


# ---- PCM helpers --------------------------------------------------------------------------------
def pcm_bytes_to_int(pcm: bytes, bit_depth: int, channels: int):
    """interleaved LE PCM -> int64 [n, ch] (20-bit: value as stored, i.e. <<4 in a 24-bit container)."""
    b = np.frombuffer(pcm, dtype=np.uint8)
    bps = {16: 2, 20: 3, 24: 3, 32: 4}[bit_depth]
    b = b.reshape(-1, bps).astype(np.int64)
    v = np.zeros(len(b), dtype=np.int64)
    for k in range(bps):
        v |= b[:, k] << (8 * k)
    sign = 1 << (8 * bps - 1)
    v = (v ^ sign) - sign
    return v.reshape(-1, channels)


def int_to_pcm_bytes(x, bit_depth: int) -> bytes:
    """int [n, ch] -> interleaved LE PCM bytes (20-bit: x << 4 in 3 bytes)."""
    x = np.asarray(x, dtype=np.int64)
    bps = {16: 2, 20: 3, 24: 3, 32: 4}[bit_depth]
    if bit_depth == 20:
        x = x << 4
    flat = x.reshape(-1)
    out = np.zeros((len(flat), bps), dtype=np.uint8)
    for k in range(bps):
        out[:, k] = (flat >> (8 * k)) & 0xff
    return out.tobytes()
