// alac_kernels.cuh -- hand-written sm_100a kernels for the ALAC packet-decode hot path.
//
// Two kernels, both parallel ACROSS packets (packets are independent, decoder.go:79-87):
//
//   alac_decode_kernel  stage 1+2. One thread per packet, one warp = 32 packets. Walks the element
//                       grammar of decodePacketInto (decoder.go:133-207), and per channel runs the
//                       adaptive Golomb-Rice decoder (DynDecomp, golomb.go:148-253) FUSED with the
//                       sign-LMS predictor (UnpcBlock, predictor.go:45-684): residuals never leave
//                       registers. Decoded channel samples are parked as int32 in a lane-interleaved
//                       scratch ([group][slot][sample][32 lanes]) so every warp store is one 128-byte line.
//                       Compressed bytes are pulled with 128-bit loads into a per-lane register
//                       bit reservoir (2 x uint4 deep prefetch).
//   alac_emit_kernel    stage 3. Fully data-parallel un-mix + shift-merge + interleaved little-endian
//                       PCM emit (WriteStereo*/WriteMono*, matrix.go:30-301) through a shared-memory
//                       transpose so global stores are coalesced / 128-bit.
//
// Integer semantics are the Go reference's: wrap-around int32/uint32, shifts >= 32 give 0 / sign
// fill (PTX shl/shr clamp exactly like that), int32 coefficients for orders 4/5/6/8 and int16-wrapping
// coefficients otherwise (predictor.go:107-110 vs :664). Where the reference would panic the packet
// gets ST_REF_PANIC.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace alacb200 {

enum : int32_t {
    ST_OK = 0,
    ST_UNSUPPORTED_ELEMENT = 3,
    ST_INVALID_HEADER = 4,
    ST_INVALID_SHIFT = 5,
    ST_BITSTREAM_OVERRUN = 6,
    ST_SAMPLE_OVERRUN = 7,
    ST_REF_PANIC = 9,
};
enum : int32_t { CTX_SCE = 1, CTX_CPE = 2, CTX_DSE = 3, CTX_FIL = 4 };
enum : int32_t { ENT_MONO = 1, ENT_U = 2, ENT_V = 3 };

struct DevConfig {
    uint32_t frame_length;
    uint32_t bit_depth;
    uint32_t num_channels;
    uint32_t bps;
    uint32_t pb, mb, kb;
};

// What stage 1+2 hands to stage 3 for one decoded element (one "write op" of matrix.go).
struct OpDesc {
    uint32_t n;             // samples this element wrote
    uint32_t shift_bitpos;  // absolute bit position of the shift data inside the packet
    uint8_t kind;           // 1 = WriteMono*, 2 = WriteStereo*
    uint8_t out_chan;       // output channel index (channelLayoutOffsets, decoder.go:55-64)
    uint8_t slot;           // scratch slot of U (V = slot+1)
    uint8_t shift;          // bytesShifted seen by the writer (0 for escape elements)
    uint8_t mix_bits;
    int8_t mix_res;
    uint16_t pad_;
};
struct PacketDesc {
    int32_t status;
    uint32_t n_final;
    uint32_t nops;
    uint32_t pad_;
    OpDesc ops[8];
};

// ---- Go shift semantics: PTX shl/shr clamp the shift amount to 32 ------------------------------
__device__ __forceinline__ uint32_t shl_go(uint32_t x, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
__device__ __forceinline__ uint32_t shr_go(uint32_t x, uint32_t s) {
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
__device__ __forceinline__ int32_t sar_go(int32_t x, uint32_t s) {
    int32_t r;
    asm("shr.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
// (del << chanShift) >> chanShift, predictor.go:69
__device__ __forceinline__ int32_t sext_go(int32_t x, uint32_t cs) { return sar_go((int32_t)shl_go((uint32_t)x, cs), cs); }
// signOfInt, predictor.go:35-39
__device__ __forceinline__ int32_t sign_of(int32_t v) { return (int32_t)((uint32_t)(-v) >> 31) | (v >> 31); }

// ---- slow bit reads straight from global memory (element headers; BitBuffer.Read*, bitbuffer.go:55-96)
struct Packet {
    const uint8_t *p;  // first byte of the packet
    uint32_t size;     // unpadded size; bytes [size, size+4) read as zero (bitbuffer.go:36-51)
};
__device__ __forceinline__ uint32_t pk_byte(const Packet &pk, uint32_t idx) { return idx < pk.size ? (uint32_t)__ldg(pk.p + idx) : 0u; }
// nb <= 16 bits at absolute bit position bp (the 24-bit window of Read)
__device__ __forceinline__ uint32_t pk_bits(const Packet &pk, uint32_t bp, uint32_t nb) {
    uint32_t b = bp >> 3;
    uint32_t w = (pk_byte(pk, b) << 16) | (pk_byte(pk, b + 1) << 8) | pk_byte(pk, b + 2);
    w = (w << (bp & 7)) & 0x00FFFFFFu;
    return shr_go(w, 24u - nb);
}
// read32bit, golomb.go:80 (big endian), zero pad semantics
__device__ __forceinline__ uint32_t pk_be32(const Packet &pk, uint32_t b) {
    return (pk_byte(pk, b) << 24) | (pk_byte(pk, b + 1) << 16) | (pk_byte(pk, b + 2) << 8) | pk_byte(pk, b + 3);
}

// Header cursor: absolute bit position + sticky "the reference would have panicked" flag.
struct Cursor {
    uint32_t bp;
    bool panic;
};
__device__ __forceinline__ uint32_t cur_read(const Packet &pk, Cursor &c, uint32_t nb) {  // Read, needs Pos+3 <= cap
    if ((c.bp >> 3) + 3u > pk.size + 4u) c.panic = true;
    uint32_t v = pk_bits(pk, c.bp, nb);
    c.bp += nb;
    return v;
}
__device__ __forceinline__ uint32_t cur_read_small(const Packet &pk, Cursor &c, uint32_t nb) {  // ReadSmall, Pos+2 <= cap
    if ((c.bp >> 3) + 2u > pk.size + 4u) c.panic = true;
    uint32_t v = pk_bits(pk, c.bp, nb);
    c.bp += nb;
    return v;
}
__device__ __forceinline__ uint32_t cur_read_one(const Packet &pk, Cursor &c) {  // ReadOne, Pos < cap
    if ((c.bp >> 3) >= pk.size + 4u) c.panic = true;
    uint32_t v = pk_bits(pk, c.bp, 1);
    c.bp += 1;
    return v;
}

// ---- register bit reservoir for the hot loops ---------------------------------------------------
// 64-bit window (hi:lo, big-endian order) + one uint4 of queued words + one uint4 in flight.
struct BitRes {
    const uint4 *base;  // 16-byte aligned address at or below the packet start
    uint32_t end_rel;   // packet end, in bytes relative to base
    uint32_t nchunk;    // next chunk index to prefetch
    uint4 cur, pf;
    uint32_t widx;  // next word of cur
    uint32_t hi, lo;
    uint32_t sh;  // bits of hi already consumed (0..31)

    __device__ __forceinline__ uint4 load_chunk(uint32_t c) const {
        uint32_t b0 = c << 4;
        if (b0 >= end_rel) return make_uint4(0, 0, 0, 0);
        uint4 v = __ldg(base + c);
        if (b0 + 16u > end_rel) {  // bytes past the packet end read as zero (the reference's 4-byte pad)
            uint32_t keep = end_rel - b0;  // 1..15 bytes valid
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t lo_b = (uint32_t)i * 4u;
                if (keep <= lo_b) w[i] = 0;
                else if (keep < lo_b + 4u) w[i] &= (1u << ((keep - lo_b) * 8u)) - 1u;
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    }
    __device__ __forceinline__ uint32_t pop() {
        uint32_t w = widx == 0 ? cur.x : widx == 1 ? cur.y : widx == 2 ? cur.z : cur.w;
        widx++;
        if (widx == 4) {
            cur = pf;
            pf = load_chunk(nchunk++);
            widx = 0;
        }
        return __byte_perm(w, 0, 0x0123);
    }
    // position the window at absolute packet bit position bp
    __device__ __forceinline__ void init(const Packet &pk, uint32_t bp) {
        uintptr_t a = (uintptr_t)pk.p;
        uint32_t mis = (uint32_t)(a & 15u);
        base = (const uint4 *)(a - mis);
        end_rel = mis + pk.size;
        uint32_t abp = bp + mis * 8u;
        uint32_t c = abp >> 7;
        cur = load_chunk(c);
        pf = load_chunk(c + 1);
        nchunk = c + 2;
        widx = (abp >> 5) & 3u;
        hi = pop();
        lo = pop();
        sh = abp & 31u;
    }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_l(lo, hi, sh); }
    __device__ __forceinline__ void consume(uint32_t nb) {
        sh += nb;
        while (sh >= 32u) {
            hi = lo;
            lo = pop();
            sh -= 32u;
        }
    }
};

// ---- adaptive Golomb-Rice state (DynDecomp, golomb.go:148-253) --------------------------------------
struct Entropy {
    uint32_t mean, zmode, zrun;
    uint32_t pb, kb, wb;
    uint32_t max_size;  // escape width = chanBits
    uint32_t size8;     // packet size in bits
};

// Decodes the residual of sample index i (0-based) of n. Returns false and sets st on error.
__device__ __forceinline__ bool entropy_next(const Packet &pk, BitRes &br, uint32_t &bp, Entropy &e, uint32_t i,
                                             uint32_t n, int32_t &res, int32_t &st) {
    if (e.zrun > 0) {  // inside a zero run (clear(predCoefs[count:end]), golomb.go:237)
        e.zrun--;
        res = 0;
        return true;
    }
    if (bp >= e.size8) {  // golomb.go:168-170
        st = ST_BITSTREAM_OVERRUN;
        return false;
    }
    uint32_t m = e.mean >> 9;
    uint32_t k = 31u - (uint32_t)__clz((int32_t)(m + 3u));
    k = min(k, e.kb);
    m = shl_go(1u, k) - 1u;
    uint32_t w = br.window();
    uint32_t r = (uint32_t)__clz((int32_t)~w);
    if (r >= 9u) {
        // getStreamBits(input, bitPos+9, maxSize), golomb.go:86-108
        uint32_t bo = bp + 9u;
        uint32_t byte_off = bo >> 3;
        if (byte_off > pk.size) { st = ST_REF_PANIC; return false; }
        uint32_t load1 = pk_be32(pk, byte_off);
        uint32_t nb = e.max_size;
        if (nb + (bo & 7u) > 32u) {
            if (byte_off >= pk.size) { st = ST_REF_PANIC; return false; }
            uint32_t v = load1 << (bo & 7u);
            uint32_t load2 = pk_byte(pk, byte_off + 4u);
            load2 = shr_go(load2, 8u - (nb + (bo & 7u) - 32u));
            v = shr_go(v, 32u - nb);
            r = v | load2;
        } else {
            uint32_t v = shr_go(load1, 32u - nb - (bo & 7u));
            if (nb < 32u) v &= shl_go(1u, nb) - 1u;
            r = v;
        }
        bp += 9u + nb;
        br.consume(9u + nb);
    } else {
        uint32_t nb = r + 1u;
        if (k != 1u) {
            uint32_t s = w << nb;
            uint32_t v = shr_go(s, 32u - k);
            if (v >= 2u) {
                r = r * m + v - 1u;
                nb += k;
            } else {
                r = r * m;
                nb += k - 1u;
            }
        }
        bp += nb;
        br.consume(nb);
    }
    uint32_t nd = r + e.zmode;
    int32_t mag = (int32_t)((nd + 1u) >> 1);
    res = (nd & 1u) ? -mag : mag;
    e.mean = e.pb * nd + e.mean - ((e.pb * e.mean) >> 9);
    if (r > 0xffffu) e.mean = 0xffffu;
    e.zmode = 0;
    if ((e.mean << 2) < 512u && i + 1u < n) {
        // zero run: dynGet, golomb.go:112-144
        e.zmode = 1;
        int32_t k32 = __clz((int32_t)e.mean) - 24 + (int32_t)((e.mean + 16u) >> 6);
        if (k32 < 0) k32 = 0;
        uint32_t mz = (shl_go(1u, (uint32_t)k32) - 1u) & e.wb;
        if ((bp >> 3) > pk.size) { st = ST_REF_PANIC; return false; }
        uint32_t w2 = br.window();
        uint32_t pre = (uint32_t)__clz((int32_t)~w2);
        uint32_t run, nb;
        if (pre >= 9u) {
            run = (w2 << 9) >> 16;
            nb = 25u;
        } else {
            nb = pre + 1u;
            uint32_t s = shl_go(w2, nb);
            uint32_t val = shr_go(s, 32u - (uint32_t)k32);
            nb += (uint32_t)k32;
            if (val < 2u) {
                run = pre * mz;
                nb -= 1u;
            } else {
                run = pre * mz + val - 1u;
            }
        }
        bp += nb;
        br.consume(nb);
        if (i + 1u + run > n) {  // golomb.go:232-234
            st = ST_SAMPLE_OVERRUN;
            return false;
        }
        e.zrun = run;
        if (run >= 65535u) e.zmode = 0;
        e.mean = 0;
    }
    return true;
}

// Per-channel header, decoder.go:275-286.
struct ChanHdr {
    uint32_t mode, den_shift, pb_factor, num;
    int16_t coefs[32];
};
__device__ __forceinline__ void read_chan_hdr(const Packet &pk, Cursor &c, ChanHdr &h) {
    uint32_t hb = cur_read(pk, c, 8);
    h.mode = hb >> 4;
    h.den_shift = hb & 0xfu;
    hb = cur_read(pk, c, 8);
    h.pb_factor = hb >> 5;
    h.num = hb & 0x1fu;
#pragma unroll 1
    for (uint32_t i = 0; i < 32; i++) h.coefs[i] = (i < h.num) ? (int16_t)cur_read(pk, c, 16) : (int16_t)0;
}

// Order-31 pre-pass state (mode != 0: UnpcBlock(pred, pred, n, nil, 31, chanBits, 0), decoder.go:306-308)
struct Delta {
    bool on;
    int32_t prev;
};
__device__ __forceinline__ int32_t delta_step(Delta &d, int32_t r, uint32_t i, uint32_t cs) {
    if (!d.on) return r;
    d.prev = (i == 0) ? r : sext_go(r + d.prev, cs);
    return d.prev;
}

// ---- fused entropy + predictor loops ---------------------------------------------------------------
// T-tap register predictor for orders with int32 coefficient semantics (unpcBlock4/5/6/8,
// predictor.go:99-618). T=6 serves orders 4,5,6 (taps >= order are masked); T=8 serves order 8.
template <int T>
__device__ __forceinline__ int32_t channel_fixed(const Packet &pk, BitRes &br, uint32_t &bp, Entropy &e,
                                                 const ChanHdr &hd, uint32_t n, uint32_t chan_bits,
                                                 int32_t *__restrict__ dst) {
    const uint32_t cs = 32u - chan_bits;
    const uint32_t den = hd.den_shift;
    const int32_t den_half = den > 0 ? (int32_t)(1u << (den - 1)) : 0;
    const int32_t order = (T == 8) ? 8 : (int32_t)hd.num;
    int32_t c[T], h[T + 1], wgt[T];
    bool in_tap[T];
#pragma unroll
    for (int j = 0; j < T; j++) {
        c[j] = (j < order) ? (int32_t)hd.coefs[j] : 0;
        wgt[j] = (j < order) ? order - j : 0;
        in_tap[j] = j < order;
    }
#pragma unroll
    for (int j = 0; j <= T; j++) h[j] = 0;
    Delta dl{hd.mode != 0, 0};
    int32_t st = ST_OK;
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        int32_t r;
        if (!entropy_next(pk, br, bp, e, i, n, r, st)) return st;
        r = delta_step(dl, r, i, cs);
        int32_t top;
        if (T == 8) top = h[8];
        else top = (order == 4) ? h[4] : (order == 5) ? h[5] : h[6];
        int32_t d[T];
        int32_t sum = den_half;
#pragma unroll
        for (int j = 0; j < T; j++) {
            d[j] = top - h[j];
            sum -= c[j] * d[j];
        }
        const int32_t fir = sext_go(r + top + (sum >> den), cs);
        const int32_t warm = (i == 0) ? r : sext_go(r + h[0], cs);
        const bool is_fir = (int32_t)i > order;
        const int32_t x = is_fir ? fir : warm;
        // sign-LMS adaptation on the residual's sign
        bool alive = is_fir && (r != 0);
        const int32_t smask = r >> 31;      // 0 / -1
        const int32_t thr = 1 + smask;      // continue while (D ^ smask) >= thr  <=>  D > 0 (r>0) / D < 0 (r<0)
        int32_t D = r;
#pragma unroll
        for (int j = T - 1; j >= 0; j--) {
            const int32_t sg = sign_of(d[j]);
            const int32_t sgn = (sg ^ smask) - smask;  // sg for r>0, -sg for r<0
            const bool act = alive && in_tap[j];
            if (act) c[j] -= sgn;
            if (j > 0) {
                const int32_t term = (sgn * d[j]) >> den;
                D -= wgt[j] * term;  // wgt == 0 for masked taps
                alive = alive && (!in_tap[j] || ((D ^ smask) >= thr));
            }
        }
#pragma unroll
        for (int j = T; j > 0; j--) h[j] = h[j - 1];
        h[0] = x;
        dst[(size_t)i * 32u] = x;
    }
    return ST_OK;
}

// Orders 0 (copy) and 31 (first-order delta), predictor.go:55-72.
__device__ __forceinline__ int32_t channel_simple(const Packet &pk, BitRes &br, uint32_t &bp, Entropy &e,
                                                  const ChanHdr &hd, uint32_t n, uint32_t chan_bits,
                                                  int32_t *__restrict__ dst) {
    const uint32_t cs = 32u - chan_bits;
    Delta dl{hd.mode != 0, 0};
    const bool acc = hd.num == 31;
    int32_t prev = 0, st = ST_OK;
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        int32_t r;
        if (!entropy_next(pk, br, bp, e, i, n, r, st)) return st;
        r = delta_step(dl, r, i, cs);
        int32_t x = r;
        if (acc && i > 0) x = sext_go(r + prev, cs);
        prev = x;
        dst[(size_t)i * 32u] = x;
    }
    return ST_OK;
}

// Every other order: unpcBlockGeneral, predictor.go:623-684 (int16 coefficients, wrap on update).
__device__ __noinline__ int32_t channel_general(const Packet &pk, BitRes &br, uint32_t &bp, Entropy &e, ChanHdr &hd,
                                                uint32_t n, uint32_t chan_bits, int32_t *__restrict__ dst) {
    const uint32_t cs = 32u - chan_bits;
    const uint32_t den = hd.den_shift;
    const int32_t den_half = den > 0 ? (int32_t)(1u << (den - 1)) : 0;
    const int32_t order = (int32_t)hd.num;
    int32_t hist[32];  // ring: out[i] at hist[i & 31]
    Delta dl{hd.mode != 0, 0};
    int32_t st = ST_OK;
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        int32_t r;
        if (!entropy_next(pk, br, bp, e, i, n, r, st)) return st;
        r = delta_step(dl, r, i, cs);
        int32_t x;
        if (i == 0) x = r;
        else if ((int32_t)i <= order) x = sext_go(r + hist[(i - 1) & 31u], cs);
        else {
            const int32_t top = hist[(i - (uint32_t)order - 1u) & 31u];
            int32_t sum1 = 0;
#pragma unroll 1
            for (int32_t k = 0; k < order; k++) sum1 += (int32_t)hd.coefs[k] * (hist[(i - 1u - (uint32_t)k) & 31u] - top);
            x = sext_go(r + top + ((sum1 + den_half) >> den), cs);
            int32_t del0 = r;
            if (r > 0) {
#pragma unroll 1
                for (int32_t k = order - 1; k >= 0; k--) {
                    int32_t dd = top - hist[(i - 1u - (uint32_t)k) & 31u];
                    int32_t sgn = sign_of(dd);
                    hd.coefs[k] = (int16_t)(hd.coefs[k] - (int16_t)sgn);
                    del0 -= (order - k) * ((sgn * dd) >> den);
                    if (del0 <= 0) break;
                }
            } else if (r < 0) {
#pragma unroll 1
                for (int32_t k = order - 1; k >= 0; k--) {
                    int32_t dd = top - hist[(i - 1u - (uint32_t)k) & 31u];
                    int32_t sgn = sign_of(dd);
                    hd.coefs[k] = (int16_t)(hd.coefs[k] + (int16_t)sgn);
                    del0 -= (order - k) * ((-sgn * dd) >> den);
                    if (del0 >= 0) break;
                }
            }
        }
        hist[i & 31u] = x;
        dst[(size_t)i * 32u] = x;
    }
    return ST_OK;
}

// One compressed channel: SetAGParams + DynDecomp + UnpcBlock (decoder.go:296-312).
__device__ __forceinline__ int32_t decode_channel(const Packet &pk, const DevConfig &cfg, BitRes &br, uint32_t &bp,
                                                  ChanHdr &hd, uint32_t n, uint32_t chan_bits,
                                                  int32_t *__restrict__ dst) {
    // DynDecomp entry: input := Buf[Pos:] (golomb.go:149) and the first read32bit when Pos > Size
    const uint32_t pos = bp >> 3;
    if (pos > pk.size + 4u) return ST_REF_PANIC;
    if (n > 0 && pos > pk.size) return ST_REF_PANIC;
    Entropy e;
    e.mean = cfg.mb;
    e.zmode = 0;
    e.zrun = 0;
    e.pb = (cfg.pb * hd.pb_factor) / 4u;
    e.kb = cfg.kb;
    e.wb = shl_go(1u, cfg.kb) - 1u;
    e.max_size = chan_bits;
    e.size8 = pk.size * 8u;
    const uint32_t ord = hd.num;
    int32_t st;
    if (ord >= 4 && ord <= 6) st = channel_fixed<6>(pk, br, bp, e, hd, n, chan_bits, dst);
    else if (ord == 8) st = channel_fixed<8>(pk, br, bp, e, hd, n, chan_bits, dst);
    else if (ord == 0 || ord == 31) st = channel_simple(pk, br, bp, e, hd, n, chan_bits, dst);
    else st = channel_general(pk, br, bp, e, hd, n, chan_bits, dst);
    // the warm-up indexes [1..numActive] of frame_length-long slices (predictor.go:76-79); it runs after
    // DynDecomp, so an entropy error wins
    if (st == ST_OK && ord != 0 && ord != 31 && ord >= cfg.frame_length) st = ST_REF_PANIC;
    return st;
}

// decodeSCEEscape / decodeCPEEscape, decoder.go:326-345, :504-535
__device__ __forceinline__ int32_t escape_sample(BitRes &br, uint32_t chan_bits) {
    const uint32_t shift = 32u - chan_bits;
    if (chan_bits <= 16u) {
        int32_t val = (int32_t)shr_go(br.window(), 32u - chan_bits);
        br.consume(chan_bits);
        return sext_go(val, shift);
    }
    const uint32_t extra = chan_bits - 16u;
    int32_t val = (int32_t)(br.window() >> 16);
    br.consume(16u);
    val = sar_go((int32_t)((uint32_t)val << 16), shift);
    int32_t lo = (int32_t)shr_go(br.window(), 32u - extra);
    br.consume(extra);
    return val | lo;
}

__constant__ int8_t k_layout[8][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0}, {2, 0, 1, 0, 0, 0, 0, 0},
                                      {2, 0, 1, 3, 0, 0, 0, 0}, {2, 0, 1, 3, 4, 0, 0, 0}, {2, 0, 1, 4, 5, 3, 0, 0},
                                      {2, 0, 1, 4, 5, 6, 3, 0}, {2, 6, 7, 0, 1, 4, 5, 3}};

// decodeSCE / decodeCPE (decoder.go:210-265, :348-414) minus the writer, which becomes an OpDesc.
__device__ __forceinline__ int32_t decode_element(const Packet &pk, const DevConfig &cfg, Cursor &cur, bool stereo,
                                                  uint32_t chan_idx, uint32_t &ns, int32_t *__restrict__ scratch_lane,
                                                  PacketDesc *__restrict__ desc, uint32_t &nops) {
    (void)cur_read_small(pk, cur, 4);
    uint32_t unused = cur_read(pk, cur, 12);
    if (cur.panic) return ST_REF_PANIC;
    if (unused != 0) return ST_INVALID_HEADER;
    uint32_t hb = cur_read(pk, cur, 4);
    if (cur.panic) return ST_REF_PANIC;
    const uint32_t partial = hb >> 3;
    uint32_t shift = (hb >> 1) & 3u;
    if (shift == 3u) return ST_INVALID_SHIFT;
    const uint32_t escape = hb & 1u;
    uint32_t chan_bits = cfg.bit_depth - shift * 8u + (stereo ? 1u : 0u);
    uint32_t n = ns;
    if (partial) {
        n = cur_read(pk, cur, 16) << 16;
        n |= cur_read(pk, cur, 16);
        if (cur.panic) return ST_REF_PANIC;
    }
    // numSamples > frame_length: every later path re-slices a frame_length buffer and panics
    // (golomb.go:155, decoder.go:328, :506)
    const bool n_too_big = n > cfg.frame_length;
    uint32_t mix_bits = 0;
    int32_t mix_res = 0;
    uint32_t shift_bitpos = 0;
    int32_t *dst_u = scratch_lane + (size_t)chan_idx * cfg.frame_length * 32u;
    int32_t *dst_v = dst_u + (size_t)cfg.frame_length * 32u;
    if (!escape) {
        mix_bits = cur_read(pk, cur, 8);
        mix_res = (int32_t)(int8_t)cur_read(pk, cur, 8);
        ChanHdr hu, hv;
        read_chan_hdr(pk, cur, hu);
        if (stereo) read_chan_hdr(pk, cur, hv);
        if (cur.panic) return ST_REF_PANIC;
        shift_bitpos = cur.bp;
        if (shift != 0) cur.bp += shift * 8u * n * (stereo ? 2u : 1u);
        if (n_too_big) return ST_REF_PANIC;
        BitRes br;
        br.init(pk, cur.bp);
        int32_t st = decode_channel(pk, cfg, br, cur.bp, hu, n, chan_bits, dst_u);
        if (st != ST_OK) return st == ST_REF_PANIC ? st : (st | ((stereo ? ENT_U : ENT_MONO) << 12));
        if (stereo) {
            st = decode_channel(pk, cfg, br, cur.bp, hv, n, chan_bits, dst_v);
            if (st != ST_OK) return st == ST_REF_PANIC ? st : (st | (ENT_V << 12));
        }
    } else {
        if (stereo) chan_bits = cfg.bit_depth;  // decoder.go:388
        if (n_too_big) return ST_REF_PANIC;
        if (n > 0) {
            // every Read needs Pos+3 <= cap; positions only grow, so checking the last one is enough
            const uint32_t per = chan_bits * (stereo ? 2u : 1u);
            const uint32_t last_nb = chan_bits <= 16u ? chan_bits : chan_bits - 16u;
            const uint32_t bp_last = cur.bp + n * per - last_nb;
            if ((bp_last >> 3) + 3u > pk.size + 4u) return ST_REF_PANIC;
            BitRes br;
            br.init(pk, cur.bp);
#pragma unroll 1
            for (uint32_t i = 0; i < n; i++) {
                dst_u[(size_t)i * 32u] = escape_sample(br, chan_bits);
                if (stereo) dst_v[(size_t)i * 32u] = escape_sample(br, chan_bits);
            }
            cur.bp += n * per;
        }
        shift = 0;
    }
    const uint32_t out_chan = (uint32_t)k_layout[cfg.num_channels - 1][chan_idx];
    // dst := out[off:off+W:off+W] past cap(out), matrix.go:44 (only when a pair is mapped onto the last channel)
    if (n > 0 && out_chan + (stereo ? 2u : 1u) > cfg.num_channels && n == cfg.frame_length) return ST_REF_PANIC;
    OpDesc op;
    op.n = n;
    op.shift_bitpos = shift_bitpos;
    op.kind = stereo ? 2 : 1;
    op.out_chan = (uint8_t)out_chan;
    op.slot = (uint8_t)chan_idx;
    op.shift = (uint8_t)shift;
    op.mix_bits = (uint8_t)mix_bits;
    op.mix_res = (int8_t)mix_res;
    op.pad_ = 0;
    desc->ops[nops++] = op;
    ns = n;
    return ST_OK;
}

// decodePacketInto, decoder.go:133-207.
__global__ void __launch_bounds__(32) alac_decode_kernel(const uint8_t *__restrict__ packed,
                                                         const uint64_t *__restrict__ offsets,
                                                         const uint32_t *__restrict__ sizes, uint32_t npackets,
                                                         DevConfig cfg, int32_t *__restrict__ scratch,
                                                         PacketDesc *__restrict__ descs,
                                                         uint32_t *__restrict__ out_bytes,
                                                         int32_t *__restrict__ status) {
    const uint32_t pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= npackets) return;
    const uint32_t group = pidx >> 5, lane = pidx & 31u;
    Packet pk{packed + offsets[pidx], sizes[pidx]};
    PacketDesc *desc = descs + pidx;
    int32_t *scratch_lane = scratch + (size_t)group * cfg.num_channels * cfg.frame_length * 32u + lane;

    Cursor cur{0, false};
    uint32_t ns = cfg.frame_length, chan_idx = 0, nops = 0;
    int32_t st = ST_OK;
    if (pk.size > 0x0FFFFFFFu) st = ST_REF_PANIC;  // bit positions are 32-bit here
#pragma unroll 1
    while (st == ST_OK) {
        if ((cur.bp >> 3) >= pk.size) {  // PastEnd, decoder.go:143-145
            st = ST_BITSTREAM_OVERRUN;
            break;
        }
        const uint32_t tag = cur_read_small(pk, cur, 3);
        if (cur.panic) { st = ST_REF_PANIC; break; }
        if (tag == 0 || tag == 3) {
            st = decode_element(pk, cfg, cur, false, chan_idx, ns, scratch_lane, desc, nops);
            if (st != ST_OK) { st |= CTX_SCE << 8; break; }
            chan_idx += 1;
        } else if (tag == 1) {
            if (chan_idx + 2 > cfg.num_channels) break;  // decoder.go:163-165
            st = decode_element(pk, cfg, cur, true, chan_idx, ns, scratch_lane, desc, nops);
            if (st != ST_OK) { st |= CTX_CPE << 8; break; }
            chan_idx += 2;
        } else if (tag == 2 || tag == 5) {
            st = ST_UNSUPPORTED_ELEMENT;
            break;
        } else if (tag == 4) {  // skipDSE, decoder.go:553-574
            (void)cur_read_small(pk, cur, 4);
            const uint32_t align = cur_read_one(pk, cur);
            uint32_t count = cur_read_small(pk, cur, 8);
            if (count == 255) count += cur_read_small(pk, cur, 8);
            if (cur.panic) { st = ST_REF_PANIC | (CTX_DSE << 8); break; }
            if (align) cur.bp = (cur.bp + 7u) & ~7u;
            cur.bp += count * 8u;
            if ((cur.bp >> 3) >= pk.size) { st = ST_BITSTREAM_OVERRUN | (CTX_DSE << 8); break; }
        } else if (tag == 6) {  // skipFIL, decoder.go:538-551
            uint32_t count = cur_read_small(pk, cur, 4);
            if (count == 15) count += cur_read_small(pk, cur, 8) - 1u;
            if (cur.panic) { st = ST_REF_PANIC | (CTX_FIL << 8); break; }
            cur.bp += count * 8u;
            if ((cur.bp >> 3) >= pk.size) { st = ST_BITSTREAM_OVERRUN | (CTX_FIL << 8); break; }
        } else {  // END, decoder.go:192-195
            break;
        }
        if (chan_idx >= cfg.num_channels) break;  // decoder.go:200-202
    }
    desc->status = st;
    desc->n_final = ns;
    desc->nops = nops;
    status[pidx] = st;
    out_bytes[pidx] = st == ST_OK ? ns * cfg.num_channels * cfg.bps : 0u;
}

// ---- stage 3 ---------------------------------------------------------------------------------------
constexpr int EMIT_TILE = 64;       // frames per block
constexpr int EMIT_THREADS = 256;   // 8 warps; lane = packet of the group, warp = frame phase

__device__ __forceinline__ void tile_put(uint8_t *row, int32_t lo_byte, int32_t hi_byte, int32_t off, int32_t v, int bps) {
    // store the bps little-endian bytes of v at packet byte offset `off`, clipped to the tile [lo_byte, hi_byte)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < bps) {
            int32_t o = off + k;
            if (o >= lo_byte && o < hi_byte) row[o - lo_byte] = (uint8_t)((uint32_t)v >> (8 * k));
        }
    }
}

// WriteStereo16/20/24/32 + WriteMono16/20/24/32 (matrix.go:30-301), ops replayed in element order so
// later elements overwrite earlier ones exactly as the sequential reference does.
__global__ void __launch_bounds__(EMIT_THREADS) alac_emit_kernel(const uint8_t *__restrict__ packed,
                                                                 const uint64_t *__restrict__ offsets,
                                                                 const uint32_t *__restrict__ sizes,
                                                                 uint32_t npackets, DevConfig cfg,
                                                                 const int32_t *__restrict__ scratch,
                                                                 const PacketDesc *__restrict__ descs,
                                                                 uint8_t *__restrict__ pcm_out, uint64_t out_stride,
                                                                 uint32_t tiles_per_packet) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_nfinal[32];
    __shared__ int32_t s_status[32];
    const uint32_t group = blockIdx.x / tiles_per_packet;
    const uint32_t tile = blockIdx.x % tiles_per_packet;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t fb = cfg.num_channels * cfg.bps;          // bytes per frame
    const uint32_t row_words = (EMIT_TILE * fb) / 4u + 1u;   // odd => conflict-free lane-per-row access
    const int32_t s0 = (int32_t)(tile * EMIT_TILE);
    const int32_t lo_byte = s0 * (int32_t)fb, hi_byte = (s0 + EMIT_TILE) * (int32_t)fb;
    const uint32_t pidx = group * 32u + lane;
    const bool valid = pidx < npackets;

    uint32_t *sw = reinterpret_cast<uint32_t *>(smem);
    for (uint32_t i = threadIdx.x; i < row_words * 32u; i += EMIT_THREADS) sw[i] = 0;  // fresh make(), decoder.go:120
    PacketDesc const *desc = descs + pidx;
    int32_t st = ST_REF_PANIC;
    uint32_t nops = 0;
    if (valid) {
        st = desc->status;
        nops = st == ST_OK ? desc->nops : 0;
        if (warp == 0) {
            s_nfinal[lane] = desc->n_final;
            s_status[lane] = st;
        }
    } else if (warp == 0) {
        s_nfinal[lane] = 0;
        s_status[lane] = ST_REF_PANIC;
    }
    __syncthreads();

    uint8_t *row = smem + (size_t)lane * row_words * 4u;
    const int32_t *sbase = scratch + (size_t)group * cfg.num_channels * cfg.frame_length * 32u + lane;
    Packet pk{nullptr, 0};
    if (valid) pk = Packet{packed + offsets[pidx], sizes[pidx]};
    const int bps = (int)cfg.bps;
    const bool depth20 = cfg.bit_depth == 20;
    const bool merges_shift = cfg.bit_depth == 24 || cfg.bit_depth == 32;  // 16/20-bit writers ignore the shift buffer

#pragma unroll 1
    for (uint32_t e = 0; e < 8; e++) {
        if (e < nops) {
            const OpDesc op = desc->ops[e];
            const bool stereo = op.kind == 2;
            const int32_t *su = sbase + (size_t)op.slot * cfg.frame_length * 32u;
            const int32_t *sv = su + (size_t)cfg.frame_length * 32u;
            const uint32_t sb = (merges_shift ? (uint32_t)op.shift : 0u) * 8u;
            const int32_t mix_res = op.mix_res;
            const uint32_t mix_bits = op.mix_bits;
            // a pair mapped onto the last channel spills R into the next frame (matrix.go:44-48)
            const bool spills = (uint32_t)op.out_chan + (stereo ? 2u : 1u) > cfg.num_channels;
            const int32_t first = spills ? s0 - 1 : s0;
            for (int32_t i = first + (int32_t)warp; i < s0 + EMIT_TILE; i += EMIT_THREADS / 32) {
                if (i < 0 || (uint32_t)i >= op.n) continue;
                int32_t left = su[(size_t)i * 32u], right = 0;
                if (stereo) {
                    right = sv[(size_t)i * 32u];
                    if (mix_res != 0) {  // matrix.go:40-41
                        const int32_t v = right;
                        left = left + v - sar_go(mix_res * v, mix_bits);
                        right = left - v;
                    }
                }
                if (depth20) {
                    left = (int32_t)((uint32_t)left << 4);
                    right = (int32_t)((uint32_t)right << 4);
                }
                if (sb) {  // shift buffer merge, matrix.go:132-135, :270-272
                    const uint32_t idx = stereo ? (uint32_t)i * 2u : (uint32_t)i;
                    left = (int32_t)shl_go((uint32_t)left, sb) | (int32_t)pk_bits(pk, op.shift_bitpos + idx * sb, sb);
                    if (stereo)
                        right = (int32_t)shl_go((uint32_t)right, sb) |
                                (int32_t)pk_bits(pk, op.shift_bitpos + (idx + 1u) * sb, sb);
                }
                const int32_t off = i * (int32_t)fb + (int32_t)op.out_chan * bps;
                tile_put(row, lo_byte, hi_byte, off, left, bps);
                if (stereo) tile_put(row, lo_byte, hi_byte, off + bps, right, bps);
            }
        }
        __syncthreads();
    }

    // flush: one warp per packet row, coalesced. The whole tile is written so the packet's slot in
    // pcm_out is fully defined: bytes past output[:n] (decoder.go:127) and failed packets read as zero.
    const uint32_t tile_frames = min((uint32_t)EMIT_TILE, cfg.frame_length - (uint32_t)s0);
    const uint32_t limit = tile_frames * fb;  // multiple of 4; multiple of 16 for a full tile
    for (uint32_t r = warp; r < 32u; r += EMIT_THREADS / 32) {
        const uint32_t p = group * 32u + r;
        if (p >= npackets) continue;
        int64_t nvalid = s_status[r] == ST_OK ? (int64_t)s_nfinal[r] * fb - lo_byte : 0;
        nvalid = nvalid < 0 ? 0 : (nvalid > (int64_t)limit ? (int64_t)limit : nvalid);
        uint32_t *src = sw + (size_t)r * row_words;
        uint8_t *srcb = reinterpret_cast<uint8_t *>(src);
        for (uint32_t b = (uint32_t)nvalid + lane; b < limit; b += 32u) srcb[b] = 0;
        __syncwarp();
        uint8_t *dst = pcm_out + (size_t)p * out_stride + (size_t)lo_byte;
        const uint32_t nvec = limit >> 4;
        if ((((uintptr_t)dst) & 15u) == 0) {
            for (uint32_t v = lane; v < nvec; v += 32u) {
                uint4 q = make_uint4(src[v * 4u], src[v * 4u + 1u], src[v * 4u + 2u], src[v * 4u + 3u]);
                reinterpret_cast<uint4 *>(dst)[v] = q;
            }
        } else {
            for (uint32_t w = lane; w < nvec * 4u; w += 32u) reinterpret_cast<uint32_t *>(dst)[w] = src[w];
        }
        for (uint32_t w = nvec * 4u + lane; w < (limit >> 2); w += 32u) reinterpret_cast<uint32_t *>(dst)[w] = src[w];
    }
}

}  // namespace alacb200
