/*
 * alac_oracle.c -- TEST INFRASTRUCTURE ONLY (see alac_oracle.h for scope and pinning).
 *
 * Plain-C restatement of the Go reference's ALAC packet decoder with Go integer semantics made
 * explicit: shifts by >= 32 yield 0 / sign fill, signed arithmetic wraps (build with -fwrapv),
 * and every place where the Go runtime would panic on an out-of-range index or slice is modelled
 * as AO_ERR_REF_PANIC assuming a FRESH PacketDecoder (BitBuffer cap == len == size+4,
 * bitbuffer.go:36-51; scratch slices of cap frame_length, decoder.go:105-108).
 */
#include "alac_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- Go shift semantics ------------------------------------------- */
static inline uint32_t shl_u(uint32_t x, uint32_t s) { return s >= 32 ? 0u : x << s; }
static inline uint32_t shr_u(uint32_t x, uint32_t s) { return s >= 32 ? 0u : x >> s; }
static inline int32_t shl_s(int32_t x, uint32_t s) { return s >= 32 ? 0 : (int32_t)((uint32_t)x << s); }
static inline int32_t sar_s(int32_t x, uint32_t s) { return s >= 32 ? (x < 0 ? -1 : 0) : x >> s; }
/* (del << chanShift) >> chanShift, predictor.go:69, :78, :134 */
static inline int32_t sext_go(int32_t x, uint32_t chan_shift) { return sar_s(shl_s(x, chan_shift), chan_shift); }

int ao_bytes_per_sample(uint8_t depth) { /* internal/alac/format.go:23-34 */
    switch (depth) {
    case 16: return 2;
    case 20:
    case 24: return 3;
    case 32: return 4;
    default: return 0;
    }
}

/* ---- config.go:47-81 ------------------------------------------------------------------------ */
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int32_t ao_parse_cookie(const uint8_t *cookie, size_t len, ao_config *out) {
    const uint8_t *d = cookie;
    memset(out, 0, sizeof(*out));
    if (len >= 12 && d[4] == 'f' && d[5] == 'r' && d[6] == 'm' && d[7] == 'a') { d += 12; len -= 12; } /* :50-52 */
    if (len >= 12 && d[4] == 'a' && d[5] == 'l' && d[6] == 'a' && d[7] == 'c') { d += 12; len -= 12; } /* :56-58 */
    if (len < 24) return AO_ERR_INVALID_COOKIE;                                                         /* :60-62 */
    if (d[4] > 0) return AO_ERR_UNSUPPORTED_VERSION;                                                    /* :64-67 */
    out->frame_length = be32(d);
    out->bit_depth = d[5];
    out->pb = d[6];
    out->mb = d[7];
    out->kb = d[8];
    out->num_channels = d[9];
    out->max_run = (uint16_t)((d[10] << 8) | d[11]);
    out->max_frame_bytes = be32(d + 12);
    out->avg_bit_rate = be32(d + 16);
    out->sample_rate = be32(d + 20);
    return AO_OK;
}

int32_t ao_check_config(const ao_config *cfg) {
    if (ao_bytes_per_sample(cfg->bit_depth) == 0) return AO_ERR_BIT_DEPTH; /* decoder.go:91-93 */
    if (cfg->num_channels < 1 || cfg->num_channels > 8) return AO_ERR_UNSUPPORTED_CONFIG;
    if (cfg->frame_length < 1 || cfg->frame_length > 65536) return AO_ERR_UNSUPPORTED_CONFIG;
    return AO_OK;
}

/* ---- bitbuffer.go:25-123 -------------------------------------------------------------------- */
typedef struct {
    uint8_t *buf;   /* padded copy, cap == len == size+4 */
    int64_t cap;
    int64_t pos;
    uint32_t bitidx;
    int64_t size;
    int panic; /* sticky: a Go runtime panic happened */
} bitbuf;

static inline uint32_t bb_read(bitbuf *b, uint32_t nbits) { /* Read, bitbuffer.go:55-66 (nbits <= 16) */
    if (b->pos + 3 > b->cap) { b->panic = 1; return 0; }
    const uint8_t *w = b->buf + b->pos;
    uint32_t r = ((uint32_t)w[0] << 16) | ((uint32_t)w[1] << 8) | w[2];
    r = (r << b->bitidx) & 0x00FFFFFFu;
    r = shr_u(r, 24u - nbits);
    b->bitidx += nbits;
    b->pos += (int64_t)(b->bitidx >> 3);
    b->bitidx &= 7;
    return r;
}
static inline uint8_t bb_read_small(bitbuf *b, uint32_t nbits) { /* ReadSmall, :70-81 (nbits <= 8) */
    if (b->pos + 2 > b->cap) { b->panic = 1; return 0; }
    const uint8_t *w = b->buf + b->pos;
    uint16_t r = (uint16_t)(((uint16_t)w[0] << 8) | w[1]);
    r = (uint16_t)(r << b->bitidx);
    r = (uint16_t)(r >> (16u - nbits));
    b->bitidx += nbits;
    b->pos += (int64_t)(b->bitidx >> 3);
    b->bitidx &= 7;
    return (uint8_t)r;
}
static inline uint8_t bb_read_one(bitbuf *b) { /* ReadOne, :84-91 */
    if (b->pos >= b->cap) { b->panic = 1; return 0; }
    uint8_t r = (uint8_t)((b->buf[b->pos] >> (7 - b->bitidx)) & 1);
    b->bitidx++;
    b->pos += (int64_t)(b->bitidx >> 3);
    b->bitidx &= 7;
    return r;
}
static inline void bb_advance(bitbuf *b, uint32_t nbits) { /* Advance, :99-103 (uint32 wrap kept) */
    b->bitidx += nbits;
    b->pos += (int64_t)(b->bitidx >> 3);
    b->bitidx &= 7;
}
static inline void bb_byte_align(bitbuf *b) { /* :106-112 */
    if (b->bitidx == 0) return;
    bb_advance(b, 8 - b->bitidx);
}
static inline int bb_past_end(const bitbuf *b) { return b->pos >= b->size; } /* :115-117 */

/* ---- golomb.go ------------------------------------------------------------------------------ */
typedef struct { uint32_t mb0, pb, kb, wb; } agparams; /* the fields DynDecomp reads, golomb.go:44-65 */

static inline void set_ag_params(agparams *p, uint32_t mean_base, uint32_t part_bound, uint32_t kbase) {
    p->mb0 = mean_base;
    p->pb = part_bound;
    p->kb = kbase;
    p->wb = shl_u(1u, kbase) - 1u; /* (1 << kBase) - 1, golomb.go:60 */
}
static inline int32_t lead(int32_t m) { return m == 0 ? 32 : (int32_t)__builtin_clz((uint32_t)m); } /* :69 */
static inline int32_t lg3a(int32_t x) { return 31 - lead(x + 3); }                                 /* :73 */
static inline uint32_t rd32(const uint8_t *in, int64_t off) { return be32(in + off); }            /* :80 */

/* DynDecomp, golomb.go:148-253. `cap_samples` is cap(predCoefs) == frame_length. */
static int32_t dyn_decomp(const agparams *p, bitbuf *bb, int32_t *pc, int64_t n, int64_t cap_samples,
                          uint32_t max_size) {
    if (bb->pos > bb->cap) return AO_ERR_REF_PANIC; /* bitBuf.Buf[bitBuf.Pos:], :149 */
    const uint8_t *in = bb->buf + bb->pos;
    const int64_t avail = bb->size - bb->pos; /* read32bit(in, off) panics iff off > avail */
    const uint32_t start_pos = bb->bitidx;
    const uint32_t max_pos = (uint32_t)avail * 8u; /* :152 (wraps when Pos > Size) */
    uint32_t bit_pos = start_pos;
    if (n > cap_samples) return AO_ERR_REF_PANIC; /* predCoefs[:numSamples:numSamples], :155 */

    uint32_t mean = p->mb0;
    int32_t zmode = 0;
    int64_t count = 0;
    const uint32_t pb = p->pb, kb = p->kb, wb = p->wb;
    uint32_t residual;

    while (count < n) {
        if (bit_pos >= max_pos) return AO_ERR_BITSTREAM_OVERRUN; /* :168-170 */
        uint32_t m = mean >> 9;
        int32_t k = lg3a((int32_t)m);
        if ((int32_t)kb < k) k = (int32_t)kb;
        m = shl_u(1u, (uint32_t)k) - 1u;
        {
            int64_t off = (int64_t)(bit_pos >> 3);
            if (off > avail) return AO_ERR_REF_PANIC;
            uint32_t stream = rd32(in, off);
            stream <<= bit_pos & 7;
            residual = (uint32_t)lead((int32_t)~stream);
            if (residual >= 9) { /* escape: getStreamBits(input, bitPos+9, maxSize), :86-108 */
                uint32_t bo = bit_pos + 9;
                int64_t byte_off = (int64_t)(bo / 8);
                if (byte_off > avail) return AO_ERR_REF_PANIC;
                uint32_t load1 = rd32(in, byte_off);
                uint32_t nb = max_size;
                if (nb + (bo & 7) > 32) {
                    uint32_t res = load1 << (bo & 7);
                    if (byte_off + 4 >= avail + 4) return AO_ERR_REF_PANIC; /* input[byteOffset+4] */
                    uint32_t load2 = in[byte_off + 4];
                    uint32_t l2s = 8u - (nb + (bo & 7) - 32u);
                    load2 = shr_u(load2, l2s);
                    res = shr_u(res, 32u - nb);
                    res |= load2;
                    residual = res;
                } else {
                    uint32_t res = shr_u(load1, 32u - nb - (bo & 7));
                    if (nb < 32) res &= shl_u(1u, nb) - 1u;
                    residual = res;
                }
                bit_pos += 9 + max_size;
            } else {
                bit_pos += residual + 1;
                if (k != 1) {
                    stream = shl_u(stream, residual + 1);
                    uint32_t v = shr_u(stream, 32u - (uint32_t)k);
                    if (v >= 2) {
                        residual = residual * m + v - 1;
                        bit_pos += (uint32_t)k;
                    } else {
                        residual *= m;
                        bit_pos += (uint32_t)k - 1u;
                    }
                }
            }
        }
        uint32_t ndecode = residual + (uint32_t)zmode;
        int32_t mult = -(int32_t)(ndecode & 1);
        mult |= 1;
        pc[count] = (int32_t)((ndecode + 1) >> 1) * mult;
        count++;

        mean = pb * (residual + (uint32_t)zmode) + mean - ((pb * mean) >> 9);
        if (residual > 0xffff) mean = 0xffff;
        zmode = 0;

        if ((mean << 2) < 512u && count < n) {
            zmode = 1;
            int32_t k32 = lead((int32_t)mean) - 24 + (int32_t)((mean + 16) >> 6);
            if (k32 < 0) k32 = 0;
            uint32_t mz = (shl_u(1u, (uint32_t)k32) - 1u) & wb;
            /* dynGet, :112-144 */
            uint32_t tb = bit_pos;
            int64_t off = (int64_t)(tb >> 3);
            if (off > avail) return AO_ERR_REF_PANIC;
            uint32_t stream = rd32(in, off);
            stream <<= tb & 7;
            uint32_t pre = (uint32_t)lead((int32_t)~stream);
            uint32_t run;
            if (pre >= 9) {
                pre = 9;
                tb += pre;
                stream <<= pre;
                run = stream >> 16;
                tb += 16;
            } else {
                tb += pre + 1;
                stream = shl_u(stream, pre + 1);
                uint32_t val = shr_u(stream, 32u - (uint32_t)k32);
                tb += (uint32_t)k32;
                if (val < 2) {
                    run = pre * mz;
                    tb--;
                } else {
                    run = pre * mz + val - 1;
                }
            }
            bit_pos = tb;
            if (count + (int64_t)run > n) return AO_ERR_SAMPLE_OVERRUN; /* :232-234 */
            memset(pc + count, 0, (size_t)run * sizeof(int32_t));
            count += (int64_t)run;
            if (run >= 65535) zmode = 0;
            mean = 0;
        }
    }
    bb_advance(bb, bit_pos - start_pos);
    return AO_OK;
}

/* ---- predictor.go --------------------------------------------------------------------------- */
static inline int32_t sign_of(int32_t v) { return (int32_t)((uint32_t)(-v) >> 31) | (v >> 31); } /* :35-39 */

/* unpcBlock4/5/6/8, predictor.go:99-618: coefficients are int32 locals for the whole block. */
#define UNPC_FIXED(ORDER)                                                                            \
    static void unpc_fixed##ORDER(const int32_t *pc1, int32_t *out, int64_t num, const int16_t *coefs, \
                                  uint32_t chan_shift, uint32_t den_shift, int32_t den_half) {       \
        int32_t c[ORDER];                                                                            \
        for (int j = 0; j < ORDER; j++) c[j] = coefs[j];                                             \
        for (int64_t idx = ORDER + 1; idx < num; idx++) {                                            \
            const int32_t *w = out + idx - (ORDER + 1);                                              \
            int32_t top = w[0];                                                                      \
            int32_t d[ORDER];                                                                        \
            int32_t sum = den_half;                                                                  \
            for (int j = 0; j < ORDER; j++) {                                                        \
                d[j] = top - w[ORDER - j];                                                           \
                sum -= c[j] * d[j];                                                                  \
            }                                                                                        \
            int32_t sum1 = sum >> den_shift;                                                         \
            int32_t del = pc1[idx];                                                                  \
            int32_t del0 = del;                                                                      \
            int32_t sign = sign_of(del);                                                             \
            del += top + sum1;                                                                       \
            out[idx] = sext_go(del, chan_shift);                                                     \
            if (sign > 0) {                                                                          \
                int j;                                                                               \
                for (j = ORDER - 1; j >= 1; j--) {                                                   \
                    int32_t sgn = sign_of(d[j]);                                                     \
                    c[j] -= sgn;                                                                     \
                    del0 -= (ORDER - j) * ((sgn * d[j]) >> den_shift);                               \
                    if (del0 <= 0) break;                                                            \
                }                                                                                    \
                if (j == 0) c[0] -= sign_of(d[0]);                                                   \
            } else if (sign < 0) {                                                                   \
                int j;                                                                               \
                for (j = ORDER - 1; j >= 1; j--) {                                                   \
                    int32_t sgn = -sign_of(d[j]);                                                    \
                    c[j] -= sgn;                                                                     \
                    del0 -= (ORDER - j) * ((sgn * d[j]) >> den_shift);                               \
                    if (del0 >= 0) break;                                                            \
                }                                                                                    \
                if (j == 0) c[0] += sign_of(d[0]);                                                   \
            }                                                                                        \
        }                                                                                            \
    }
UNPC_FIXED(4)
UNPC_FIXED(5)
UNPC_FIXED(6)
UNPC_FIXED(8)

/* unpcBlockGeneral, predictor.go:623-684: coefficients wrap at int16 on every update. */
static void unpc_general(const int32_t *pc1, int32_t *out, int64_t num, int16_t *coefs, int order,
                         uint32_t chan_shift, uint32_t den_shift, int32_t den_half) {
    const int lim = order + 1;
    for (int64_t idx = lim; idx < num; idx++) {
        const int32_t *hist = out + idx - lim;
        int32_t top = hist[0];
        int32_t sum1 = 0;
        for (int k = 0; k < order; k++) sum1 += (int32_t)coefs[k] * (hist[order - k] - top);
        int32_t del = pc1[idx];
        int32_t del0 = del;
        int32_t sign = sign_of(del);
        del += top + ((sum1 + den_half) >> den_shift);
        out[idx] = sext_go(del, chan_shift);
        if (sign > 0) {
            for (int k = order - 1; k >= 0; k--) {
                int32_t dd = top - hist[order - k];
                int32_t sgn = sign_of(dd);
                coefs[k] = (int16_t)(coefs[k] - (int16_t)sgn);
                del0 -= (int32_t)(order - k) * ((sgn * dd) >> den_shift);
                if (del0 <= 0) break;
            }
        } else if (sign < 0) {
            for (int k = order - 1; k >= 0; k--) {
                int32_t dd = top - hist[order - k];
                int32_t sgn = sign_of(dd);
                coefs[k] = (int16_t)(coefs[k] + (int16_t)sgn);
                del0 -= (int32_t)(order - k) * ((-sgn * dd) >> den_shift);
                if (del0 >= 0) break;
            }
        }
    }
}

/* UnpcBlock, predictor.go:45-94. cap = frame_length (len of pc1/out). Returns 0 or REF_PANIC. */
static int32_t unpc_block(const int32_t *pc1, int32_t *out, int64_t num, int16_t *coefs, int32_t num_active,
                          uint32_t chan_bits, uint32_t den_shift, int64_t cap) {
    uint32_t chan_shift = 32u - chan_bits; /* wraps for chanBits 33, :46 */
    int32_t den_half = den_shift > 0 ? (int32_t)(1u << (den_shift - 1)) : 0;
    if (cap < 1) return AO_ERR_REF_PANIC;
    out[0] = pc1[0];
    if (num_active == 0) {
        if (num > 1 && pc1 != out) memmove(out + 1, pc1 + 1, (size_t)(num - 1) * sizeof(int32_t));
        return AO_OK;
    }
    if (num_active == 31) {
        int32_t prev = out[0];
        for (int64_t idx = 1; idx < num; idx++) {
            int32_t del = pc1[idx] + prev;
            prev = sext_go(del, chan_shift);
            out[idx] = prev;
        }
        return AO_OK;
    }
    if ((int64_t)num_active >= cap) return AO_ERR_REF_PANIC; /* warm-up indexes [1..numActive], :76-79 */
    for (int32_t idx = 1; idx <= num_active; idx++) {
        int32_t del = pc1[idx] + out[idx - 1];
        out[idx] = sext_go(del, chan_shift);
    }
    switch (num_active) {
    case 4: unpc_fixed4(pc1, out, num, coefs, chan_shift, den_shift, den_half); break;
    case 5: unpc_fixed5(pc1, out, num, coefs, chan_shift, den_shift, den_half); break;
    case 6: unpc_fixed6(pc1, out, num, coefs, chan_shift, den_shift, den_half); break;
    case 8: unpc_fixed8(pc1, out, num, coefs, chan_shift, den_shift, den_half); break;
    default: unpc_general(pc1, out, num, coefs, num_active, chan_shift, den_shift, den_half); break;
    }
    return AO_OK;
}

/* ---- matrix.go ------------------------------------------------------------------------------ */
static inline void put_le(uint8_t *dst, int32_t v, int bps) {
    dst[0] = (uint8_t)v;
    dst[1] = (uint8_t)(v >> 8);
    if (bps >= 3) dst[2] = (uint8_t)(v >> 16);
    if (bps == 4) dst[3] = (uint8_t)(v >> 24);
}

/* WriteStereo16/20/24/32, matrix.go:30-215. */
static void write_stereo(uint8_t *out, const int32_t *mix_u, const int32_t *mix_v, int chan_idx, int num_chan,
                         int64_t n, int32_t mix_bits, int32_t mix_res, const uint16_t *shift_buf,
                         int bytes_shifted, int depth) {
    const int bps = ao_bytes_per_sample((uint8_t)depth);
    const int64_t stride = (int64_t)num_chan * bps;
    int64_t off = (int64_t)chan_idx * bps;
    const int use_shift = (depth == 24 || depth == 32) && bytes_shifted != 0; /* 16/20 ignore it */
    const uint32_t shift = (uint32_t)bytes_shifted * 8u;
    for (int64_t i = 0; i < n; i++) {
        int32_t left, right;
        if (mix_res != 0) {
            left = mix_u[i] + mix_v[i] - sar_s(mix_res * mix_v[i], (uint32_t)mix_bits);
            right = left - mix_v[i];
        } else {
            left = mix_u[i];
            right = mix_v[i];
        }
        if (depth == 20) {
            left = shl_s(left, 4);
            right = shl_s(right, 4);
        }
        if (use_shift) {
            left = shl_s(left, shift) | (int32_t)shift_buf[i * 2 + 0];
            right = shl_s(right, shift) | (int32_t)shift_buf[i * 2 + 1];
        }
        put_le(out + off, left, bps);
        put_le(out + off + bps, right, bps);
        off += stride;
    }
}

/* WriteMono16/20/24/32, matrix.go:220-301. */
static void write_mono(uint8_t *out, const int32_t *mix_u, int chan_idx, int num_chan, int64_t n,
                       const uint16_t *shift_buf, int bytes_shifted, int depth) {
    const int bps = ao_bytes_per_sample((uint8_t)depth);
    const int64_t stride = (int64_t)num_chan * bps;
    int64_t off = (int64_t)chan_idx * bps;
    const int use_shift = (depth == 24 || depth == 32) && bytes_shifted != 0;
    const uint32_t shift = (uint32_t)bytes_shifted * 8u;
    for (int64_t i = 0; i < n; i++) {
        int32_t val = mix_u[i];
        if (depth == 20) val = shl_s(val, 4);
        if (use_shift) val = shl_s(val, shift) | (int32_t)shift_buf[i];
        put_le(out + off, val, bps);
        off += stride;
    }
}

/* ---- decoder.go ----------------------------------------------------------------------------- */
static const int8_t k_layout[8][8] = { /* channelLayoutOffsets, decoder.go:55-64 */
    {0}, {0, 1}, {2, 0, 1}, {2, 0, 1, 3}, {2, 0, 1, 3, 4}, {2, 0, 1, 4, 5, 3}, {2, 0, 1, 4, 5, 6, 3},
    {2, 6, 7, 0, 1, 4, 5, 3}};

typedef struct {
    const ao_config *cfg;
    int32_t *mix_u, *mix_v, *pred;
    uint16_t *shift_buf;
    uint8_t *bitmem;
    size_t bitmem_cap;
    bitbuf bits;
} decoder;

static int decoder_init(decoder *d, const ao_config *cfg) {
    size_t fl = cfg->frame_length;
    memset(d, 0, sizeof(*d));
    d->cfg = cfg;
    d->mix_u = (int32_t *)calloc(fl, sizeof(int32_t));
    d->mix_v = (int32_t *)calloc(fl, sizeof(int32_t));
    d->pred = (int32_t *)calloc(fl, sizeof(int32_t));
    d->shift_buf = (uint16_t *)calloc(fl * 2, sizeof(uint16_t));
    return d->mix_u && d->mix_v && d->pred && d->shift_buf ? 0 : -1;
}
static void decoder_free(decoder *d) {
    free(d->mix_u);
    free(d->mix_v);
    free(d->pred);
    free(d->shift_buf);
    free(d->bitmem);
}

typedef struct { uint32_t mode, den_shift, pb_factor, num; int16_t coefs[32]; } chan_hdr;

static void read_chan_hdr(bitbuf *b, chan_hdr *h) { /* decoder.go:275-286, :424-448 */
    uint32_t hb = bb_read(b, 8);
    h->mode = hb >> 4;
    h->den_shift = hb & 0xf;
    hb = bb_read(b, 8);
    h->pb_factor = hb >> 5;
    h->num = hb & 0x1f;
    memset(h->coefs, 0, sizeof(h->coefs));
    for (uint32_t i = 0; i < h->num; i++) h->coefs[i] = (int16_t)bb_read(b, 16);
}

/* One channel of decodeSCECompressed / decodeCPECompressed: entropy + optional delta + predictor. */
static int32_t decode_channel(decoder *d, bitbuf *b, chan_hdr *h, int32_t *dst, int64_t n, uint32_t chan_bits) {
    const ao_config *cfg = d->cfg;
    const int64_t cap = cfg->frame_length;
    agparams ag;
    set_ag_params(&ag, cfg->mb, ((uint32_t)cfg->pb * h->pb_factor) / 4, cfg->kb); /* :296-300 */
    int32_t st = dyn_decomp(&ag, b, d->pred, n, cap, chan_bits);
    if (st != AO_OK) return st;
    if (h->mode != 0) { /* :306-308 */
        st = unpc_block(d->pred, d->pred, n, NULL, 31, chan_bits, 0, cap);
        if (st != AO_OK) return st;
    }
    return unpc_block(d->pred, dst, n, h->coefs, (int32_t)h->num, chan_bits, h->den_shift, cap);
}

/* decodeSCEEscape / decodeCPEEscape sample, decoder.go:326-345, :504-535 */
static inline int32_t read_escape_sample(bitbuf *b, uint32_t chan_bits) {
    uint32_t shift = 32u - chan_bits;
    if (chan_bits <= 16) {
        int32_t val = (int32_t)bb_read(b, chan_bits);
        return sar_s(shl_s(val, shift), shift);
    }
    uint32_t extra = chan_bits - 16;
    int32_t val = (int32_t)bb_read(b, 16);
    val = sar_s(shl_s(val, 16), shift);
    return val | (int32_t)bb_read(b, extra);
}

/* Would a writer starting at out channel `chan_idx`, `width` channels wide, index past cap(out)?
 * (dst := out[off:off+N:off+N], matrix.go:44 etc.; cap(out) = frame_length*num_chan*bps) */
static int writer_panics(const ao_config *cfg, int chan_idx, int width, int64_t n) {
    if (n <= 0) return 0;
    const int64_t bps = ao_bytes_per_sample(cfg->bit_depth);
    const int64_t stride = (int64_t)cfg->num_channels * bps;
    return (n - 1) * stride + (int64_t)(chan_idx + width) * bps > (int64_t)cfg->frame_length * stride;
}

/* decodeSCE, decoder.go:210-265. Returns status; *ns updated. */
static int32_t decode_sce(decoder *d, uint8_t *out, int chan_idx, uint32_t *ns) {
    const ao_config *cfg = d->cfg;
    bitbuf *b = &d->bits;
    const int64_t cap = cfg->frame_length;
    (void)bb_read_small(b, 4);
    uint32_t unused = bb_read(b, 12);
    if (b->panic) return AO_ERR_REF_PANIC;
    if (unused != 0) return AO_ERR_INVALID_HEADER;
    uint32_t hb = bb_read(b, 4);
    if (b->panic) return AO_ERR_REF_PANIC;
    uint32_t partial = hb >> 3;
    int bytes_shifted = (int)((hb >> 1) & 3);
    if (bytes_shifted == 3) return AO_ERR_INVALID_SHIFT;
    uint32_t escape = hb & 1;
    uint32_t chan_bits = (uint32_t)cfg->bit_depth - (uint32_t)bytes_shifted * 8u;
    uint32_t n = *ns;
    if (partial != 0) {
        n = bb_read(b, 16) << 16;
        n |= bb_read(b, 16);
        if (b->panic) return AO_ERR_REF_PANIC;
    }
    if (escape == 0) { /* decodeSCECompressed, :267-324 */
        (void)bb_read(b, 8);
        (void)bb_read(b, 8);
        chan_hdr hu;
        read_chan_hdr(b, &hu);
        if (b->panic) return AO_ERR_REF_PANIC;
        bitbuf shift_bits = *b;
        if (bytes_shifted != 0) bb_advance(b, (uint32_t)bytes_shifted * 8u * n);
        int32_t st = decode_channel(d, b, &hu, d->mix_u, (int64_t)n, chan_bits);
        if (st != AO_OK) return st == AO_ERR_REF_PANIC ? st : AO_STATUS(st, 0, AO_ENT_MONO);
        if (bytes_shifted != 0) {
            if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
            for (uint32_t i = 0; i < n; i++) d->shift_buf[i] = (uint16_t)bb_read(&shift_bits, (uint32_t)bytes_shifted * 8u);
            if (shift_bits.panic) return AO_ERR_REF_PANIC;
        }
    } else { /* decodeSCEEscape, :326-345 */
        if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
        for (uint32_t i = 0; i < n; i++) {
            d->mix_u[i] = read_escape_sample(b, chan_bits);
            if (b->panic) return AO_ERR_REF_PANIC;
        }
        bytes_shifted = 0;
    }
    if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
    if (writer_panics(cfg, chan_idx, 1, n)) return AO_ERR_REF_PANIC;
    write_mono(out, d->mix_u, chan_idx, cfg->num_channels, n, d->shift_buf, bytes_shifted, cfg->bit_depth);
    *ns = n;
    return AO_OK;
}

/* decodeCPE, decoder.go:348-414 */
static int32_t decode_cpe(decoder *d, uint8_t *out, int chan_idx, uint32_t *ns) {
    const ao_config *cfg = d->cfg;
    bitbuf *b = &d->bits;
    const int64_t cap = cfg->frame_length;
    (void)bb_read_small(b, 4);
    uint32_t unused = bb_read(b, 12);
    if (b->panic) return AO_ERR_REF_PANIC;
    if (unused != 0) return AO_ERR_INVALID_HEADER;
    uint32_t hb = bb_read(b, 4);
    if (b->panic) return AO_ERR_REF_PANIC;
    uint32_t partial = hb >> 3;
    int bytes_shifted = (int)((hb >> 1) & 3);
    if (bytes_shifted == 3) return AO_ERR_INVALID_SHIFT;
    uint32_t escape = hb & 1;
    uint32_t chan_bits = (uint32_t)cfg->bit_depth - (uint32_t)bytes_shifted * 8u + 1u;
    uint32_t n = *ns;
    if (partial != 0) {
        n = bb_read(b, 16) << 16;
        n |= bb_read(b, 16);
        if (b->panic) return AO_ERR_REF_PANIC;
    }
    int32_t mix_bits = 0, mix_res = 0;
    if (escape == 0) { /* decodeCPECompressed, :416-502 */
        mix_bits = (int32_t)bb_read(b, 8);
        mix_res = (int32_t)(int8_t)bb_read(b, 8);
        chan_hdr hu, hv;
        read_chan_hdr(b, &hu);
        read_chan_hdr(b, &hv);
        if (b->panic) return AO_ERR_REF_PANIC;
        bitbuf shift_bits = *b;
        if (bytes_shifted != 0) bb_advance(b, (uint32_t)bytes_shifted * 8u * 2u * n);
        int32_t st = decode_channel(d, b, &hu, d->mix_u, (int64_t)n, chan_bits);
        if (st != AO_OK) return st == AO_ERR_REF_PANIC ? st : AO_STATUS(st, 0, AO_ENT_U);
        st = decode_channel(d, b, &hv, d->mix_v, (int64_t)n, chan_bits);
        if (st != AO_OK) return st == AO_ERR_REF_PANIC ? st : AO_STATUS(st, 0, AO_ENT_V);
        if (bytes_shifted != 0) {
            if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
            for (uint32_t i = 0; i < 2 * n; i++) d->shift_buf[i] = (uint16_t)bb_read(&shift_bits, (uint32_t)bytes_shifted * 8u);
            if (shift_bits.panic) return AO_ERR_REF_PANIC;
        }
    } else { /* decodeCPEEscape, :504-535 */
        chan_bits = cfg->bit_depth;
        if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
        for (uint32_t i = 0; i < n; i++) {
            d->mix_u[i] = read_escape_sample(b, chan_bits);
            d->mix_v[i] = read_escape_sample(b, chan_bits);
            if (b->panic) return AO_ERR_REF_PANIC;
        }
        bytes_shifted = 0;
    }
    if ((int64_t)n > cap) return AO_ERR_REF_PANIC;
    if (writer_panics(cfg, chan_idx, 2, n)) return AO_ERR_REF_PANIC;
    write_stereo(out, d->mix_u, d->mix_v, chan_idx, cfg->num_channels, n, mix_bits, mix_res, d->shift_buf,
                 bytes_shifted, cfg->bit_depth);
    *ns = n;
    return AO_OK;
}

/* decodePacketInto, decoder.go:133-207 */
static int32_t decode_packet_into(decoder *d, const uint8_t *packet, size_t size, uint8_t *out, uint32_t *out_bytes) {
    const ao_config *cfg = d->cfg;
    *out_bytes = 0;
    /* bits.Reset(packet): private padded copy, bitbuffer.go:36-51 */
    if (d->bitmem_cap < size + 4) {
        free(d->bitmem);
        d->bitmem_cap = size + 4 + 4096;
        d->bitmem = (uint8_t *)malloc(d->bitmem_cap);
    }
    if (size) memcpy(d->bitmem, packet, size);
    memset(d->bitmem + size, 0, 4);
    bitbuf *b = &d->bits;
    b->buf = d->bitmem;
    b->cap = (int64_t)size + 4;
    b->size = (int64_t)size;
    b->pos = 0;
    b->bitidx = 0;
    b->panic = 0;

    uint32_t ns = cfg->frame_length;
    const int num_chan = cfg->num_channels;
    const int bps = ao_bytes_per_sample(cfg->bit_depth);
    int chan_idx = 0;
    const int8_t *offsets = k_layout[num_chan - 1];

    for (;;) {
        if (bb_past_end(b)) return AO_ERR_BITSTREAM_OVERRUN; /* :143-145 */
        uint8_t tag = bb_read_small(b, 3);
        if (b->panic) return AO_ERR_REF_PANIC;
        switch (tag) {
        case 0:
        case 3: {
            int32_t st = decode_sce(d, out, offsets[chan_idx], &ns);
            if (st != AO_OK) return st | (AO_CTX_SCE << 8);
            chan_idx++;
            break;
        }
        case 1: {
            if (chan_idx + 2 > num_chan) goto done; /* :163-165 */
            int32_t st = decode_cpe(d, out, offsets[chan_idx], &ns);
            if (st != AO_OK) return st | (AO_CTX_CPE << 8);
            chan_idx += 2;
            break;
        }
        case 2:
        case 5: return AO_ERR_UNSUPPORTED_ELEMENT; /* :179-180 */
        case 4: {                                  /* skipDSE, :553-574 */
            (void)bb_read_small(b, 4);
            uint8_t align = bb_read_one(b);
            uint16_t count = bb_read_small(b, 8);
            if (count == 255) count = (uint16_t)(count + bb_read_small(b, 8));
            if (b->panic) return AO_STATUS(AO_ERR_REF_PANIC, AO_CTX_DSE, 0);
            if (align != 0) bb_byte_align(b);
            bb_advance(b, (uint32_t)count * 8u);
            if (bb_past_end(b)) return AO_STATUS(AO_ERR_BITSTREAM_OVERRUN, AO_CTX_DSE, 0);
            break;
        }
        case 6: { /* skipFIL, :538-551 */
            int16_t count = (int16_t)bb_read_small(b, 4);
            if (count == 15) count = (int16_t)(count + (int16_t)bb_read_small(b, 8) - 1);
            if (b->panic) return AO_STATUS(AO_ERR_REF_PANIC, AO_CTX_FIL, 0);
            bb_advance(b, (uint32_t)count * 8u);
            if (bb_past_end(b)) return AO_STATUS(AO_ERR_BITSTREAM_OVERRUN, AO_CTX_FIL, 0);
            break;
        }
        case 7: /* :192-195 */
            bb_byte_align(b);
            goto done;
        }
        if (chan_idx >= num_chan) break; /* :200-202 */
    }
done:
    *out_bytes = ns * (uint32_t)num_chan * (uint32_t)bps; /* :206 */
    return AO_OK;
}

int32_t ao_decode_packet(const ao_config *cfg, const uint8_t *packet, size_t size, uint8_t *out,
                         uint32_t *out_bytes) {
    *out_bytes = 0;
    int32_t st = ao_check_config(cfg);
    if (st != AO_OK) return st;
    decoder d;
    if (decoder_init(&d, cfg) != 0) {
        decoder_free(&d);
        return AO_ERR_REF_PANIC;
    }
    memset(out, 0, (size_t)cfg->frame_length * cfg->num_channels * (size_t)ao_bytes_per_sample(cfg->bit_depth));
    st = decode_packet_into(&d, packet, size, out, out_bytes);
    if (st != AO_OK) *out_bytes = 0;
    decoder_free(&d);
    return st;
}

/* ---- batch driver (CPU baseline) ------------------------------------------------------------- */
typedef struct {
    const ao_config *cfg;
    const uint8_t *packed;
    const uint64_t *offsets;
    const uint32_t *sizes;
    uint32_t lo, hi;
    uint8_t *out;
    uint64_t out_stride;
    uint32_t *out_bytes;
    int32_t *status;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    decoder d;
    const size_t frame_bytes =
        (size_t)j->cfg->frame_length * j->cfg->num_channels * (size_t)ao_bytes_per_sample(j->cfg->bit_depth);
    if (decoder_init(&d, j->cfg) != 0) {
        for (uint32_t i = j->lo; i < j->hi; i++) { j->status[i] = AO_ERR_REF_PANIC; j->out_bytes[i] = 0; }
        decoder_free(&d);
        return NULL;
    }
    for (uint32_t i = j->lo; i < j->hi; i++) {
        uint8_t *o = j->out + (uint64_t)i * j->out_stride;
        memset(o, 0, frame_bytes);
        j->status[i] = decode_packet_into(&d, j->packed + j->offsets[i], j->sizes[i], o, &j->out_bytes[i]);
        if (j->status[i] != AO_OK) j->out_bytes[i] = 0;
    }
    decoder_free(&d);
    return NULL;
}

void ao_decode_batch(const ao_config *cfg, const uint8_t *packed, const uint64_t *offsets,
                     const uint32_t *sizes, uint32_t n, uint8_t *out, uint64_t out_stride,
                     uint32_t *out_bytes, int32_t *status, int nthreads) {
    int32_t st = ao_check_config(cfg);
    if (st != AO_OK) {
        for (uint32_t i = 0; i < n; i++) { status[i] = st; out_bytes[i] = 0; }
        return;
    }
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
    batch_job *jobs = (batch_job *)calloc((size_t)nthreads, sizeof(batch_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    /* contiguous ranges balanced by compressed bytes */
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += sizes[i];
    uint32_t lo = 0;
    uint64_t acc = 0;
    for (int t = 0; t < nthreads; t++) {
        uint32_t hi = lo;
        uint64_t target = total * (uint64_t)(t + 1) / (uint64_t)nthreads;
        while (hi < n && (acc + sizes[hi] <= target || hi == lo) && (n - hi) > (uint32_t)(nthreads - 1 - t)) {
            acc += sizes[hi];
            hi++;
        }
        if (t == nthreads - 1) hi = n;
        jobs[t] = (batch_job){cfg, packed, offsets, sizes, lo, hi, out, out_stride, out_bytes, status};
        lo = hi;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&tids[t], NULL, batch_worker, &jobs[t]);
    batch_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
}
