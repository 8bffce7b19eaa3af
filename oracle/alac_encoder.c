/*
 * alac_encoder.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Test-side ALAC packet encoder / bitstream synthesiser. The reference ships no encoder and no
 * fixtures; its tests get packets from external encoders. This file makes
 * packets of every shape the decoder grammar accepts, including the ones
 * FFmpeg never emits (20/32-bit, mode != 0, order 0..31, pbFactor != 4, tag 3, DSE/FIL,
 * bytesShifted 2, escape + partial), and the large synthetic streams bench.py decodes.
 *
 * It is the exact inverse of the decoder arithmetic:
 *   - un-mix inverse of matrix.go:40-41      (v = L-R, u = R + ((mixRes*v) >> mixBits))
 *   - forward sign-LMS predictor mirroring predictor.go:45-684 state updates
 *   - adaptive Golomb-Rice writer mirroring golomb.go:148-253 state updates
 * Streams it writes are cross-checked against FFmpeg's independent ALAC decoder in
 * tests/golden/gen_ffmpeg_fixtures.py (run in the build container, results committed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Go shift semantics -------------------------------------------------------------------- */
static inline uint32_t shl_u(uint32_t x, uint32_t s) { return s >= 32 ? 0u : x << s; }
static inline int32_t shl_s(int32_t x, uint32_t s) { return s >= 32 ? 0 : (int32_t)((uint32_t)x << s); }
static inline int32_t sar_s(int32_t x, uint32_t s) { return s >= 32 ? (x < 0 ? -1 : 0) : x >> s; }
static inline int32_t sext_go(int32_t x, uint32_t cs) { return sar_s(shl_s(x, cs), cs); }
static inline int32_t sign_of(int32_t v) { return (int32_t)((uint32_t)(-v) >> 31) | (v >> 31); }
static inline int32_t lead(int32_t m) { return m == 0 ? 32 : (int32_t)__builtin_clz((uint32_t)m); }
static inline int32_t lg3a(int32_t x) { return 31 - lead(x + 3); }

/* ---- MSB-first bit writer -------------------------------------------------------------------- */
typedef struct {
    uint8_t *buf;
    size_t cap;
    uint64_t bitpos;
    int overflow;
} ae_writer;

void ae_writer_init(ae_writer *w, uint8_t *buf, size_t cap) {
    w->buf = buf;
    w->cap = cap;
    w->bitpos = 0;
    w->overflow = 0;
    memset(buf, 0, cap);
}
void ae_put_bits(ae_writer *w, uint32_t value, uint32_t nbits) { /* nbits <= 32 */
    for (uint32_t i = 0; i < nbits; i++) {
        uint32_t bit = (value >> (nbits - 1 - i)) & 1u;
        size_t byte = (size_t)(w->bitpos >> 3);
        if (byte >= w->cap) { w->overflow = 1; return; }
        if (bit) w->buf[byte] |= (uint8_t)(0x80u >> (w->bitpos & 7));
        w->bitpos++;
    }
}
static void put_ones(ae_writer *w, uint32_t n) { while (n--) ae_put_bits(w, 1, 1); }
void ae_byte_align(ae_writer *w) { w->bitpos = (w->bitpos + 7) & ~(uint64_t)7; }
uint64_t ae_writer_bits(const ae_writer *w) { return w->bitpos; }
size_t ae_writer_bytes(const ae_writer *w) { return (size_t)((w->bitpos + 7) >> 3); }
int ae_writer_overflow(const ae_writer *w) { return w->overflow; }

/* ---- adaptive Golomb writer (mirror of golomb.go:148-253) ---------------------------------- */
typedef struct { uint32_t mb, pb, kb; } ae_ag;

/* Encodes residual array pc[0..n) (signed) with escape width max_size bits. Returns 0, or -1 when
 * a value cannot be represented (caller falls back to an escape element). */
static int golomb_write(ae_writer *w, const ae_ag *p, const int32_t *pc, int64_t n, uint32_t max_size) {
    uint32_t mean = p->mb;
    uint32_t zmode = 0;
    const uint32_t wb = shl_u(1u, p->kb) - 1u;
    int64_t count = 0;
    while (count < n) {
        uint32_t m = mean >> 9;
        int32_t k = lg3a((int32_t)m);
        if ((int32_t)p->kb < k) k = (int32_t)p->kb;
        m = shl_u(1u, (uint32_t)k) - 1u;
        int32_t x = pc[count];
        uint32_t ndecode = x >= 0 ? (uint32_t)x * 2u : (uint32_t)(-(int64_t)x) * 2u - 1u;
        if (ndecode < zmode) return -1; /* a zero directly after a non-maximal run: caller bug */
        uint32_t r = ndecode - zmode;
        uint32_t div = m ? r / m : (r ? 9u : 0u);
        uint32_t mod = m ? r % m : 0u;
        if (div >= 9) {
            if (max_size < 32 && (r >> max_size) != 0) return -1;
            if (max_size > 32) return -1; /* the reference's 33-bit escape read is not invertible */
            put_ones(w, 9);
            ae_put_bits(w, r, max_size);
        } else {
            put_ones(w, div);
            ae_put_bits(w, 0, 1);
            if (k != 1) {
                if (k == 0) return -1;
                if (mod == 0) ae_put_bits(w, 0, (uint32_t)k - 1u);
                else ae_put_bits(w, mod + 1u, (uint32_t)k);
            }
        }
        count++;
        mean = p->pb * (r + zmode) + mean - ((p->pb * mean) >> 9);
        if (r > 0xffff) mean = 0xffff;
        zmode = 0;
        if ((mean << 2) < 512u && count < n) {
            zmode = 1;
            int32_t k32 = lead((int32_t)mean) - 24 + (int32_t)((mean + 16) >> 6);
            if (k32 < 0) k32 = 0;
            uint32_t mz = (shl_u(1u, (uint32_t)k32) - 1u) & wb;
            uint32_t run = 0;
            while (count + run < n && pc[count + run] == 0 && run < 65535u) run++;
            uint32_t d2 = mz ? run / mz : 9u;
            uint32_t m2 = mz ? run % mz : 0u;
            if (d2 >= 9 || k32 == 0) {
                put_ones(w, 9);
                ae_put_bits(w, run, 16);
            } else {
                put_ones(w, d2);
                ae_put_bits(w, 0, 1);
                if (m2 == 0) ae_put_bits(w, 0, (uint32_t)k32 - 1u);
                else ae_put_bits(w, m2 + 1u, (uint32_t)k32);
            }
            count += run;
            if (run >= 65535u) zmode = 0;
            mean = 0;
        }
    }
    return 0;
}

/* ---- forward predictor (mirror of predictor.go) --------------------------------------------- */
/* Given the target signal x[0..n) (already representable in chan_bits), produce pc1[0..n) such
 * that UnpcBlock(pc1, order, coefs, den_shift) returns x. coefs are the INITIAL coefficients. */
static void forward_predict(const int32_t *x, int32_t *pc1, int64_t n, const int16_t *coefs_in, int order,
                            uint32_t chan_bits, uint32_t den_shift) {
    const uint32_t cs = 32u - chan_bits;
    const int32_t den_half = den_shift > 0 ? (int32_t)(1u << (den_shift - 1)) : 0;
    if (n <= 0) return;
    pc1[0] = x[0];
    if (order == 0) {
        for (int64_t i = 1; i < n; i++) pc1[i] = x[i];
        return;
    }
    if (order == 31) {
        for (int64_t i = 1; i < n; i++) pc1[i] = sext_go(x[i] - x[i - 1], cs);
        return;
    }
    for (int64_t i = 1; i <= order && i < n; i++) pc1[i] = sext_go(x[i] - x[i - 1], cs);
    const int fixed = (order == 4 || order == 5 || order == 6 || order == 8);
    int32_t c32[32];
    int16_t c16[32];
    for (int j = 0; j < order; j++) { c32[j] = coefs_in[j]; c16[j] = coefs_in[j]; }
    const int lim = order + 1;
    for (int64_t idx = lim; idx < n; idx++) {
        const int32_t *h = x + idx - lim;
        int32_t top = h[0];
        int32_t sum = 0;
        for (int j = 0; j < order; j++) sum += (fixed ? c32[j] : (int32_t)c16[j]) * (h[order - j] - top);
        int32_t pred = top + ((sum + den_half) >> den_shift);
        int32_t del = sext_go(x[idx] - pred, cs);
        pc1[idx] = del;
        int32_t del0 = del;
        int32_t sign = sign_of(del);
        if (sign == 0) continue;
        for (int j = order - 1; j >= 0; j--) {
            int32_t dd = top - h[order - j];
            int32_t sgn = sign_of(dd);
            if (sign > 0) {
                if (fixed) c32[j] -= sgn; else c16[j] = (int16_t)(c16[j] - (int16_t)sgn);
                del0 -= (int32_t)(order - j) * ((sgn * dd) >> den_shift);
                if (del0 <= 0) break;
            } else {
                if (fixed) c32[j] += sgn; else c16[j] = (int16_t)(c16[j] + (int16_t)sgn);
                del0 -= (int32_t)(order - j) * ((-sgn * dd) >> den_shift);
                if (del0 >= 0) break;
            }
        }
    }
}

/* ---- LPC analysis (Levinson-Durbin) for realistic coefficients -------------------------------- */
/* Returns chosen order in [min_order,max_order]; coefs[j] pairs with x[i-1-j]. */
static int lpc_analyse(const int32_t *x, int64_t n, int min_order, int max_order, uint32_t den_shift, int16_t *coefs) {
    double r[33], a[33], tmp[33], err, best_cost = 1e300;
    int best_order = min_order;
    double best_a[33];
    memset(best_a, 0, sizeof(best_a));
    if (max_order > 30) max_order = 30;
    for (int lag = 0; lag <= max_order; lag++) {
        double s = 0;
        for (int64_t i = lag; i < n; i++) {
            double wi = 1.0, wl = 1.0;
            s += wi * wl * (double)x[i] * (double)x[i - lag];
        }
        r[lag] = s;
    }
    if (r[0] <= 0) {
        memset(coefs, 0, 32 * sizeof(int16_t));
        return min_order;
    }
    r[0] *= 1.0000001;
    err = r[0];
    memset(a, 0, sizeof(a));
    for (int m = 1; m <= max_order; m++) {
        double acc = r[m];
        for (int j = 1; j < m; j++) acc -= a[j] * r[m - j];
        double kref = acc / err;
        memcpy(tmp, a, sizeof(a));
        a[m] = kref;
        for (int j = 1; j < m; j++) a[j] = tmp[j] - kref * tmp[m - j];
        err *= (1.0 - kref * kref);
        if (err <= 0) err = 1e-9;
        if (m >= min_order) {
            double cost = 0.5 * log2(err / (double)n + 1e-9) * (double)n + 16.0 * m;
            if (cost < best_cost) {
                best_cost = cost;
                best_order = m;
                memcpy(best_a, a, sizeof(a));
            }
        }
    }
    memset(coefs, 0, 32 * sizeof(int16_t));
    for (int j = 0; j < best_order; j++) {
        double q = best_a[j + 1] * (double)(1u << den_shift);
        long v = lrint(q);
        if (v > 32767) v = 32767;
        if (v < -32768) v = -32768;
        coefs[j] = (int16_t)v;
    }
    return best_order;
}

/* ---- element encoder ---------------------------------------------------------------------------- */
typedef struct {
    uint8_t mode;      /* 0 or non-zero: decoder applies the order-31 delta pass first */
    uint8_t den_shift; /* 0..15 */
    uint8_t pb_factor; /* 0..7 */
    uint8_t order;     /* 0..31 */
    int16_t coefs[32];
} ae_chan_params;

typedef struct {
    uint8_t tag;           /* 0 SCE, 1 CPE, 3 LFE */
    uint8_t instance;      /* 4-bit element instance tag */
    uint8_t escape;        /* write raw samples */
    uint8_t bytes_shifted; /* 0..2 */
    uint8_t partial;       /* write the 32-bit sample count */
    uint8_t mix_bits;
    int8_t mix_res;
    uint8_t pad_;
    ae_chan_params ch[2];
} ae_element;

typedef struct {
    uint32_t frame_length;
    uint8_t bit_depth, num_channels, pb, mb, kb, pad_;
    uint16_t max_run;
    uint32_t max_frame_bytes, avg_bit_rate, sample_rate;
} ae_config; /* same layout as ao_config */

/* Encodes one element. c0/c1: the element's channel samples (c1 NULL for SCE/LFE), n samples,
 * sign-extended at cfg->bit_depth. Returns 0, -1 if a residual was not representable (use
 * escape), -2 on bad params. */
int ae_encode_element(ae_writer *w, const ae_config *cfg, const ae_element *e, const int32_t *c0,
                      const int32_t *c1, uint32_t n) {
    const int stereo = e->tag == 1;
    const uint32_t depth = cfg->bit_depth;
    const uint32_t shift = e->bytes_shifted;
    if (shift > 2 || (stereo && !c1)) return -2;
    ae_put_bits(w, e->tag, 3);
    ae_put_bits(w, e->instance, 4);
    ae_put_bits(w, 0, 12);
    ae_put_bits(w, (uint32_t)((e->partial ? 8 : 0) | (shift << 1) | (e->escape ? 1 : 0)), 4);
    if (e->partial) {
        ae_put_bits(w, n >> 16, 16);
        ae_put_bits(w, n & 0xffff, 16);
    }
    if (e->escape) {
        /* decoder.go:326-345 (SCE keeps chanBits = depth - 8*shift), :504-535 (CPE resets to depth) */
        uint32_t cb = stereo ? depth : depth - 8 * shift;
        for (uint32_t i = 0; i < n; i++) {
            for (int c = 0; c < (stereo ? 2 : 1); c++) {
                uint32_t v = (uint32_t)(c ? c1[i] : c0[i]);
                if (cb <= 16) ae_put_bits(w, v & (shl_u(1u, cb) - 1u), cb);
                else {
                    ae_put_bits(w, (v >> (cb - 16)) & 0xffff, 16);
                    ae_put_bits(w, v & (shl_u(1u, cb - 16) - 1u), cb - 16);
                }
            }
        }
        return 0;
    }
    const uint32_t chan_bits = depth - 8 * shift + (stereo ? 1 : 0);
    int32_t *u = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1) * 4);
    int32_t *v = u + (n + 1), *pc = v + (n + 1), *tmp = pc + (n + 1);
    const uint32_t sb = shift * 8;
    const uint32_t smask = shl_u(1u, sb) - 1u;
    for (uint32_t i = 0; i < n; i++) {
        int32_t a = c0[i], b = stereo ? c1[i] : 0;
        if (shift) { a = sar_s(a, sb); b = sar_s(b, sb); }
        if (stereo && e->mix_res != 0) {
            int32_t d = a - b;
            u[i] = b + sar_s((int32_t)e->mix_res * d, e->mix_bits);
            v[i] = d;
        } else {
            u[i] = a;
            v[i] = b;
        }
    }
    ae_put_bits(w, e->mix_bits, 8);
    ae_put_bits(w, (uint32_t)(uint8_t)e->mix_res, 8);
    for (int c = 0; c < (stereo ? 2 : 1); c++) {
        const ae_chan_params *p = &e->ch[c];
        ae_put_bits(w, (uint32_t)((p->mode << 4) | (p->den_shift & 0xf)), 8);
        ae_put_bits(w, (uint32_t)((p->pb_factor << 5) | (p->order & 0x1f)), 8);
        for (int j = 0; j < p->order; j++) ae_put_bits(w, (uint16_t)p->coefs[j], 16);
    }
    if (shift) {
        for (uint32_t i = 0; i < n; i++) {
            ae_put_bits(w, (uint32_t)c0[i] & smask, sb);
            if (stereo) ae_put_bits(w, (uint32_t)c1[i] & smask, sb);
        }
    }
    int rc = 0;
    for (int c = 0; c < (stereo ? 2 : 1) && rc == 0; c++) {
        const ae_chan_params *p = &e->ch[c];
        const int32_t *sig = c ? v : u;
        forward_predict(sig, pc, n, p->coefs, p->order, chan_bits, p->den_shift);
        const int32_t *res = pc;
        if (p->mode != 0 && n > 0) { /* invert the order-31 pre-pass, decoder.go:306-308 */
            const uint32_t cs = 32u - chan_bits;
            tmp[0] = pc[0];
            for (uint32_t i = 1; i < n; i++) tmp[i] = sext_go(pc[i] - pc[i - 1], cs);
            res = tmp;
        }
        ae_ag ag = {cfg->mb, ((uint32_t)cfg->pb * p->pb_factor) / 4, cfg->kb};
        rc = golomb_write(w, &ag, res, n, chan_bits);
    }
    free(u);
    return rc;
}

void ae_write_dse(ae_writer *w, uint32_t instance, int align, const uint8_t *data, uint32_t count) {
    ae_put_bits(w, 4, 3); /* decoder.go:553-574 */
    ae_put_bits(w, instance, 4);
    ae_put_bits(w, align ? 1 : 0, 1);
    if (count >= 255) {
        ae_put_bits(w, 255, 8);
        ae_put_bits(w, count - 255, 8);
    } else ae_put_bits(w, count, 8);
    if (align) ae_byte_align(w);
    for (uint32_t i = 0; i < count; i++) ae_put_bits(w, data ? data[i] : 0, 8);
}
void ae_write_fil(ae_writer *w, uint32_t count) { /* decoder.go:538-551; count <= 269 */
    ae_put_bits(w, 6, 3);
    if (count >= 15) {
        ae_put_bits(w, 15, 4);
        ae_put_bits(w, count - 15 + 1, 8);
    } else ae_put_bits(w, count, 4);
    for (uint32_t i = 0; i < count; i++) ae_put_bits(w, 0xA5, 8);
}
void ae_write_end(ae_writer *w) {
    ae_put_bits(w, 7, 3);
    ae_byte_align(w);
}

/* ---- cookie --------------------------------------------------------------------------------------- */
static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
/* wrappers: bit0 = 'alac' atom, bit1 = 'frma' atom in front (config.go:50-58). Returns length. */
size_t ae_write_cookie(const ae_config *cfg, int wrappers, uint8_t *out) {
    uint8_t *p = out;
    if (wrappers & 2) { put_be32(p, 12); memcpy(p + 4, "frma", 4); memcpy(p + 8, "alac", 4); p += 12; }
    if (wrappers & 1) { put_be32(p, 36); memcpy(p + 4, "alac", 4); put_be32(p + 8, 0); p += 12; }
    put_be32(p, cfg->frame_length);
    p[4] = 0;
    p[5] = cfg->bit_depth;
    p[6] = cfg->pb;
    p[7] = cfg->mb;
    p[8] = cfg->kb;
    p[9] = cfg->num_channels;
    p[10] = (uint8_t)(cfg->max_run >> 8);
    p[11] = (uint8_t)cfg->max_run;
    put_be32(p + 12, cfg->max_frame_bytes);
    put_be32(p + 16, cfg->avg_bit_rate);
    put_be32(p + 20, cfg->sample_rate);
    return (size_t)(p + 24 - out);
}

/* ---- whole-packet convenience encoder ------------------------------------------------------------ */
typedef struct {
    int32_t min_order, max_order; /* LPC search range (1..30); both 0 => order 0 */
    int32_t den_shift;            /* default 9 */
    int32_t mode;                 /* 0 */
    int32_t pb_factor;            /* default 4 */
    int32_t mix_bits, mix_res;    /* stereo; mix_res == -128 => pick (0,0) or (1,1) by energy */
    int32_t bytes_shifted;        /* -1 => FFmpeg rule (depth-16)/8 clipped to 0..2 */
    int32_t force_escape;
    int32_t lfe_tag3;             /* write the LFE element with tag 3 instead of 0 */
    int32_t fil_bytes;            /* >0: a FIL element before END */
    int32_t dse_bytes;            /* >0: a DSE element before the first audio element */
    int32_t no_end;               /* omit END (decoder stops at chanIdx >= numChan) */
    int32_t always_partial;       /* write the partial header even when n == frame_length */
} ae_packet_opts;

static const int8_t k_layout[8][8] = {
    {0}, {0, 1}, {2, 0, 1}, {2, 0, 1, 3}, {2, 0, 1, 3, 4}, {2, 0, 1, 4, 5, 3}, {2, 0, 1, 4, 5, 6, 3},
    {2, 6, 7, 0, 1, 4, 5, 3}};
/* element sequence per channel count: 0 SCE, 1 CPE, 3 LFE */
static const int8_t k_elems[8][6] = {{0, -1}, {1, -1}, {0, 1, -1}, {0, 1, 0, -1}, {0, 1, 1, -1},
                                     {0, 1, 1, 3, -1}, {0, 1, 1, 0, 3, -1}, {0, 1, 1, 1, 3, -1}};

/* pcm: interleaved int32 [n][num_channels] in OUTPUT channel order, sign-extended at bit_depth.
 * Returns packet size in bytes, or -1 on overflow/bad input. */
int64_t ae_encode_packet(const ae_config *cfg, const ae_packet_opts *o, const int32_t *pcm, uint32_t n,
                         uint8_t *out, size_t cap) {
    const int nc = cfg->num_channels;
    if (nc < 1 || nc > 8) return -1;
    ae_writer w;
    ae_writer_init(&w, out, cap);
    int32_t *c0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1) * 2);
    int32_t *c1 = c0 + (n + 1);
    if (o->dse_bytes > 0) ae_write_dse(&w, 0, 1, NULL, (uint32_t)o->dse_bytes);
    int chan_idx = 0, inst[4] = {0, 0, 0, 0};
    for (int ei = 0; k_elems[nc - 1][ei] >= 0; ei++) {
        int tag = k_elems[nc - 1][ei];
        const int stereo = tag == 1;
        int oc = k_layout[nc - 1][chan_idx];
        for (uint32_t i = 0; i < n; i++) {
            c0[i] = pcm[(size_t)i * nc + oc];
            if (stereo) c1[i] = pcm[(size_t)i * nc + oc + 1];
        }
        ae_element e;
        memset(&e, 0, sizeof(e));
        e.tag = (uint8_t)((tag == 3 && !o->lfe_tag3) ? 0 : tag);
        e.instance = (uint8_t)(inst[tag]++ & 0xf);
        e.partial = (uint8_t)((n != cfg->frame_length || o->always_partial) ? 1 : 0);
        int bs = o->bytes_shifted;
        if (bs < 0) bs = cfg->bit_depth > 16 ? (cfg->bit_depth - 16) / 8 : 0;
        if (bs > 2) bs = 2;
        e.bytes_shifted = (uint8_t)bs;
        e.escape = (uint8_t)(o->force_escape ? 1 : 0);
        const uint32_t sb = (uint32_t)bs * 8;
        if (stereo) {
            if (o->mix_res == -128) {
                double e_lr = 0, e_ms = 0;
                for (uint32_t i = 0; i < n; i++) {
                    double l = (double)sar_s(c0[i], sb), r = (double)sar_s(c1[i], sb);
                    e_lr += fabs(l) + fabs(r);
                    e_ms += fabs((l + r) * 0.5) + fabs(l - r);
                }
                if (e_ms < e_lr) { e.mix_bits = 1; e.mix_res = 1; }
            } else {
                e.mix_bits = (uint8_t)o->mix_bits;
                e.mix_res = (int8_t)o->mix_res;
            }
        }
        /* the signals the predictor will see, for the LPC analysis */
        for (int c = 0; c < (stereo ? 2 : 1); c++) {
            ae_chan_params *p = &e.ch[c];
            p->mode = (uint8_t)o->mode;
            p->den_shift = (uint8_t)o->den_shift;
            p->pb_factor = (uint8_t)o->pb_factor;
            if (o->max_order <= 0) { p->order = 0; continue; }
            if (o->min_order == 31) { p->order = 31; continue; }
            int32_t *sig = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
            for (uint32_t i = 0; i < n; i++) {
                int32_t a = sar_s(c0[i], sb), b = stereo ? sar_s(c1[i], sb) : 0;
                if (stereo && e.mix_res != 0) {
                    int32_t d = a - b;
                    sig[i] = c ? d : b + sar_s((int32_t)e.mix_res * d, e.mix_bits);
                } else sig[i] = c ? b : a;
            }
            p->order = (uint8_t)lpc_analyse(sig, n, o->min_order, o->max_order, (uint32_t)o->den_shift, p->coefs);
            free(sig);
        }
        uint64_t mark = w.bitpos;
        int rc = ae_encode_element(&w, cfg, &e, c0, stereo ? c1 : NULL, n);
        if (rc == -2) { free(c0); return -1; }
        const uint64_t raw_bits = (uint64_t)n * cfg->bit_depth * (stereo ? 2 : 1) + 64;
        if (rc == -1 || (!e.escape && w.bitpos - mark > raw_bits)) {
            /* incompressible or unrepresentable: rewrite as an escape element (what encoders do) */
            size_t from = (size_t)(mark >> 3);
            uint8_t keep = (uint8_t)(w.buf[from] & (uint8_t)(0xff00u >> (mark & 7)));
            memset(w.buf + from, 0, w.cap - from);
            w.buf[from] = keep;
            w.bitpos = mark;
            e.escape = 1;
            e.bytes_shifted = 0;
            ae_encode_element(&w, cfg, &e, c0, stereo ? c1 : NULL, n);
        }
        chan_idx += stereo ? 2 : 1;
    }
    if (o->fil_bytes > 0) ae_write_fil(&w, (uint32_t)o->fil_bytes);
    if (!o->no_end) ae_write_end(&w);
    else ae_byte_align(&w);
    free(c0);
    if (w.overflow) return -1;
    return (int64_t)ae_writer_bytes(&w);
}

size_t ae_sizeof_writer(void) { return sizeof(ae_writer); }
