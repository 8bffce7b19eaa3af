/*
 * alac_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the saprobe-alac packet-decode path. It exists to CHECK the
 * CUDA product path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * legs). Nothing under saprobe-alac_b200/ may include, link or call it.
 *
 * Pinning: the reference ships no golden vectors and its Go toolchain is
 * absent from this image, so the oracle is pinned against an independent ALAC encoder+decoder
 * (FFmpeg 8 libavcodec, driven by tests/golden/gen_ffmpeg_fixtures.py): on every committed fixture
 *   oracle(packets) == FFmpeg-decode(packets) == source PCM
 * which is exactly what the reference's own conformance suite asserts
 * (tests/conformance_test.go:282-332).  Paths FFmpeg cannot emit (20/32-bit, mode!=0, order 31,
 * DSE/FIL, ...) are pinned only by this restatement ("parity = restatement only" in test names).
 *
 * Reference files restated (all under /root/reference):
 *   internal/alac/bitbuffer.go:36-123   bit reader
 *   internal/alac/golomb.go:55-253      adaptive Golomb-Rice (DynDecomp, dynGet, getStreamBits)
 *   internal/alac/predictor.go:35-684   sign-LMS predictor (UnpcBlock + 4/5/6/8/general)
 *   internal/alac/matrix.go:30-301      un-mix + shift merge + LE PCM emit
 *   internal/alac/format.go:23-34       BytesPerSample
 *   decoder.go:55-64, 133-574           element grammar
 *   config.go:47-81                     magic cookie
 */
#ifndef ALAC_ORACLE_H
#define ALAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status word. Low byte = sentinel (1:1 with internal/alac/errors.go:24-33, plus two additions),
 * bits 8-11 = element context of decoder.go's error wrapping, bits 12-13 = entropy context. */
enum {
    AO_OK = 0,
    AO_ERR_INVALID_COOKIE = 1,      /* ErrInvalidCookie      errors.go:25 */
    AO_ERR_UNSUPPORTED_VERSION = 2, /* ErrUnsupportedVersion errors.go:26 */
    AO_ERR_UNSUPPORTED_ELEMENT = 3, /* ErrUnsupportedElement errors.go:27 */
    AO_ERR_INVALID_HEADER = 4,      /* ErrInvalidHeader      errors.go:28 */
    AO_ERR_INVALID_SHIFT = 5,       /* ErrInvalidShift       errors.go:29 */
    AO_ERR_BITSTREAM_OVERRUN = 6,   /* ErrBitstreamOverrun   errors.go:30 */
    AO_ERR_SAMPLE_OVERRUN = 7,      /* ErrSampleOverrun      errors.go:31 */
    AO_ERR_BIT_DEPTH = 8,           /* ErrBitDepth           errors.go:32 */
    AO_ERR_REF_PANIC = 9,           /* the Go reference would panic (index/slice out of range) */
    AO_ERR_UNSUPPORTED_CONFIG = 10  /* cookie the replacement refuses (channels not 1..8, frame length 0 or > 65536) */
};
enum { AO_CTX_NONE = 0, AO_CTX_SCE = 1, AO_CTX_CPE = 2, AO_CTX_DSE = 3, AO_CTX_FIL = 4 };
enum { AO_ENT_NONE = 0, AO_ENT_MONO = 1, AO_ENT_U = 2, AO_ENT_V = 3 };
#define AO_STATUS(code, ctx, ent) ((int32_t)((code) | ((ctx) << 8) | ((ent) << 12)))
#define AO_CODE(status) ((status) & 0xff)

typedef struct {
    uint32_t frame_length;
    uint8_t bit_depth;
    uint8_t num_channels;
    uint8_t pb;
    uint8_t mb;
    uint8_t kb;
    uint8_t pad_;
    uint16_t max_run;
    uint32_t max_frame_bytes;
    uint32_t avg_bit_rate;
    uint32_t sample_rate;
} ao_config; /* field-for-field PacketConfig, config.go:27-38 */

/* config.go:47-81 */
int32_t ao_parse_cookie(const uint8_t *cookie, size_t len, ao_config *out);
/* decoder.go:90-93 plus the replacement's own limits */
int32_t ao_check_config(const ao_config *cfg);
/* internal/alac/format.go:23-34; 0 for an unsupported depth */
int ao_bytes_per_sample(uint8_t depth);

/* decoder.go:117-128 DecodePacket on a FRESH PacketDecoder. `out` must hold
 * frame_length*num_channels*bps bytes; it is zero-filled first (fresh make(), decoder.go:120).
 * *out_bytes = numSamples*numChannels*bps on success, 0 on error. Returns the status word. */
int32_t ao_decode_packet(const ao_config *cfg, const uint8_t *packet, size_t size, uint8_t *out,
                         uint32_t *out_bytes);

/* Batch driver used as the CPU baseline: packets i at packed+offsets[i] (sizes[i] bytes) decoded
 * into out + i*out_stride by `nthreads` threads over contiguous packet ranges. */
void ao_decode_batch(const ao_config *cfg, const uint8_t *packed, const uint64_t *offsets,
                     const uint32_t *sizes, uint32_t n, uint8_t *out, uint64_t out_stride,
                     uint32_t *out_bytes, int32_t *status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
