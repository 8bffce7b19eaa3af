// alac_abi.cu -- the C ABI of include/alac_b200.h over the kernels in alac_kernels.cuh.
//
// Host-side plumbing only: handles, device buffers, the chunked H2D -> decode kernel -> D2H pipeline over
// four slots (a CUDA stream and its buffers each), pinned memory, error text. No decode arithmetic lives
// here and there is no CPU fallback: if CUDA is unusable every decode entry point fails.
#include "alac_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/alac_b200.h"

using namespace alacb200;

namespace {

thread_local std::string g_last_error;

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define CU(call)                                       \
    do {                                               \
        if (!cuda_ok((call), #call)) return ALACB200_E_CUDA; \
    } while (0)

constexpr int kSlots = 4;  // pipeline depth of the host-buffer path

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    // grow-only; contents are not preserved
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            g_last_error = "cudaMalloc failed";
            cudaGetLastError();
            return false;
        }
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            g_last_error = "cudaHostAlloc failed";
            cudaGetLastError();
            return false;
        }
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// What one in-flight batch needs besides its inputs and outputs: the parked samples and the element lists.
struct Work {
    DevBuf scratch;  // int32 [groups][channels][frame_length][32]
    DevBuf descs;    // PacketDesc [n]
};

struct Slot {
    cudaStream_t stream = nullptr;
    Work work;
    DevBuf packed, offsets, sizes, pcm, out_bytes, status;
    PinBuf h_in;   // pinned staging of this chunk's rebased offsets (u64) + sizes (u32)
    PinBuf h_out;  // pinned staging of this chunk's out_bytes (u32) + status (i32)
    cudaEvent_t done = nullptr;
    // where h_out goes once the chunk has finished (caller arrays may be pageable: copying into them straight
    // from the stream would make every chunk synchronous)
    uint32_t *user_out_bytes = nullptr;
    int32_t *user_status = nullptr;
    uint32_t pending = 0;
    uint64_t pcm_stride = 0;  // out_stride the gaps of `pcm` were last zeroed for
};

struct ProfEvents {
    cudaEvent_t e0, e1;
};

}  // namespace

struct alacb200_decoder {
    alacb200_config cfg;
    DevConfig dev_cfg;
    int device;
    uint64_t frame_bytes;
    Work device_path_work;  // scratch of alacb200_decode_packets_device
    Slot slots[kSlots];
    bool profiling = false;
    std::vector<ProfEvents> prof_events;
    alacb200_profile prof{};
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cuda_ok(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int32_t check_config(const alacb200_config *cfg) {
    if (alacb200_bytes_per_sample(cfg->bit_depth) == 0) return ALACB200_ST_BIT_DEPTH;  // decoder.go:91-93
    if (cfg->num_channels < 1 || cfg->num_channels > 8) return ALACB200_ST_UNSUPPORTED_CONFIG;
    if (cfg->frame_length < 1 || cfg->frame_length > 65536) return ALACB200_ST_UNSUPPORTED_CONFIG;
    return ALACB200_ST_OK;
}


// Enqueue the decode kernel for n device-resident packets on `stream`.
int32_t launch(alacb200_decoder *dec, Work &work, const uint8_t *d_packed, const uint64_t *d_offsets,
               const uint32_t *d_sizes, uint32_t n, uint8_t *d_pcm, uint64_t out_stride, uint32_t *d_out_bytes,
               int32_t *d_status, cudaStream_t stream) {
    if (n == 0) return ALACB200_OK;
    const DevConfig &c = dec->dev_cfg;
    // Bound the parked-sample scratch: very large batches (a library shard is millions of packets) run as a
    // sequence of launches on the same stream, each reusing the scratch of the one before it.
    const uint64_t per_group = (uint64_t)c.num_channels * c.frame_length * 32u * sizeof(int32_t);
    const uint32_t max_groups = (uint32_t)std::max<uint64_t>(64, (8ull << 30) / per_group);
    if ((n + 31u) / 32u > max_groups) {
        for (uint32_t a = 0; a < n; a += max_groups * 32u) {
            const uint32_t m = std::min(n - a, max_groups * 32u);
            int32_t rc = launch(dec, work, d_packed, d_offsets + a, d_sizes + a, m, d_pcm + (size_t)a * out_stride, out_stride,
                                d_out_bytes + a, d_status + a, stream);
            if (rc != ALACB200_OK) return rc;
        }
        return ALACB200_OK;
    }
    const uint32_t groups = (n + 31u) / 32u;
    const size_t scratch_bytes = (size_t)groups * c.num_channels * c.frame_length * 32u * sizeof(int32_t);
    if (scratch_bytes > work.scratch.cap || (size_t)n * sizeof(PacketDesc) > work.descs.cap) {
        // growing frees the old buffers: make sure nothing in flight still uses them
        CU(cudaDeviceSynchronize());
        if (!work.scratch.reserve(scratch_bytes) || !work.descs.reserve((size_t)n * sizeof(PacketDesc)))
            return ALACB200_E_NOMEM;
    }
    ProfEvents pe{};
    if (dec->profiling) {
        CU(cudaEventCreate(&pe.e0));
        CU(cudaEventCreate(&pe.e1));
        CU(cudaEventRecord(pe.e0, stream));
    }
    alac_decode_kernel<<<groups, DEC_THREADS, sizeof(DecShared), stream>>>(d_packed, d_offsets, d_sizes, n, c,
                                                                           (int32_t *)work.scratch.p,
                                                                           (PacketDesc *)work.descs.p, d_pcm, out_stride,
                                                                           d_out_bytes, d_status);
    CU(cudaGetLastError());
    if (dec->profiling) {
        CU(cudaEventRecord(pe.e1, stream));
        dec->prof_events.push_back(pe);
        dec->prof.launches_decode++;
    }
    return ALACB200_OK;
}

int32_t drain_profile(alacb200_decoder *dec) {
    for (auto &pe : dec->prof_events) {
        CU(cudaEventSynchronize(pe.e1));
        float a = 0;
        CU(cudaEventElapsedTime(&a, pe.e0, pe.e1));
        dec->prof.ms_decode += a;
        cudaEventDestroy(pe.e0);
        cudaEventDestroy(pe.e1);
    }
    dec->prof_events.clear();
    return ALACB200_OK;
}

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

}  // namespace

extern "C" {

int32_t alacb200_parse_cookie(const uint8_t *cookie, size_t len, alacb200_config *out) {
    if (!out) return ALACB200_ST_INVALID_COOKIE;
    std::memset(out, 0, sizeof(*out));
    const uint8_t *d = cookie;
    if (!d) len = 0;
    // optional 'frma' and 'alac' atom wrappers, config.go:50-58
    if (len >= 12 && std::memcmp(d + 4, "frma", 4) == 0) { d += 12; len -= 12; }
    if (len >= 12 && std::memcmp(d + 4, "alac", 4) == 0) { d += 12; len -= 12; }
    if (len < 24) return ALACB200_ST_INVALID_COOKIE;       // config.go:60-62
    if (d[4] > 0) return ALACB200_ST_UNSUPPORTED_VERSION;  // config.go:64-67
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
    return ALACB200_ST_OK;
}

int32_t alacb200_bytes_per_sample(uint8_t bit_depth) {
    switch (bit_depth) {
    case 16: return 2;
    case 20:
    case 24: return 3;
    case 32: return 4;
    default: return 0;
    }
}

int32_t alacb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t alacb200_create(const alacb200_config *cfg, int device, alacb200_decoder **out, int32_t *status_out) {
    if (status_out) *status_out = ALACB200_ST_OK;
    if (!cfg || !out) return ALACB200_E_ARG;
    *out = nullptr;
    int32_t st = check_config(cfg);
    if (st != ALACB200_ST_OK) {
        if (status_out) *status_out = st;
        return ALACB200_E_CONFIG;
    }
    int ndev = alacb200_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) {
        g_last_error = "no usable CUDA device (this library has no CPU fallback)";
        return ALACB200_E_NO_DEVICE;
    }
    DeviceGuard guard(device);
    if (!guard.ok) return ALACB200_E_CUDA;
    auto *dec = new alacb200_decoder();
    dec->cfg = *cfg;
    dec->device = device;
    dec->dev_cfg.frame_length = cfg->frame_length;
    dec->dev_cfg.bit_depth = cfg->bit_depth;
    dec->dev_cfg.num_channels = cfg->num_channels;
    dec->dev_cfg.bps = (uint32_t)alacb200_bytes_per_sample(cfg->bit_depth);
    dec->dev_cfg.pb = cfg->pb;
    dec->dev_cfg.mb = cfg->mb;
    dec->dev_cfg.kb = cfg->kb;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
        dec->dev_cfg.num_sms = (uint32_t)sms;
    }
    dec->frame_bytes = (uint64_t)cfg->frame_length * cfg->num_channels * dec->dev_cfg.bps;
    cudaError_t e = cudaFuncSetAttribute(alac_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecShared));
    if (!cuda_ok(e, "cudaFuncSetAttribute(alac_decode_kernel)")) {
        delete dec;
        return ALACB200_E_CUDA;
    }
    for (auto &s : dec->slots) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !cuda_ok(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate")) {
            alacb200_destroy(dec);
            return ALACB200_E_CUDA;
        }
    }
    *out = dec;
    return ALACB200_OK;
}

void alacb200_destroy(alacb200_decoder *dec) {
    if (!dec) return;
    DeviceGuard guard(dec->device);
    cudaDeviceSynchronize();
    for (auto &pe : dec->prof_events) {
        cudaEventDestroy(pe.e0);
        cudaEventDestroy(pe.e1);
    }
    dec->device_path_work.scratch.release();
    dec->device_path_work.descs.release();
    for (auto &s : dec->slots) {
        s.work.scratch.release();
        s.work.descs.release();
        s.packed.release();
        s.offsets.release();
        s.sizes.release();
        s.pcm.release();
        s.out_bytes.release();
        s.status.release();
        s.h_in.release();
        s.h_out.release();
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete dec;
}

int32_t alacb200_format(const alacb200_decoder *dec, alacb200_pcm_format *out) {
    if (!dec || !out) return ALACB200_E_ARG;
    out->sample_rate = (int32_t)dec->cfg.sample_rate;  // decoder.go:99-103
    out->bit_depth = dec->cfg.bit_depth;
    out->channels = dec->cfg.num_channels;
    return ALACB200_OK;
}

int32_t alacb200_get_config(const alacb200_decoder *dec, alacb200_config *out) {
    if (!dec || !out) return ALACB200_E_ARG;
    *out = dec->cfg;
    return ALACB200_OK;
}

uint64_t alacb200_max_packet_pcm_bytes(const alacb200_decoder *dec) { return dec ? dec->frame_bytes : 0; }

int32_t alacb200_decode_packets_device(alacb200_decoder *dec, const uint8_t *d_packed, uint64_t packed_bytes,
                                       const uint64_t *d_offsets, const uint32_t *d_sizes, uint32_t n,
                                       uint8_t *d_pcm_out, uint64_t out_stride, uint32_t *d_out_bytes,
                                       int32_t *d_status, void *stream) {
    (void)packed_bytes;
    if (!dec) return ALACB200_E_ARG;
    if (n == 0) return ALACB200_OK;
    if (!d_packed || !d_offsets || !d_sizes || !d_pcm_out || !d_out_bytes || !d_status) return ALACB200_E_ARG;
    if (out_stride < dec->frame_bytes || (out_stride & 3u) || (((uintptr_t)d_pcm_out) & 3u) ||
        (((uintptr_t)d_packed) & 15u)) {
        g_last_error = "out_stride must be >= max_packet_pcm_bytes and a multiple of 4; d_packed 16-byte aligned";
        return ALACB200_E_ARG;
    }
    DeviceGuard guard(dec->device);
    if (!guard.ok) return ALACB200_E_CUDA;
    return launch(dec, dec->device_path_work, d_packed, d_offsets, d_sizes, n, d_pcm_out, out_stride, d_out_bytes,
                  d_status, (cudaStream_t)stream);
}

int32_t alacb200_decode_packets(alacb200_decoder *dec, const uint8_t *packed, const uint64_t *offsets,
                                const uint32_t *sizes, uint32_t n, uint8_t *pcm_out, uint64_t out_stride,
                                uint32_t *out_bytes, int32_t *status) {
    if (!dec) return ALACB200_E_ARG;
    if (n == 0) return ALACB200_OK;
    if (!packed || !offsets || !sizes || !pcm_out || !out_bytes || !status) return ALACB200_E_ARG;
    if (out_stride < dec->frame_bytes || (out_stride & 3u)) {
        g_last_error = "out_stride must be >= max_packet_pcm_bytes and a multiple of 4";
        return ALACB200_E_ARG;
    }
    DeviceGuard guard(dec->device);
    if (!guard.ok) return ALACB200_E_CUDA;

    // Chunk the batch so copies of chunk k+1 / k-1 overlap the kernels of chunk k.
    uint32_t nchunks_target = 6;
    if (const char *env = std::getenv("ALACB200_CHUNKS")) nchunks_target = (uint32_t)std::max(1, std::atoi(env));  // tuning knob
    uint32_t chunk = (n + nchunks_target - 1u) / nchunks_target;
    chunk = std::min(std::max(chunk, 512u), 16384u);
    const uint64_t max_chunk_pcm = 512ull << 20;
    while (chunk > 32u && (uint64_t)chunk * out_stride > max_chunk_pcm) chunk /= 2u;
    chunk = (chunk + 31u) & ~31u;

    // Whatever way this call ends, nothing may still be copying into the caller's buffers afterwards and no slot may
    // keep pointers into the caller's arrays for a later call to write through.
    struct Drain {
        alacb200_decoder *d;
        bool clean = false;  // every slot was retired: nothing in flight
        ~Drain() {
            for (auto &s : d->slots) {
                if (!clean) cudaStreamSynchronize(s.stream);
                s.pending = 0;
                s.user_out_bytes = nullptr;
                s.user_status = nullptr;
            }
        }
    } drain{dec};

    int32_t rc = ALACB200_OK;
    uint32_t slot_idx = 0;
    auto retire = [](Slot &s) -> bool {  // wait for the slot's chunk and hand its per-packet results to the caller
        if (cudaEventSynchronize(s.done) != cudaSuccess) return false;
        if (s.pending) {
            const uint32_t *ob = (const uint32_t *)s.h_out.p;
            std::memcpy(s.user_out_bytes, ob, (size_t)s.pending * 4);
            std::memcpy(s.user_status, ob + s.pending, (size_t)s.pending * 4);
            s.pending = 0;
        }
        return true;
    };
    for (uint32_t a = 0; a < n && rc == ALACB200_OK; a += chunk, slot_idx = (slot_idx + 1) % kSlots) {
        const uint32_t b = std::min(n, a + chunk), m = b - a;
        Slot &s = dec->slots[slot_idx];
        if (!retire(s)) return ALACB200_E_CUDA;  // the slot's previous chunk (and its staging) is finished
        // byte range of this chunk inside `packed`
        uint64_t lo = UINT64_MAX, hi = 0;
        for (uint32_t i = a; i < b; i++) {
            lo = std::min(lo, offsets[i]);
            hi = std::max(hi, offsets[i] + sizes[i]);
        }
        if (hi < lo) lo = hi = 0;
        const uint32_t mis = (uint32_t)(lo & 15u);  // keep each packet's alignment relative to 16 bytes
        const uint64_t span = hi - lo;
        if (!s.packed.reserve(mis + span + 64) || !s.offsets.reserve((size_t)m * 8) || !s.sizes.reserve((size_t)m * 4) ||
            !s.out_bytes.reserve((size_t)m * 8) || !s.h_in.reserve((size_t)m * 12) || !s.h_out.reserve((size_t)m * 8))
            return ALACB200_E_NOMEM;
        if ((size_t)m * out_stride > s.pcm.cap || s.pcm_stride != out_stride) {
            if (!s.pcm.reserve((size_t)m * out_stride)) return ALACB200_E_NOMEM;
            // the kernel never touches the gap between frame_bytes and out_stride, and the gap travels back with the
            // slot: define it once per buffer and per stride (PCM of an earlier call must not show up in it)
            CU(cudaMemsetAsync(s.pcm.p, 0, s.pcm.cap, s.stream));
            s.pcm_stride = out_stride;
        }
        uint64_t *ho = (uint64_t *)s.h_in.p;
        uint32_t *hs = (uint32_t *)(ho + m);
        for (uint32_t i = 0; i < m; i++) ho[i] = offsets[a + i] - lo + mis;
        std::memcpy(hs, sizes + a, (size_t)m * 4);
        uint32_t *d_ob = (uint32_t *)s.out_bytes.p;  // out_bytes[m] then status[m], one D2H copy
        int32_t *d_st = (int32_t *)(d_ob + m);
        if (span) CU(cudaMemcpyAsync((uint8_t *)s.packed.p + mis, packed + lo, span, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.offsets.p, ho, (size_t)m * 8, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.sizes.p, hs, (size_t)m * 4, cudaMemcpyHostToDevice, s.stream));
        rc = launch(dec, s.work, (const uint8_t *)s.packed.p, (const uint64_t *)s.offsets.p, (const uint32_t *)s.sizes.p, m,
                    (uint8_t *)s.pcm.p, out_stride, d_ob, d_st, s.stream);
        if (rc != ALACB200_OK) break;
        CU(cudaMemcpyAsync(pcm_out + (size_t)a * out_stride, s.pcm.p, (size_t)(m - 1) * out_stride + dec->frame_bytes,
                           cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_out.p, d_ob, (size_t)m * 8, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.done, s.stream));
        s.user_out_bytes = out_bytes + a;
        s.user_status = status + a;
        s.pending = m;
    }
    for (auto &s : dec->slots)
        if (!retire(s)) return ALACB200_E_CUDA;
    drain.clean = rc == ALACB200_OK;
    return rc;
}

void *alacb200_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        g_last_error = "cudaHostAlloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void alacb200_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
}

const char *alacb200_strerror(int32_t status) {
    switch (ALACB200_ST_CODE(status)) {  // strings of internal/alac/errors.go:24-33
    case ALACB200_ST_OK: return "ok";
    case ALACB200_ST_INVALID_COOKIE: return "alac: invalid magic cookie";
    case ALACB200_ST_UNSUPPORTED_VERSION: return "alac: unsupported compatible version";
    case ALACB200_ST_UNSUPPORTED_ELEMENT: return "alac: unsupported element type (CCE/PCE)";
    case ALACB200_ST_INVALID_HEADER: return "alac: invalid frame header";
    case ALACB200_ST_INVALID_SHIFT: return "alac: invalid bytesShifted value";
    case ALACB200_ST_BITSTREAM_OVERRUN: return "alac: bitstream overrun";
    case ALACB200_ST_SAMPLE_OVERRUN: return "alac: sample count exceeds buffer";
    case ALACB200_ST_BIT_DEPTH: return "alac: unsupported bit depth";
    case ALACB200_ST_REF_PANIC: return "alac: malformed packet (the reference decoder would panic)";
    case ALACB200_ST_UNSUPPORTED_CONFIG: return "alac: unsupported channel count or frame length";
    default: return "alac: unknown status";
    }
}

size_t alacb200_format_error(int32_t status, char *buf, size_t cap) {
    if (!buf || cap == 0) return 0;
    const int code = ALACB200_ST_CODE(status);
    if (code == ALACB200_ST_OK) return (size_t)std::snprintf(buf, cap, "ok");
    const bool is_config = code == ALACB200_ST_INVALID_COOKIE || code == ALACB200_ST_UNSUPPORTED_VERSION ||
                           code == ALACB200_ST_BIT_DEPTH || code == ALACB200_ST_UNSUPPORTED_CONFIG;
    static const char *ctx[] = {"", "SCE/LFE: ", "CPE: ", "DSE: ", "FIL: "};
    static const char *ent[] = {"", "entropy decode: ", "entropy decode U: ", "entropy decode V: "};
    const int c = ALACB200_ST_CTX(status), e = ALACB200_ST_ENT(status);
    int w = std::snprintf(buf, cap, "%s: %s%s%s", is_config ? "invalid configuration" : "decode failed",
                          c <= 4 ? ctx[c] : "", ent[e], alacb200_strerror(status));
    return w < 0 ? 0 : (size_t)w;
}

const char *alacb200_last_error(void) { return g_last_error.c_str(); }

#ifdef ALACB200_DEV
// Developer hooks (only in the `make dev` build, not part of include/alac_b200.h): point the decode kernel's per-role
// clock64 counters at a device buffer of n_ctas*16 u64 (or NULL to switch them off). Used by tools/role_cycles.py.
int32_t alacb200_debug_flags(unsigned int flags) {
    CU(cudaMemcpyToSymbol(g_debug_flags, &flags, sizeof(flags)));
    return ALACB200_OK;
}

int32_t alacb200_debug_role_cycles(unsigned long long *d_buf) {
    CU(cudaMemcpyToSymbol(g_role_cycles, &d_buf, sizeof(d_buf)));
    return ALACB200_OK;
}
#endif

int32_t alacb200_set_profiling(alacb200_decoder *dec, int enable) {
    if (!dec) return ALACB200_E_ARG;
    DeviceGuard guard(dec->device);
    if (!guard.ok) return ALACB200_E_CUDA;
    int32_t rc = drain_profile(dec);
    if (rc != ALACB200_OK) return rc;
    dec->profiling = enable != 0;
    if (enable) dec->prof = alacb200_profile{};
    return ALACB200_OK;
}

int32_t alacb200_get_profile(alacb200_decoder *dec, alacb200_profile *out) {
    if (!dec || !out) return ALACB200_E_ARG;
    DeviceGuard guard(dec->device);
    if (!guard.ok) return ALACB200_E_CUDA;
    int32_t rc = drain_profile(dec);
    if (rc != ALACB200_OK) return rc;
    *out = dec->prof;
    return ALACB200_OK;
}

}  // extern "C"
