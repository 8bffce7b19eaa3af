// alac.hpp -- C++ host mirror of saprobe-alac's exported Go API over the C ABI (include/alac_b200.h).
//
// The reference's host language is Go; this image has no Go toolchain, so the compiled host side above
// the C ABI is C++ (header-only). Names, argument meaning and error behaviour follow the Go package:
//
//   alac::ParseMagicCookie            config.go:47-81
//   alac::PacketDecoder               decoder.go:79-128   (NewPacketDecoder, DecodePacket, Format)
//   alac::PacketDecoder::DecodePackets  NEW: the batched entry point of the north star
//   alac::LibraryDecoder              NEW: many tracks of mixed cookies in one call, sharded over devices
//   alac::Decoder                     decode.go:32-190    (NewDecoder, Read, Seek, Duration, Position, Format)
//   alac::Error{Config,NoTrack,Decode}  errors.go:22-34, message = the reference's %w chain
//
// All decoding runs in libalacb200.so (CUDA); nothing here decodes on the CPU.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/alac_b200.h"

namespace alac {

using PacketConfig = alacb200_config;  // PacketConfig, config.go:27-38

struct PCMFormat {  // format.go:20-24
    int SampleRate, BitDepth, Channels;
};

enum class ErrKind { Config, NoTrack, Decode, Device, Read };  // ErrConfig / ErrNoTrack / ErrDecode (+ CUDA unusable, reader failure)

class Error : public std::runtime_error {
public:
    Error(ErrKind k, int32_t status, const std::string &msg) : std::runtime_error(msg), kind(k), status(status) {}
    ErrKind kind;
    int32_t status;  // status word of the C ABI (0 for container / device errors)
};

inline Error error_from_status(int32_t status, const std::string &prefix = "") {
    char buf[256];
    alacb200_format_error(status, buf, sizeof buf);
    const int code = ALACB200_ST_CODE(status);
    const bool cfg = code == ALACB200_ST_INVALID_COOKIE || code == ALACB200_ST_UNSUPPORTED_VERSION ||
                     code == ALACB200_ST_BIT_DEPTH || code == ALACB200_ST_UNSUPPORTED_CONFIG;
    return Error(cfg ? ErrKind::Config : ErrKind::Decode, status, prefix + buf);
}

inline PacketConfig ParseMagicCookie(const uint8_t *cookie, size_t len) {
    PacketConfig cfg;
    const int32_t st = alacb200_parse_cookie(cookie, len, &cfg);
    if (st != ALACB200_ST_OK) throw error_from_status(st);
    return cfg;
}
inline PacketConfig ParseMagicCookie(const std::vector<uint8_t> &cookie) { return ParseMagicCookie(cookie.data(), cookie.size()); }

// The host side of a big batch is memcpy-bound (packing [][]byte into the pinned arena, copying every packet's PCM into
// its own fresh buffer): split it over a few host threads. fn(lo, hi) works on the index range [lo, hi).
template <class F>
inline void parallel_ranges(size_t n, size_t bytes, F &&fn) {
    size_t nt = std::min<size_t>({8, std::max<size_t>(1, std::thread::hardware_concurrency()), bytes / (4u << 20) + 1, n ? n : 1});
    if (nt <= 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back([&, t] { fn(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

struct PacketResult {
    std::vector<uint8_t> pcm;     // empty on error
    std::unique_ptr<Error> err;   // nullptr on success
};

class PacketDecoder {
public:
    // NewPacketDecoder, decoder.go:90-110
    static std::unique_ptr<PacketDecoder> New(const PacketConfig &config, int device = 0) {
        alacb200_decoder *h = nullptr;
        int32_t st = 0;
        const int32_t rc = alacb200_create(&config, device, &h, &st);
        if (rc == ALACB200_E_CONFIG) {
            Error e = error_from_status(st);
            if (st == ALACB200_ST_BIT_DEPTH) throw Error(e.kind, st, std::string(e.what()) + ": " + std::to_string(config.bit_depth));
            throw e;
        }
        if (rc != ALACB200_OK) throw Error(ErrKind::Device, 0, std::string("alacb200_create: ") + alacb200_last_error());
        return std::unique_ptr<PacketDecoder>(new PacketDecoder(h, config));
    }
    ~PacketDecoder() { alacb200_destroy(h_); }
    PacketDecoder(const PacketDecoder &) = delete;
    PacketDecoder &operator=(const PacketDecoder &) = delete;

    PCMFormat Format() const { return PCMFormat{(int)cfg_.sample_rate, cfg_.bit_depth, cfg_.num_channels}; }  // decoder.go:112
    const PacketConfig &Config() const { return cfg_; }

    // DecodePackets: one batched GPU call for many packets (possibly of several tracks with this cookie). The packets
    // are packed into the decoder's own pinned arena and the PCM comes back through it: no pinned allocation per call.
    std::vector<PacketResult> DecodePackets(const std::vector<std::pair<const uint8_t *, size_t>> &packets) {
        const uint32_t n = (uint32_t)packets.size();
        std::vector<PacketResult> out(n);
        if (n == 0) return out;
        std::vector<uint64_t> offsets(n);
        std::vector<uint32_t> sizes(n);
        uint64_t pos = 0;
        for (uint32_t i = 0; i < n; i++) {  // host packer: every packet on a 16-byte boundary
            offsets[i] = pos;
            sizes[i] = (uint32_t)packets[i].second;
            pos += (packets[i].second + 15) / 16 * 16;
        }
        const uint64_t stride = Stride();
        uint8_t *in = nullptr, *pcm = nullptr;
        if (alacb200_arena(h_, pos, (uint64_t)n * stride, &in, &pcm) != ALACB200_OK)
            throw Error(ErrKind::Device, 0, std::string("alacb200_arena: ") + alacb200_last_error());
        parallel_ranges(n, pos, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++)
                if (sizes[i]) std::memcpy(in + offsets[i], packets[i].first, sizes[i]);
        });
        std::vector<uint32_t> nbytes(n);
        std::vector<int32_t> status(n);
        DecodeInPlace(in, pos, offsets.data(), sizes.data(), n, pcm, nbytes.data(), status.data());
        parallel_ranges(n, (size_t)n * stride, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                if (status[i] == ALACB200_ST_OK) out[i].pcm.assign(pcm + i * stride, pcm + i * stride + nbytes[i]);
                else out[i].err.reset(new Error(error_from_status(status[i])));
            }
        });
        return out;
    }

    // Packets read in place from `data` (a file image with its sample table, or packed packets): PCM rows of Stride()
    // bytes into pcm. status ALACB200_ST_IO_TRUNCATED marks a packet outside data (a READ error, decode.go:172-174).
    void DecodeInPlace(const uint8_t *data, uint64_t data_len, const uint64_t *offsets, const uint32_t *sizes, uint32_t n, uint8_t *pcm,
                       uint32_t *nbytes, int32_t *status) {
        const int32_t rc = alacb200_decode_packets(h_, data, data_len, offsets, sizes, n, pcm, Stride(), nbytes, status);
        if (rc != ALACB200_OK) throw Error(ErrKind::Device, 0, std::string("alacb200_decode_packets: ") + alacb200_last_error());
    }
    uint64_t Stride() const { return (alacb200_max_packet_pcm_bytes(h_) + 3) / 4 * 4; }
    // the decoder's pinned arena (grow-only, valid until the next Arena / DecodePackets call)
    void Arena(uint64_t in_bytes, uint64_t out_bytes, uint8_t **in, uint8_t **out) {
        if (alacb200_arena(h_, in_bytes, out_bytes, in, out) != ALACB200_OK)
            throw Error(ErrKind::Device, 0, std::string("alacb200_arena: ") + alacb200_last_error());
    }

    // DecodePacket, decoder.go:117-128: a fresh buffer of numSamples*channels*bps bytes, or the error.
    std::vector<uint8_t> DecodePacket(const uint8_t *packet, size_t len) {
        auto r = DecodePackets({{packet, len}});
        if (r[0].err) throw *r[0].err;
        return std::move(r[0].pcm);
    }

private:
    PacketDecoder(alacb200_decoder *h, const PacketConfig &cfg) : h_(h), cfg_(cfg) {}
    alacb200_decoder *h_;
    PacketConfig cfg_;
};

// Streaming decoder over an in-memory M4A/MP4 image, decode.go:32-190. Read decodes a window of packets per
// GPU call and serves bytes from it; packet order, short reads, EOF, Seek's packet alignment and Duration's
// over-count of a partial last packet are the reference's.
class Decoder {
public:
    using Duration_ns = int64_t;

    static std::unique_ptr<Decoder> New(const std::vector<uint8_t> &file, int device = 0, size_t window = 2048) {  // NewDecoder, decode.go:50-75
        alacb200_track *t = nullptr;
        const int32_t rc = alacb200_mp4_find_alac_track(file.data(), file.size(), &t);
        std::unique_ptr<alacb200_track, void (*)(alacb200_track *)> guard(t, alacb200_mp4_free_track);
        if (rc != ALACB200_OK) throw Error(ErrKind::NoTrack, 0, std::string("no track found: ") + alacb200_mp4_error(t));  // decode.go:53
        size_t clen = 0;
        const uint8_t *cookie = alacb200_mp4_cookie(t, &clen);
        PacketConfig cfg;
        try {
            cfg = ParseMagicCookie(cookie, clen);
        } catch (const Error &e) {
            throw Error(e.kind, e.status, std::string("parsing ALAC config: ") + e.what());  // decode.go:58
        }
        uint64_t ns = 0;
        const alacb200_sample_info *si = alacb200_mp4_samples(t, &ns);
        std::unique_ptr<Decoder> d(new Decoder());
        d->samples_.assign(si, si + ns);
        d->offsets_.resize(ns);
        d->sizes_.resize(ns);
        for (uint64_t k = 0; k < ns; k++) {
            d->offsets_[k] = si[k].offset;
            d->sizes_[k] = si[k].size;
        }
        d->dec_ = PacketDecoder::New(cfg, device);
        d->window_ = std::max<size_t>(1, window);
        // the file image lives in pinned memory for the life of the decoder: every window is read in place from it
        // (image + sample table straight to the C ABI, asynchronous H2D), nothing is re-packed
        d->image_len_ = file.size();
        d->image_ = (uint8_t *)alacb200_pinned_alloc(std::max<size_t>(1, file.size()));
        if (!d->image_) throw Error(ErrKind::Device, 0, std::string("alacb200_pinned_alloc: ") + alacb200_last_error());
        std::memcpy(d->image_, file.data(), file.size());
        return d;
    }
    ~Decoder() {
        dec_.reset();
        alacb200_pinned_free(image_);
    }
    Decoder(const Decoder &) = delete;
    Decoder &operator=(const Decoder &) = delete;

    PCMFormat Format() const { return dec_->Format(); }
    Duration_ns Duration() const {  // decode.go:82-88
        const auto &c = dec_->Config();
        return (int64_t)samples_.size() * c.frame_length * 1000000000ll / c.sample_rate;
    }
    Duration_ns Position() const {  // decode.go:91-97
        const auto &c = dec_->Config();
        return (int64_t)sample_idx_ * c.frame_length * 1000000000ll / c.sample_rate;
    }
    Duration_ns Seek(Duration_ns t) {  // decode.go:103-124
        const auto &c = dec_->Config();
        // time.Duration.Seconds(): whole seconds plus the nanosecond rest, each converted on its own
        const double seconds = (double)(t / 1000000000ll) + (double)(t % 1000000000ll) / 1e9;
        const int64_t frame = (int64_t)(seconds * (double)c.sample_rate);
        int64_t target = frame / (int64_t)c.frame_length;
        target = std::max<int64_t>(0, std::min<int64_t>(target, (int64_t)samples_.size()));
        sample_idx_ = (size_t)target;
        buf_.clear();
        buf_off_ = 0;
        eof_ = sample_idx_ >= samples_.size();
        return Position();
    }
    // io.Reader: returns bytes copied; 0 at EOF. Throws the decode / read error once the bytes before it are drained.
    size_t Read(uint8_t *p, size_t len) {  // decode.go:127-190
        size_t total = 0;
        while (len > 0) {
            if (buf_off_ < buf_.size()) {
                const size_t n = std::min(len, buf_.size() - buf_off_);
                std::memcpy(p, buf_.data() + buf_off_, n);
                buf_off_ += n;
                total += n;
                p += n;
                len -= n;
                continue;
            }
            if (eof_ || sample_idx_ >= samples_.size()) {
                eof_ = true;
                return total;
            }
            try {
                fill();
            } catch (...) {
                if (total > 0) return total;  // Go returns (total, err): the error surfaces on the next call
                throw;
            }
        }
        return total;
    }

private:
    Decoder() = default;
    void fill() {
        const size_t idx = sample_idx_;
        if (!(idx >= ready_base_ && idx < ready_base_ + ready_n_)) {
            const size_t hi = std::min(samples_.size(), idx + window_);
            const uint32_t n = (uint32_t)(hi - idx);
            const uint64_t stride = dec_->Stride();
            uint8_t *in = nullptr, *pcm = nullptr;
            dec_->Arena(0, (uint64_t)n * stride, &in, &pcm);
            ready_nbytes_.assign(n, 0);
            ready_status_.assign(n, 0);
            dec_->DecodeInPlace(image_, image_len_, offsets_.data() + idx, sizes_.data() + idx, n, pcm, ready_nbytes_.data(), ready_status_.data());
            ready_pcm_ = pcm;  // the arena is only reused by this decoder's next window
            ready_stride_ = stride;
            ready_base_ = idx;
            ready_n_ = n;
        }
        const size_t k = idx - ready_base_;
        const int32_t st = ready_status_[k];
        if (ALACB200_ST_CODE(st) == ALACB200_ST_IO_TRUNCATED)  // io.ReadFull fails, decode.go:172-174: not a decode error
            throw Error(ErrKind::Read, st, "reading sample " + std::to_string(idx) + ": unexpected EOF");
        if (st != ALACB200_ST_OK) {
            const Error e = error_from_status(st);
            throw Error(e.kind, e.status, "decoding packet " + std::to_string(idx) + ": " + e.what());  // decode.go:181
        }
        buf_.assign(ready_pcm_ + k * ready_stride_, ready_pcm_ + k * ready_stride_ + ready_nbytes_[k]);
        buf_off_ = 0;
        sample_idx_++;
    }

    uint8_t *image_ = nullptr;  // pinned copy of the file
    size_t image_len_ = 0;
    std::vector<alacb200_sample_info> samples_;
    std::vector<uint64_t> offsets_;
    std::vector<uint32_t> sizes_;
    std::unique_ptr<PacketDecoder> dec_;
    size_t sample_idx_ = 0, window_ = 2048, ready_base_ = 0, ready_n_ = 0, buf_off_ = 0;
    const uint8_t *ready_pcm_ = nullptr;
    uint64_t ready_stride_ = 0;
    std::vector<uint32_t> ready_nbytes_;
    std::vector<int32_t> ready_status_;
    std::vector<uint8_t> buf_;
    bool eof_ = false;
};

// Many tracks of mixed cookies in one call, sharded over the devices of the box (BASELINE configs[4]). A
// PacketDecoder holds one cookie (decoder.go:79-87); the library handle checks every track like ParseMagicCookie +
// NewPacketDecoder, balances contiguous track ranges over the devices by compressed bytes and reads every track's
// packets in place.
struct TrackInput {
    const uint8_t *cookie;
    size_t cookie_len;
    const uint8_t *data;  // file image or packed packets (pin it for asynchronous copies)
    uint64_t data_len;
    const uint64_t *offsets;
    const uint32_t *sizes;
    uint32_t n;
};
struct TrackOutput {
    PacketConfig config{};
    std::unique_ptr<Error> err;  // ErrConfig: the track has no decoder
    uint64_t stride = 0;
    std::vector<uint8_t> pcm;    // n rows of `stride` bytes; packet i is the first out_bytes[i] of row i
    std::vector<uint32_t> out_bytes;
    std::vector<int32_t> status;
    int device = -1;
};
class LibraryDecoder {
public:
    static std::unique_ptr<LibraryDecoder> New(const std::vector<int> &devices = {0}) {
        alacb200_library *h = nullptr;
        if (alacb200_library_create(devices.data(), (int)devices.size(), &h) != ALACB200_OK)
            throw Error(ErrKind::Device, 0, std::string("alacb200_library_create: ") + alacb200_last_error());
        return std::unique_ptr<LibraryDecoder>(new LibraryDecoder(h));
    }
    ~LibraryDecoder() { alacb200_library_destroy(h_); }
    LibraryDecoder(const LibraryDecoder &) = delete;
    LibraryDecoder &operator=(const LibraryDecoder &) = delete;

    std::vector<TrackOutput> DecodeTracks(const std::vector<TrackInput> &tracks) {
        const uint32_t nt = (uint32_t)tracks.size();
        std::vector<TrackOutput> out(nt);
        std::vector<alacb200_track_desc> descs(nt);
        for (uint32_t t = 0; t < nt; t++) {
            const TrackInput &in = tracks[t];
            alacb200_track_desc &d = descs[t];
            std::memset(&d, 0, sizeof d);
            PacketConfig cfg;
            uint64_t stride = 4;
            if (alacb200_parse_cookie(in.cookie, in.cookie_len, &cfg) == ALACB200_ST_OK && alacb200_bytes_per_sample(cfg.bit_depth) > 0)
                stride = ((uint64_t)cfg.frame_length * cfg.num_channels * alacb200_bytes_per_sample(cfg.bit_depth) + 3) / 4 * 4;
            out[t].stride = stride;
            out[t].pcm.assign((size_t)in.n * stride + 1, 0);
            out[t].out_bytes.assign(in.n + 1, 0);
            out[t].status.assign(in.n + 1, 0);
            d.cookie = in.cookie; d.cookie_len = in.cookie_len;
            d.data = in.data; d.data_len = in.data_len;
            d.offsets = in.offsets; d.sizes = in.sizes; d.n = in.n;
            d.pcm_out = out[t].pcm.data(); d.out_stride = stride;
            d.out_bytes = out[t].out_bytes.data(); d.status = out[t].status.data();
        }
        if (alacb200_library_decode_tracks(h_, descs.data(), nt) != ALACB200_OK)
            throw Error(ErrKind::Device, 0, std::string("alacb200_library_decode_tracks: ") + alacb200_last_error());
        for (uint32_t t = 0; t < nt; t++) {
            out[t].config = descs[t].config;
            out[t].device = descs[t].device;
            if (descs[t].result == ALACB200_E_CONFIG) {
                Error e = error_from_status(descs[t].track_status);
                if (descs[t].track_status == ALACB200_ST_BIT_DEPTH)
                    e = Error(e.kind, e.status, std::string(e.what()) + ": " + std::to_string(descs[t].config.bit_depth));  // decoder.go:92
                out[t].err.reset(new Error(e));
            } else if (descs[t].result != ALACB200_OK) {
                out[t].err.reset(new Error(ErrKind::Device, 0, "track " + std::to_string(t) + ": rc=" + std::to_string(descs[t].result)));
            }
        }
        return out;
    }

private:
    explicit LibraryDecoder(alacb200_library *h) : h_(h) {}
    alacb200_library *h_;
};

}  // namespace alac
