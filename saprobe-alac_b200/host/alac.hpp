// alac.hpp -- C++ host mirror of saprobe-alac's exported Go API over the C ABI (include/alac_b200.h).
//
// The reference's host language is Go; this image has no Go toolchain, so the compiled host side above
// the C ABI is C++ (header-only). Names, argument meaning and error behaviour follow the Go package:
//
//   alac::ParseMagicCookie            config.go:47-81
//   alac::PacketDecoder               decoder.go:79-128   (NewPacketDecoder, DecodePacket, Format)
//   alac::PacketDecoder::DecodePackets  NEW: the batched entry point of the north star
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
#include <utility>
#include <vector>

#include "../../include/alac_b200.h"

namespace alac {

using PacketConfig = alacb200_config;  // PacketConfig, config.go:27-38

struct PCMFormat {  // format.go:20-24
    int SampleRate, BitDepth, Channels;
};

enum class ErrKind { Config, NoTrack, Decode, Device };  // ErrConfig / ErrNoTrack / ErrDecode (+ CUDA unusable)

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

    // DecodePackets: one batched GPU call for many packets (possibly of several tracks with this cookie).
    std::vector<PacketResult> DecodePackets(const std::vector<std::pair<const uint8_t *, size_t>> &packets) {
        const uint32_t n = (uint32_t)packets.size();
        std::vector<PacketResult> out(n);
        if (n == 0) return out;
        // host packer: contiguous buffer, every packet on a 16-byte boundary
        std::vector<uint64_t> offsets(n);
        std::vector<uint32_t> sizes(n);
        uint64_t pos = 0;
        for (uint32_t i = 0; i < n; i++) {
            offsets[i] = pos;
            sizes[i] = (uint32_t)packets[i].second;
            pos += (packets[i].second + 15) / 16 * 16;
        }
        std::vector<uint8_t> packed(pos + 64, 0);
        for (uint32_t i = 0; i < n; i++)
            if (sizes[i]) std::memcpy(packed.data() + offsets[i], packets[i].first, sizes[i]);
        const uint64_t stride = (alacb200_max_packet_pcm_bytes(h_) + 3) / 4 * 4;
        std::vector<uint8_t> pcm((size_t)n * stride);
        std::vector<uint32_t> nbytes(n);
        std::vector<int32_t> status(n);
        const int32_t rc = alacb200_decode_packets(h_, packed.data(), offsets.data(), sizes.data(), n, pcm.data(), stride,
                                                   nbytes.data(), status.data());
        if (rc != ALACB200_OK) throw Error(ErrKind::Device, 0, std::string("alacb200_decode_packets: ") + alacb200_last_error());
        for (uint32_t i = 0; i < n; i++) {
            if (status[i] == ALACB200_ST_OK) out[i].pcm.assign(pcm.begin() + (size_t)i * stride, pcm.begin() + (size_t)i * stride + nbytes[i]);
            else out[i].err.reset(new Error(error_from_status(status[i])));
        }
        return out;
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

    static std::unique_ptr<Decoder> New(std::vector<uint8_t> file, int device = 0, size_t window = 2048) {  // NewDecoder, decode.go:50-75
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
        d->file_ = std::move(file);
        d->samples_.assign(si, si + ns);
        d->dec_ = PacketDecoder::New(cfg, device);
        d->window_ = std::max<size_t>(1, window);
        return d;
    }

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
        const int64_t frame = (int64_t)(((double)t / 1e9) * (double)c.sample_rate);
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
        if (!(idx >= ready_base_ && idx < ready_base_ + ready_.size())) {
            const size_t hi = std::min(samples_.size(), idx + window_);
            std::vector<std::pair<const uint8_t *, size_t>> pk;
            bool short_read = false;
            for (size_t k = idx; k < hi; k++) {
                const auto &s = samples_[k];
                if (s.offset > file_.size() || s.size > file_.size() - s.offset) {  // io.ReadFull fails, decode.go:172-174
                    short_read = true;
                    break;
                }
                pk.emplace_back(file_.data() + s.offset, s.size);
            }
            ready_ = dec_->DecodePackets(pk);
            if (short_read) {
                PacketResult r;
                r.err.reset(new Error(ErrKind::Decode, 0, "reading sample " + std::to_string(idx + pk.size()) + ": unexpected EOF"));
                ready_.push_back(std::move(r));
            }
            ready_base_ = idx;
        }
        PacketResult &r = ready_[idx - ready_base_];
        if (r.err) throw Error(r.err->kind, r.err->status, "decoding packet " + std::to_string(idx) + ": " + r.err->what());  // decode.go:181
        buf_ = r.pcm;
        buf_off_ = 0;
        sample_idx_++;
    }

    std::vector<uint8_t> file_;
    std::vector<alacb200_sample_info> samples_;
    std::unique_ptr<PacketDecoder> dec_;
    size_t sample_idx_ = 0, window_ = 2048, ready_base_ = 0, buf_off_ = 0;
    std::vector<PacketResult> ready_;
    std::vector<uint8_t> buf_;
    bool eof_ = false;
};

}  // namespace alac
