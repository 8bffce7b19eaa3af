// Exercises the C++ host mirror (saprobe-alac_b200/host/alac.hpp) the way a user of the Go package would.
//   host_api_test cpu                      cookie / error / container logic, no kernel launch
//   host_api_test gpu file.m4a want.pcm    NewDecoder + Read (+ Seek) and DecodePackets on the GPU vs expected PCM
#include <cstdio>
#include <fstream>
#include <iterator>

#include "../../saprobe-alac_b200/host/alac.hpp"

static std::vector<uint8_t> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
#define CHECK(c)                                                         \
    do {                                                                 \
        if (!(c)) {                                                      \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    if (mode == "cpu") {
        // error_test.go:81-122: short / empty cookie -> ErrConfig
        try {
            alac::ParseMagicCookie(nullptr, 0);
            return 1;
        } catch (const alac::Error &e) {
            CHECK(e.kind == alac::ErrKind::Config);
            CHECK(std::string(e.what()) == "invalid configuration: alac: invalid magic cookie");
        }
        uint8_t ck[24] = {0, 0, 0x10, 0, 0, 16, 40, 10, 14, 2, 0, 255, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xac, 0x44};
        alac::PacketConfig cfg = alac::ParseMagicCookie(ck, sizeof ck);
        CHECK(cfg.frame_length == 4096 && cfg.bit_depth == 16 && cfg.num_channels == 2 && cfg.sample_rate == 44100);
        cfg.bit_depth = 12;  // error_test.go:126-142
        try {
            alac::PacketDecoder::New(cfg);
            return 1;
        } catch (const alac::Error &e) {
            CHECK(e.kind == alac::ErrKind::Config);
            CHECK(std::string(e.what()).find("unsupported bit depth: 12") != std::string::npos);
        }
        // error_test.go:146-172: empty / garbage reader -> ErrNoTrack
        try {
            alac::Decoder::New(std::vector<uint8_t>(64, 0x5a));
            return 1;
        } catch (const alac::Error &e) {
            CHECK(e.kind == alac::ErrKind::NoTrack);
        }
        if (alacb200_device_count() == 0) {  // no CPU fallback: creating a decoder must fail loudly
            cfg.bit_depth = 16;
            try {
                alac::PacketDecoder::New(cfg);
                return 1;
            } catch (const alac::Error &e) {
                CHECK(e.kind == alac::ErrKind::Device);
            }
        }
        std::puts("cpu ok");
        return 0;
    }
    CHECK(argc >= 4);
    auto file = slurp(argv[2]);
    auto want = slurp(argv[3]);
    auto dec = alac::Decoder::New(file, 0, 7);  // a small window so several GPU calls happen
    std::vector<uint8_t> got;
    std::vector<uint8_t> buf(10007);  // odd read size: short reads must stitch
    for (;;) {
        size_t n = dec->Read(buf.data(), buf.size());
        if (n == 0) break;
        got.insert(got.end(), buf.begin(), buf.begin() + n);
    }
    CHECK(got == want);
    // conformance_test.go:343-421: seek to 50 % and compare with the tail of the full decode
    const auto fmt = dec->Format();
    const int bps = alacb200_bytes_per_sample((uint8_t)fmt.BitDepth);
    const int64_t at = dec->Seek(dec->Duration() / 2);
    CHECK(at == dec->Position());
    const int64_t frame = at * fmt.SampleRate / 1000000000ll;
    std::vector<uint8_t> tail;
    for (;;) {
        size_t n = dec->Read(buf.data(), buf.size());
        if (n == 0) break;
        tail.insert(tail.end(), buf.begin(), buf.begin() + n);
    }
    const size_t off = (size_t)frame * fmt.Channels * bps;
    CHECK(off <= want.size() && tail == std::vector<uint8_t>(want.begin() + off, want.end()));
    CHECK(dec->Seek(dec->Duration() * 2) == dec->Duration() && dec->Read(buf.data(), 16) == 0);
    std::printf("gpu ok: %zu bytes, seek at frame %lld\n", got.size(), (long long)frame);
    return 0;
}
