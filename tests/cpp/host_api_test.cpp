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
    // PacketDecoder::DecodePackets ([]byte in, a fresh []byte per packet out) over the file's own packets: the
    // concatenation is the file's PCM (the partial last packet comes back shorter, decoder.go:127)
    alacb200_track *trk = nullptr;
    CHECK(alacb200_mp4_find_alac_track(file.data(), file.size(), &trk) == ALACB200_OK);
    size_t clen = 0;
    const uint8_t *cookie = alacb200_mp4_cookie(trk, &clen);
    uint64_t ns = 0;
    const alacb200_sample_info *si = alacb200_mp4_samples(trk, &ns);
    CHECK(ns > 2);
    std::vector<std::pair<const uint8_t *, size_t>> packets;
    std::vector<uint64_t> offsets(ns);
    std::vector<uint32_t> sizes(ns);
    for (uint64_t k = 0; k < ns; k++) {
        packets.push_back({file.data() + si[k].offset, si[k].size});
        offsets[k] = si[k].offset;
        sizes[k] = si[k].size;
    }
    auto pdec = alac::PacketDecoder::New(alac::ParseMagicCookie(cookie, clen));
    auto res = pdec->DecodePackets(packets);
    std::vector<uint8_t> cat;
    for (auto &r : res) {
        CHECK(!r.err);
        cat.insert(cat.end(), r.pcm.begin(), r.pcm.end());
    }
    CHECK(cat == want);
    CHECK(pdec->DecodePacket(packets[1].first, packets[1].second) == res[1].pcm);  // decoder.go:117-128, one packet
    // a garbage packet among good ones: its error, the others untouched
    std::vector<uint8_t> junk(64, 0xA5);
    auto mixed = pdec->DecodePackets({packets[0], {junk.data(), junk.size()}, packets[2]});
    CHECK(!mixed[0].err && mixed[0].pcm == res[0].pcm && mixed[1].err && mixed[1].pcm.empty() && !mixed[2].err && mixed[2].pcm == res[2].pcm);
    CHECK(mixed[1].err->kind == alac::ErrKind::Decode);
    // LibraryDecoder: the same track twice plus a track with a bad cookie, read in place from the file image
    std::vector<uint8_t> bad_cookie(cookie, cookie + clen);
    bad_cookie[bad_cookie.size() - 24 + 4] = 1;  // compatibleVersion != 0 (config.go:66-68)
    auto lib = alac::LibraryDecoder::New({0});
    std::vector<alac::TrackInput> tracks(3, alac::TrackInput{cookie, clen, file.data(), file.size(), offsets.data(), sizes.data(), (uint32_t)ns});
    tracks[1].cookie = bad_cookie.data();
    auto outs = lib->DecodeTracks(tracks);
    CHECK(outs.size() == 3 && !outs[0].err && outs[1].err && outs[1].err->kind == alac::ErrKind::Config && !outs[2].err);
    for (int t : {0, 2}) {
        std::vector<uint8_t> lcat;
        for (uint64_t k = 0; k < ns; k++) {
            CHECK(outs[t].status[k] == ALACB200_ST_OK);
            lcat.insert(lcat.end(), outs[t].pcm.begin() + k * outs[t].stride, outs[t].pcm.begin() + k * outs[t].stride + outs[t].out_bytes[k]);
        }
        CHECK(lcat == want);
    }
    alacb200_mp4_free_track(trk);
    std::printf("gpu ok: %zu bytes, seek at frame %lld, DecodePackets and LibraryDecoder agree\n", got.size(), (long long)frame);
    return 0;
}
