// api_bench -- what a caller of the drop-in API pays: the C++ host mirror (saprobe-alac_b200/host/alac.hpp, same names
// as the Go package) timed end to end, host buffers in and out. Used by bench.py for the `e2e_api` legs.
//   api_bench packets  cookie.bin packed.bin offsets.u64 sizes.u32 steps   PacketDecoder::DecodePackets, all packets per call
//   api_bench read     file.m4a steps                                      NewDecoder + Read to EOF in 32 KiB pieces (io.Copy)
// Prints one line: seconds_per_step pcm_bytes fnv1a32_of_the_pcm [seconds_per_step with one decoder kept open, then the
// NewDecoder / first Read / close shares of a step]
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>

#include "../saprobe-alac_b200/host/alac.hpp"

static std::vector<uint8_t> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static uint32_t fnv(uint32_t h, const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 16777619u;
    return h;
}
using clk = std::chrono::steady_clock;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string mode = argv[1];
    try {
        if (mode == "packets" && argc >= 7) {
            auto cookie = slurp(argv[2]), packed = slurp(argv[3]), offb = slurp(argv[4]), szb = slurp(argv[5]);
            const int steps = std::atoi(argv[6]);
            const size_t n = szb.size() / 4;
            const uint64_t *offs = reinterpret_cast<const uint64_t *>(offb.data());
            const uint32_t *sizes = reinterpret_cast<const uint32_t *>(szb.data());
            std::vector<std::pair<const uint8_t *, size_t>> packets(n);
            for (size_t i = 0; i < n; i++) packets[i] = {packed.data() + offs[i], sizes[i]};
            auto dec = alac::PacketDecoder::New(alac::ParseMagicCookie(cookie));
            size_t bytes = 0;
            uint32_t h = 2166136261u;
            for (int warm = 0; warm < 2; warm++) (void)dec->DecodePackets(packets);
            const auto t0 = clk::now();
            for (int s = 0; s < steps; s++) {
                auto res = dec->DecodePackets(packets);
                if (s == steps - 1) {
                    for (auto &r : res) {
                        if (r.err) throw *r.err;
                        bytes += r.pcm.size();
                        h = fnv(h, r.pcm.data(), r.pcm.size());
                    }
                }
            }
            const double dt = std::chrono::duration<double>(clk::now() - t0).count() / steps;
            std::printf("%.9f %zu %u\n", dt, bytes, h);
            return 0;
        }
        if (mode == "read" && argc >= 4) {
            auto file = slurp(argv[2]);
            const int steps = std::atoi(argv[3]);
            std::vector<uint8_t> buf(32 * 1024);
            size_t bytes = 0;
            uint32_t h = 2166136261u;
            double total = 0, t_new = 0, t_first = 0, t_close = 0;
            for (int s = -2; s < steps; s++) {  // two warm-up passes
                const auto t0 = clk::now();
                auto dec = alac::Decoder::New(file);  // NewDecoder is inside the timed region, like tests/benchmark_test.go:261-286
                const auto t1 = clk::now();
                size_t got = 0;
                uint32_t hh = 2166136261u;
                auto t2 = t1;
                for (;;) {
                    const size_t k = dec->Read(buf.data(), buf.size());
                    if (got == 0) t2 = clk::now();  // the first Read decodes the first window
                    if (k == 0) break;
                    got += k;
                    if (s == steps - 1) hh = fnv(hh, buf.data(), k);
                }
                const auto t3 = clk::now();
                dec.reset();
                if (s >= 0) {
                    total += std::chrono::duration<double>(clk::now() - t0).count();
                    t_new += std::chrono::duration<double>(t1 - t0).count();
                    t_first += std::chrono::duration<double>(t2 - t1).count();
                    t_close += std::chrono::duration<double>(clk::now() - t3).count();
                }
                bytes = got;
                h = hh;
            }
            // the same with ONE decoder kept open: Seek(0) + Read to EOF per step (what a player that re-reads pays)
            auto dec = alac::Decoder::New(file);
            double reuse = 0;
            for (int s = -1; s < steps; s++) {
                const auto t0 = clk::now();
                dec->Seek(0);
                for (;;)
                    if (dec->Read(buf.data(), buf.size()) == 0) break;
                if (s >= 0) reuse += std::chrono::duration<double>(clk::now() - t0).count();
            }
            std::printf("%.9f %zu %u %.9f %.9f %.9f %.9f\n", total / steps, bytes, h, reuse / steps, t_new / steps, t_first / steps, t_close / steps);
            return 0;
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "api_bench: %s\n", e.what());
        return 1;
    }
    return 2;
}
