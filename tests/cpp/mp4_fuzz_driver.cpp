// mp4_fuzz_driver.cpp -- test infrastructure: runs the container parser (saprobe-alac_b200/host/mp4.cpp, compiled
// into this binary with AddressSanitizer + UBSan) over a file of mutated M4A images written by the Python test.
// Input format: repeated { u32 little-endian length, bytes }.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/alac_b200.h"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> buf;
    unsigned ok = 0, bad = 0;
    for (;;) {
        uint8_t h[4];
        if (std::fread(h, 1, 4, f) != 4) break;
        const uint32_t n = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24);
        buf.resize(n);
        if (n && std::fread(buf.data(), 1, n, f) != n) return 3;
        // hand the parser an exact-size heap copy so that any read past the image is an ASan error
        uint8_t *img = new uint8_t[n ? n : 1];
        if (n) std::memcpy(img, buf.data(), n);
        alacb200_track *t = nullptr;
        const int32_t rc = alacb200_mp4_find_alac_track(img, n, &t);
        if (rc == ALACB200_OK) {
            ok++;
            size_t clen = 0;
            const uint8_t *c = alacb200_mp4_cookie(t, &clen);
            uint64_t ns = 0;
            const alacb200_sample_info *si = alacb200_mp4_samples(t, &ns);
            volatile uint64_t sink = 0;
            for (size_t i = 0; i < clen; i++) sink += c[i];
            for (uint64_t i = 0; i < ns; i++) sink += si[i].offset + si[i].size;
        } else {
            bad++;
            (void)alacb200_mp4_error(t);
        }
        alacb200_mp4_free_track(t);
        delete[] img;
    }
    std::fclose(f);
    std::printf("fuzz ok=%u bad=%u\n", ok, bad);
    return 0;
}
