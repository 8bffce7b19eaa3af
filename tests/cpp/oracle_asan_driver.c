/* oracle_asan_driver.c -- test infrastructure: the CPU oracle (oracle/alac_oracle.c, compiled into this binary with
 * AddressSanitizer + UBSan) over a blob of packets written by the Python test. Every packet is handed over in an
 * exact-size heap block and decoded into an exact-size output block, so any access outside either is an ASan error.
 * Prints one "status out_bytes fnv1a(pcm)" line per packet so the test can compare with the regular oracle build.
 * Blob: repeated cases { u32 cookie_len, cookie, u32 npackets, { u32 len, bytes } x npackets }, little endian. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/alac_oracle.h"

static int rd32(FILE *f, uint32_t *v) {
    uint8_t h[4];
    if (fread(h, 1, 4, f) != 4) return 0;
    *v = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24);
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t clen;
    while (rd32(f, &clen)) {
        uint8_t *cookie = (uint8_t *)malloc(clen ? clen : 1);
        if (clen && fread(cookie, 1, clen, f) != clen) return 3;
        ao_config cfg;
        const int32_t cst = ao_parse_cookie(cookie, clen, &cfg);
        free(cookie);
        uint32_t np;
        if (!rd32(f, &np)) return 3;
        const int usable = cst == AO_OK && ao_check_config(&cfg) == AO_OK;
        const size_t cap = usable ? (size_t)cfg.frame_length * cfg.num_channels * (size_t)ao_bytes_per_sample(cfg.bit_depth) : 0;
        for (uint32_t i = 0; i < np; i++) {
            uint32_t n;
            if (!rd32(f, &n)) return 3;
            uint8_t *pk = (uint8_t *)malloc(n ? n : 1);
            if (n && fread(pk, 1, n, f) != n) return 3;
            if (usable) {
                uint8_t *out = (uint8_t *)malloc(cap ? cap : 1);
                uint32_t nb = 0;
                const int32_t st = ao_decode_packet(&cfg, pk, n, out, &nb);
                uint32_t h = 2166136261u;
                for (uint32_t k = 0; k < nb; k++) h = (h ^ out[k]) * 16777619u;
                printf("%d %u %u\n", st, nb, h);
                free(out);
            } else {
                printf("%d 0 0\n", cst);
            }
            free(pk);
        }
    }
    fclose(f);
    return 0;
}
