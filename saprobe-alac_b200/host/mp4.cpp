// mp4.cpp -- host-side container reader: the sample-table packer in front of the GPU decoder.
//
// Mirrors internal/mp4 of the reference (mp4.go:60-591) over an in-memory file image: box walker,
// FindALACTrack (moov/trak/mdia/minf/stbl), cookie from stsd (QuickTime v1 entries included), sample
// table = stco|co64 x stsc x stsz (constant-size stsz included). Behaviour kept on purpose: stts is
// never read, the first trak with an 'alac' sample entry wins, a trak whose stsd is unusable is skipped.
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/alac_b200.h"

namespace {

struct Reader {  // io.ReadSeeker over a byte image
    const uint8_t *data;
    int64_t len;
    int64_t pos;
    bool has(int64_t n) const { return pos >= 0 && pos <= len && n >= 0 && n <= len - pos; }
    bool read_full(uint8_t *dst, int64_t n) {  // io.ReadFull
        if (pos < 0 || pos > len || n > len - pos) {
            pos = len;
            return false;
        }
        if (n > 0) std::memcpy(dst, data + pos, (size_t)n);  // dst may be null for an empty table
        pos += n;
        return true;
    }
};

struct Box {  // boxInfo, mp4.go:38-46
    int64_t offset = 0, size = 0, header = 0;
    char fourcc[4] = {0, 0, 0, 0};
    int64_t payload_offset() const { return offset + header; }
    int64_t payload_size() const { return size - header; }
    bool is(const char *t) const { return std::memcmp(fourcc, t, 4) == 0; }
};

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }

enum class Rd { Ok, Eof, Bad };

// readBoxInfo, mp4.go:60-113
Rd read_box(Reader &r, Box &b, std::string &err) {
    b = Box();
    b.offset = r.pos;
    uint8_t h[16];
    if (!r.read_full(h, 8)) return Rd::Eof;
    b.header = 8;
    std::memcpy(b.fourcc, h + 4, 4);
    uint32_t raw = be32(h);
    if (raw == 0) {
        b.size = r.len - b.offset;  // box extends to end of file
        r.pos = b.offset + b.header;
    } else if (raw == 1) {
        if (!r.read_full(h + 8, 8)) return Rd::Eof;
        b.header = 16;
        b.size = (int64_t)be64(h + 8);
    } else {
        b.size = raw;
    }
    if (b.size < b.header) {
        err = "mp4: invalid box size";
        return Rd::Bad;
    }
    // a 64-bit size near INT64_MAX must not overflow offset + size below: a box cannot end behind the image anyway
    // (the reference only notices when it reads, mp4.go:98-139)
    if (b.size > INT64_MAX - b.offset) b.size = INT64_MAX - b.offset;
    return Rd::Ok;
}

// iterChildren + findChild, mp4.go:141-205. Returns false on a structural error.
template <class F>
bool iter_children(Reader &r, const Box &parent, std::string &err, F &&cb) {
    r.pos = parent.payload_offset();
    const int64_t end = parent.offset + parent.size;
    for (;;) {
        if (r.pos >= end) return true;
        Box child;
        Rd rc = read_box(r, child, err);
        if (rc == Rd::Eof) return true;  // a truncated tail ends the walk quietly (mp4.go:165-167)
        if (rc == Rd::Bad) return false;
        bool stop = false;
        if (!cb(child, stop)) return false;
        if (stop) return true;
        r.pos = child.offset + child.size;
    }
}
bool find_child(Reader &r, const Box &parent, const char *target, Box &out, bool &found, std::string &err) {
    found = false;
    return iter_children(r, parent, err, [&](const Box &c, bool &stop) {
        if (c.is(target)) {
            out = c;
            found = true;
            stop = true;
        }
        return true;
    });
}

// extractCookie, mp4.go:313-378. Returns false when this is not an ALAC track.
bool extract_cookie(Reader &r, const Box &stbl, std::vector<uint8_t> &cookie) {
    Box stsd;
    bool found;
    std::string err;
    if (!find_child(r, stbl, "stsd", stsd, found, err) || !found) return false;
    int64_t plen = stsd.payload_size();
    if (plen < 0 || plen > r.len) return false;
    std::vector<uint8_t> data((size_t)plen);
    r.pos = stsd.payload_offset();
    if (!r.read_full(data.data(), plen)) return false;
    if (plen < 8) return false;
    uint32_t entries = be32(&data[4]);
    int64_t pos = 8;
    for (uint32_t e = 0; e < entries; e++) {
        if (pos + 8 > plen) break;
        int64_t esize = be32(&data[(size_t)pos]);
        if (esize < 8 + 28 || pos + esize > plen) {
            pos += esize;
            continue;
        }
        if (std::memcmp(&data[(size_t)pos + 4], "alac", 4) != 0) {
            pos += esize;
            continue;
        }
        uint32_t version = ((uint32_t)data[(size_t)pos + 16] << 8) | data[(size_t)pos + 17];
        int64_t skip = 8 + 28 + (version == 1 ? 16 : 0);  // QuickTime v1 extra fields, mp4.go:360-365
        int64_t cs = pos + skip, ce = pos + esize;
        if (cs >= ce) return false;  // ErrInvalidEntry
        cookie.assign(data.begin() + cs, data.begin() + ce);
        return true;
    }
    return false;
}

bool read_table(Reader &r, const Box &box, int64_t header_bytes, std::vector<uint8_t> &hdr) {
    r.pos = box.payload_offset();
    hdr.resize((size_t)header_bytes);
    return r.read_full(hdr.data(), header_bytes);
}

// buildSampleTable, mp4.go:382-420
bool build_sample_table(Reader &r, const Box &stbl, std::vector<alacb200_sample_info> &out, std::string &err) {
    std::vector<uint64_t> chunk_offsets;
    Box box;
    bool found;
    std::string e2;
    std::vector<uint8_t> hdr, buf;
    // readChunkOffsets, mp4.go:422-436
    if (find_child(r, stbl, "stco", box, found, e2) && found) {
        if (!read_table(r, box, 8, hdr)) { err = "mp4: no chunk offset box (stco/co64)"; return false; }
        uint32_t count = be32(&hdr[4]);
        if (!r.has((int64_t)count * 4)) { err = "mp4: no chunk offset box (stco/co64)"; return false; }  // before any allocation
        buf.resize((size_t)count * 4);
        if (!r.read_full(buf.data(), (int64_t)count * 4)) { err = "mp4: no chunk offset box (stco/co64)"; return false; }
        chunk_offsets.resize(count);
        for (uint32_t i = 0; i < count; i++) chunk_offsets[i] = be32(&buf[(size_t)i * 4]);
    } else {
        if (!find_child(r, stbl, "co64", box, found, e2) || !found) { err = "mp4: no chunk offset box (stco/co64)"; return false; }
        if (!read_table(r, box, 8, hdr)) { err = "mp4: invalid co64 payload"; return false; }
        uint32_t count = be32(&hdr[4]);
        if (!r.has((int64_t)count * 8)) { err = "mp4: invalid co64 payload"; return false; }
        buf.resize((size_t)count * 8);
        if (!r.read_full(buf.data(), (int64_t)count * 8)) { err = "mp4: invalid co64 payload"; return false; }
        chunk_offsets.resize(count);
        for (uint32_t i = 0; i < count; i++) chunk_offsets[i] = be64(&buf[(size_t)i * 8]);
    }
    // readStsc, mp4.go:488-523
    if (!find_child(r, stbl, "stsc", box, found, e2) || !found) { err = "mp4: no stsc box"; return false; }
    if (!read_table(r, box, 8, hdr)) { err = "mp4: invalid stsc payload"; return false; }
    uint32_t nstsc = be32(&hdr[4]);
    if (!r.has((int64_t)nstsc * 12)) { err = "mp4: invalid stsc payload"; return false; }
    buf.resize((size_t)nstsc * 12);
    if (!r.read_full(buf.data(), (int64_t)nstsc * 12)) { err = "mp4: invalid stsc payload"; return false; }
    std::vector<std::pair<uint32_t, uint32_t>> stsc(nstsc);
    for (uint32_t i = 0; i < nstsc; i++) stsc[i] = {be32(&buf[(size_t)i * 12]), be32(&buf[(size_t)i * 12 + 4])};
    // readStsz, mp4.go:525-565
    if (!find_child(r, stbl, "stsz", box, found, e2) || !found) { err = "mp4: no stsz box"; return false; }
    if (!read_table(r, box, 12, hdr)) { err = "mp4: invalid stsz payload"; return false; }
    uint32_t constant = be32(&hdr[4]), count = be32(&hdr[8]);
    std::vector<uint32_t> sizes;
    if (constant == 0) {
        if (!r.has((int64_t)count * 4)) { err = "mp4: invalid stsz payload"; return false; }
        buf.resize((size_t)count * 4);
        if (!r.read_full(buf.data(), (int64_t)count * 4)) { err = "mp4: invalid stsz payload"; return false; }
        sizes.resize(count);
        for (uint32_t i = 0; i < count; i++) sizes[i] = be32(&buf[(size_t)i * 4]);
    }
    out.clear();
    // A table cannot name more samples than the image has bytes (a constant-size stsz carries no per-sample data, so its
    // count is otherwise unbounded; the reference would try to allocate it).
    if ((int64_t)count > r.len) { err = "mp4: invalid stsz payload"; return false; }
    out.reserve(count);
    uint64_t sample_idx = 0;
    size_t k = 0;  // first stsc entry whose first_chunk exceeds the current chunk number: it only ever moves forward
    for (size_t ci = 0; ci < chunk_offsets.size() && sample_idx < count; ci++) {
        // lookupSamplesPerChunk, mp4.go:580-591 (1-based chunk numbers): the last entry before the first one that
        // starts at a later chunk. Same result as the reference's scan from the top, without its quadratic cost.
        while (k < stsc.size() && stsc[k].first <= (uint32_t)(ci + 1)) k++;
        const uint32_t per = k ? stsc[k - 1].second : 0;
        uint64_t off = chunk_offsets[ci];
        for (uint32_t it = 0; it < per && sample_idx < count; it++) {
            uint32_t sz = constant ? constant : sizes[(size_t)sample_idx];
            out.push_back(alacb200_sample_info{off, sz, 0});
            off += sz;
            sample_idx++;
        }
    }
    return true;
}

}  // namespace

struct alacb200_track {
    std::vector<uint8_t> cookie;
    std::vector<alacb200_sample_info> samples;
    std::string error;
};

extern "C" {

// FindALACTrack, mp4.go:233-300
static int32_t find_alac_track(const uint8_t *file, uint64_t file_len, alacb200_track *t);

int32_t alacb200_mp4_find_alac_track(const uint8_t *file, uint64_t file_len, alacb200_track **out) {
    if (!out) return ALACB200_E_ARG;
    *out = nullptr;
    alacb200_track *t = nullptr;
    try {  // nothing may unwind through the C ABI
        t = new alacb200_track();
        *out = t;
        return find_alac_track(file, file_len, t);
    } catch (const std::exception &e) {
        if (t) {
            t->samples.clear();
            t->cookie.clear();
            try { t->error = std::string("mp4: ") + e.what(); } catch (...) {}
        }
        return ALACB200_E_NO_TRACK;
    }
}

static int32_t find_alac_track(const uint8_t *file, uint64_t file_len, alacb200_track *t) {
    if (!file && file_len) {
        t->error = "null file image";
        return ALACB200_E_ARG;
    }
    Reader r{file, (int64_t)file_len, 0};
    Box root;
    root.offset = 0;
    root.size = (int64_t)file_len;
    root.header = 0;
    Box moov;
    bool found = false;
    if (!find_child(r, root, "moov", moov, found, t->error)) {
        t->error = "reading container structure: " + t->error;
        return ALACB200_E_NO_TRACK;
    }
    if (!found) {
        t->error = "mp4: no ALAC track found in container";
        return ALACB200_E_NO_TRACK;
    }
    bool have = false, failed = false;
    bool ok = iter_children(r, moov, t->error, [&](const Box &trak, bool &stop) {
        if (!trak.is("trak")) return true;
        Box cur = trak, next;
        for (const char *name : {"mdia", "minf", "stbl"}) {
            bool f = false;
            if (!find_child(r, cur, name, next, f, t->error)) return false;
            if (!f) return true;  // not this trak
            cur = next;
        }
        std::vector<uint8_t> cookie;
        if (!extract_cookie(r, cur, cookie)) return true;  // "not an ALAC track": continue, mp4.go:279-281
        std::string terr;
        if (!build_sample_table(r, cur, t->samples, terr)) {
            t->error = "building sample table: " + terr;
            failed = true;
            return false;
        }
        t->cookie = std::move(cookie);
        have = true;
        stop = true;
        return true;
    });
    if (!ok || failed) return ALACB200_E_NO_TRACK;
    if (!have) {
        t->error = "mp4: no ALAC track found in container";
        return ALACB200_E_NO_TRACK;
    }
    return ALACB200_OK;
}

void alacb200_mp4_free_track(alacb200_track *t) { delete t; }

const uint8_t *alacb200_mp4_cookie(const alacb200_track *t, size_t *len) {
    if (len) *len = t ? t->cookie.size() : 0;
    return t && !t->cookie.empty() ? t->cookie.data() : nullptr;
}
const alacb200_sample_info *alacb200_mp4_samples(const alacb200_track *t, uint64_t *count) {
    if (count) *count = t ? t->samples.size() : 0;
    return t && !t->samples.empty() ? t->samples.data() : nullptr;
}
const char *alacb200_mp4_error(const alacb200_track *t) { return t ? t->error.c_str() : ""; }

}  // extern "C"
