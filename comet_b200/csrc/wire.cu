// wire.cu -- shared pieces of the reference's index wire formats (see wire.cuh): file sinks / sources (plain and
// gzip), the common header, the roaring blob of the deleted set.  Host code only.
#include <zlib.h>

#include <cstdio>

#include "wire.cuh"

namespace cm {
namespace wire {

namespace {

bool ends_with_gz(const char *path) {
    size_t n = strlen(path);
    return n >= 3 && strcmp(path + n - 3, ".gz") == 0;
}

struct GzSink : Sink {
    gzFile f = nullptr;
    bool ok = true;
    bool put(const void *p, size_t n) override {
        const uint8_t *b = (const uint8_t *)p;
        while (n > 0 && ok) {
            unsigned m = (unsigned)(n > (1u << 30) ? (1u << 30) : n);
            if (gzwrite(f, b, m) != (int)m) ok = false;
            b += m; n -= m; total += m;
        }
        return ok;
    }
};
struct FileSink : Sink {
    FILE *f = nullptr;
    bool ok = true;
    bool put(const void *p, size_t n) override {
        if (ok && fwrite(p, 1, n, f) != n) ok = false;
        total += (int64_t)n;
        return ok;
    }
};
// gzread reads plain files transparently, so one source serves both
struct GzSource : Source {
    gzFile f = nullptr;
    bool get(void *p, size_t n) override {
        uint8_t *b = (uint8_t *)p;
        while (n > 0) {
            unsigned m = (unsigned)(n > (1u << 30) ? (1u << 30) : n);
            if (gzread(f, b, m) != (int)m) return false;
            b += m; n -= m; total += m;
        }
        return true;
    }
};

}  // namespace

Sink *open_file_sink(const char *path) {
    if (ends_with_gz(path)) {
        gzFile f = gzopen(path, "wb6");
        if (!f) return nullptr;
        gzbuffer(f, 1u << 20);
        GzSink *s = new GzSink();
        s->f = f;
        return s;
    }
    FILE *f = fopen(path, "wb");
    if (!f) return nullptr;
    FileSink *s = new FileSink();
    s->f = f;
    return s;
}
int close_file_sink(Sink *s) {
    bool ok = true;
    if (GzSink *g = dynamic_cast<GzSink *>(s)) {
        ok = g->ok;
        if (gzclose(g->f) != Z_OK) ok = false;
    } else if (FileSink *p = dynamic_cast<FileSink *>(s)) {
        ok = p->ok;
        if (fclose(p->f) != 0) ok = false;
    }
    delete s;
    return ok ? CM_OK : fail(CM_ERR_INVALID_ARG, "short write while serialising the index");
}
Source *open_file_source(const char *path) {
    gzFile f = gzopen(path, "rb");
    if (!f) return nullptr;
    gzbuffer(f, 1u << 20);
    GzSource *s = new GzSource();
    s->f = f;
    return s;
}
void close_file_source(Source *s) {
    if (GzSource *g = dynamic_cast<GzSource *>(s)) gzclose(g->f);
    delete s;
}

const char *kind_name(int metric) {
    return metric == CM_L2 ? "l2" : metric == CM_L2SQ ? "l2_squared" : "cosine";
}

int write_header(Sink &s, const char *magic, int dim, int metric) {
    const char *kind = kind_name(metric);
    CM_WIRE_PUT(s.put(magic, 4), "magic number");
    CM_WIRE_PUT(s.u32(1), "version");
    CM_WIRE_PUT(s.u32((uint32_t)dim), "dimensionality");
    CM_WIRE_PUT(s.u32((uint32_t)strlen(kind)), "distance kind length");
    CM_WIRE_PUT(s.put(kind, strlen(kind)), "distance kind");
    return CM_OK;
}

int read_header(Source &s, const char *magic, int dim, int metric) {
    char m[5] = {0, 0, 0, 0, 0};
    CM_WIRE_GET(s.get(m, 4), "magic number");
    if (memcmp(m, magic, 4) != 0) return fail(CM_ERR_INVALID_ARG, "invalid magic number: expected '%s', got '%s'", magic, m);
    uint32_t version = 0, d = 0, klen = 0;
    CM_WIRE_GET(s.u32(&version), "version");
    if (version != 1) return fail(CM_ERR_UNSUPPORTED, "unsupported version: %u", version);
    CM_WIRE_GET(s.u32(&d), "dimensionality");
    if ((int64_t)d != dim) return fail(CM_ERR_DIM_MISMATCH, "dimension mismatch: index has dim=%d, serialized data has dim=%u", dim, d);
    CM_WIRE_GET(s.u32(&klen), "distance kind length");
    if (klen > 64) return fail(CM_ERR_INVALID_ARG, "distance kind mismatch: index uses '%s', serialized data uses a %u-byte name", kind_name(metric), klen);
    std::string kind(klen, '\0');
    CM_WIRE_GET(s.get(&kind[0], klen), "distance kind");
    if (kind != kind_name(metric))
        return fail(CM_ERR_INVALID_ARG, "distance kind mismatch: index uses '%s', serialized data uses '%s'", kind_name(metric), kind.c_str());
    return CM_OK;
}

bool write_empty_bitmap(Sink &s) {
    static const uint8_t empty[8] = {0x3A, 0x30, 0, 0, 0, 0, 0, 0};   // SERIAL_COOKIE_NO_RUNCONTAINER (12346), 0 containers
    return s.u32(8) && s.put(empty, 8);
}

// RoaringFormatSpec (the portable format roaring v1.9.4's ToBytes / UnmarshalBinary use): cookie; container count;
// [run flags]; per container (key u16, cardinality-1 u16); [offsets]; containers (array: u16 values; bitmap: 1024 x
// u64; run: count u16 then (start u16, length-1 u16) pairs).
int decode_roaring(const uint8_t *p, size_t n, std::vector<uint32_t> *ids) {
    ids->clear();
    if (n == 0) return CM_OK;
    size_t at = 0;
    auto need = [&](size_t k) { return at + k <= n; };
    auto rd16 = [&](uint16_t *v) { memcpy(v, p + at, 2); at += 2; };
    auto rd32 = [&](uint32_t *v) { memcpy(v, p + at, 4); at += 4; };
    const char *bad = "failed to deserialize deleted nodes bitmap: truncated roaring data";
    if (!need(4)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
    uint32_t cookie = 0, size = 0;
    rd32(&cookie);
    std::vector<uint8_t> is_run;
    bool has_run = false;
    if ((cookie & 0xFFFF) == 12347) {
        has_run = true;
        size = (cookie >> 16) + 1;
        size_t nb = (size + 7) / 8;
        if (!need(nb)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
        is_run.assign(p + at, p + at + nb);
        at += nb;
    } else if (cookie == 12346) {
        if (!need(4)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
        rd32(&size);
    } else {
        return fail(CM_ERR_INVALID_ARG, "failed to deserialize deleted nodes bitmap: unknown roaring cookie %u", cookie);
    }
    if (size > 65536) return fail(CM_ERR_INVALID_ARG, "failed to deserialize deleted nodes bitmap: %u containers", size);
    std::vector<uint16_t> keys(size), cards(size);
    if (!need((size_t)size * 4)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
    for (uint32_t i = 0; i < size; i++) { rd16(&keys[i]); rd16(&cards[i]); }
    if (!has_run || size >= 4) {                  // offset header (skipped: containers follow in order)
        if (!need((size_t)size * 4)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
        at += (size_t)size * 4;
    }
    for (uint32_t i = 0; i < size; i++) {
        const uint32_t hi = (uint32_t)keys[i] << 16;
        const uint32_t card = (uint32_t)cards[i] + 1;
        if (has_run && (is_run[i / 8] >> (i % 8) & 1)) {
            uint16_t nruns = 0;
            if (!need(2)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
            rd16(&nruns);
            if (!need((size_t)nruns * 4)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
            for (uint16_t r = 0; r < nruns; r++) {
                uint16_t start = 0, len1 = 0;
                rd16(&start); rd16(&len1);
                for (uint32_t v = start; v <= (uint32_t)start + len1; v++) ids->push_back(hi | v);
            }
        } else if (card > 4096) {
            if (!need(8192)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
            for (uint32_t w = 0; w < 1024; w++) {
                uint64_t bits = 0;
                memcpy(&bits, p + at + (size_t)w * 8, 8);
                while (bits) {
                    int b = __builtin_ctzll(bits);
                    ids->push_back(hi | (w * 64 + (uint32_t)b));
                    bits &= bits - 1;
                }
            }
            at += 8192;
        } else {
            if (!need((size_t)card * 2)) return fail(CM_ERR_INVALID_ARG, "%s", bad);
            for (uint32_t j = 0; j < card; j++) {
                uint16_t v = 0;
                rd16(&v);
                ids->push_back(hi | v);
            }
        }
    }
    return CM_OK;
}

int read_bitmap(Source &s, std::vector<uint32_t> *ids) {
    uint32_t size = 0;
    CM_WIRE_GET(s.u32(&size), "bitmap size");
    std::vector<uint8_t> blob(size);
    CM_WIRE_GET(size == 0 || s.get(blob.data(), size), "bitmap data");
    return decode_roaring(blob.data(), blob.size(), ids);
}

}  // namespace wire
}  // namespace cm

// exposed for the tests of the roaring decoder (a hand-built blob per container type)
extern "C" int cm_debug_decode_roaring(const uint8_t *blob, int64_t len, uint32_t *out_ids, int64_t cap, int64_t *count) {
    std::vector<uint32_t> ids;
    int rc = cm::wire::decode_roaring(blob, (size_t)(len > 0 ? len : 0), &ids);
    if (rc != CM_OK) return rc;
    if (count) *count = (int64_t)ids.size();
    if (out_ids)
        for (size_t i = 0; i < ids.size() && (int64_t)i < cap; i++) out_ids[i] = ids[i];
    return CM_OK;
}
