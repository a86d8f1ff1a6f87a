// wire.cuh -- the reference's index wire formats (WriteTo / ReadFrom): FLAT, IVFX, PQIX, IVPQ, HNSW, all
// little-endian, magic + version 1, ending in the roaring blob of the deleted set (flat_index.go:366-614,
// ivf_index.go:468-785, pq_index.go:509-846, ivfpq_index.go:544-960, hnsw_index.go:734-1096).
//
// Byte sinks / sources (memory, plain file, gzip file -- the LSM layer stores `vector_%06d.bin.gz`,
// storage_provider.go:163-166) and the pieces every format shares: header, distance-kind string, roaring blob.
// Each index's field order lives next to its state (flat.cu, ivf.cu, pq.cu, hnsw.cu).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace cm {
namespace wire {

struct Sink {
    int64_t total = 0;
    virtual ~Sink() {}
    virtual bool put(const void *p, size_t n) = 0;
    bool u8(uint8_t v) { return put(&v, 1); }
    bool u32(uint32_t v) { return put(&v, 4); }          // the library only runs on little-endian hosts (x86-64 / arm64)
    bool i32(int32_t v) { return put(&v, 4); }
    bool f64(double v) { return put(&v, 8); }
};

struct Source {
    int64_t total = 0;
    virtual ~Source() {}
    virtual bool get(void *p, size_t n) = 0;
    bool u8(uint8_t *v) { return get(v, 1); }
    bool u32(uint32_t *v) { return get(v, 4); }
    bool i32(int32_t *v) { return get(v, 4); }
    bool f64(double *v) { return get(v, 8); }
};

// counts every byte; stores them while they fit (a NULL buffer turns it into a size query)
struct MemSink : Sink {
    uint8_t *buf;
    int64_t cap;
    MemSink(uint8_t *b, int64_t c) : buf(b), cap(c) {}
    bool put(const void *p, size_t n) override {
        if (buf && total + (int64_t)n <= cap) memcpy(buf + total, p, n);
        total += (int64_t)n;
        return true;
    }
};
struct MemSource : Source {
    const uint8_t *buf;
    int64_t len;
    MemSource(const uint8_t *b, int64_t l) : buf(b), len(l) {}
    bool get(void *p, size_t n) override {
        if (total + (int64_t)n > len) return false;
        memcpy(p, buf + total, n);
        total += (int64_t)n;
        return true;
    }
};

// wire.cu: file sinks / sources; a path ending in ".gz" is written gzip-compressed, reading detects gzip by itself
Sink *open_file_sink(const char *path);
int close_file_sink(Sink *s);               // CM_OK when every byte reached the file
Source *open_file_source(const char *path);
void close_file_source(Source *s);

const char *kind_name(int metric);           // distance.go:21-38: "l2" | "l2_squared" | "cosine"

// magic, version 1, dimensionality, distance kind (length + bytes): the first fields of every format
int write_header(Sink &s, const char *magic, int dim, int metric);
// the same fields back, validated against the pre-constructed index with the reference's error texts
int read_header(Source &s, const char *magic, int dim, int metric);

// The deleted set of a flushed index is empty and WriteTo always flushes first: the roaring blob is the 8-byte
// portable encoding of an empty bitmap (cookie 12346, zero containers), preceded by its length.
bool write_empty_bitmap(Sink &s);
// length + roaring portable format (RoaringFormatSpec: array, bitmap and run containers) -> the IDs it holds
int read_bitmap(Source &s, std::vector<uint32_t> *ids);
// decode one blob (exposed for tests)
int decode_roaring(const uint8_t *p, size_t n, std::vector<uint32_t> *ids);

#define CM_WIRE_GET(expr, what)                                                                   \
    do {                                                                                          \
        if (!(expr)) return ::cm::fail(CM_ERR_INVALID_ARG, "failed to read %s: unexpected EOF", what); \
    } while (0)
#define CM_WIRE_PUT(expr, what)                                                           \
    do {                                                                                  \
        if (!(expr)) return ::cm::fail(CM_ERR_INVALID_ARG, "failed to write %s", what);   \
    } while (0)

// Run `save(Sink &)` against a caller buffer (NULL = size query) / a file, `load(Source &)` against bytes / a file.
template <class F>
int save_to_buffer(F &&save, uint8_t *buf, int64_t cap, int64_t *bytes) {
    MemSink s(buf, cap);
    int rc = save(s);
    if (bytes) *bytes = s.total;
    if (rc != CM_OK) return rc;
    if (buf && s.total > cap) return fail(CM_ERR_BUFFER_TOO_SMALL, "serialised index needs %lld bytes, buffer holds %lld", (long long)s.total, (long long)cap);
    return CM_OK;
}
template <class F>
int save_to_file(F &&save, const char *path) {
    if (!path) return fail(CM_ERR_INVALID_ARG, "null path");
    Sink *s = open_file_sink(path);
    if (!s) return fail(CM_ERR_INVALID_ARG, "cannot open %s for writing", path);
    int rc = save(*s);
    int rc2 = close_file_sink(s);
    return rc != CM_OK ? rc : rc2;
}
template <class F>
int load_from_buffer(F &&load, const uint8_t *buf, int64_t len, int64_t *consumed) {
    if (!buf || len < 0) return fail(CM_ERR_INVALID_ARG, "null buffer");
    MemSource s(buf, len);
    int rc = load(s);
    if (consumed) *consumed = s.total;
    return rc;
}
template <class F>
int load_from_file(F &&load, const char *path) {
    if (!path) return fail(CM_ERR_INVALID_ARG, "null path");
    Source *s = open_file_source(path);
    if (!s) return fail(CM_ERR_INVALID_ARG, "cannot open %s for reading", path);
    int rc = load(*s);
    close_file_source(s);
    return rc;
}

}  // namespace wire
}  // namespace cm
