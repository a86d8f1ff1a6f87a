// flat.cu -- FlatIndex device state and the cm_flat_* / distance entry points of the C ABI.
//
// Device layout (replaces []VectorNode with one heap slice per node, flat_index.go:82, node.go:30-33):
//   rows     fp32 [cap][ld]   row-major, ld = dim rounded up to 32 floats (128 B) and zero padded, so
//                             every TMA box is full and padding adds exact zeros to the sums;
//   ids      u32  [cap]       node IDs by scan position;
//   deleted  u8   [cap]       1 when the row's ID is in the soft-delete set (flat_index.go:87);
//   rows_bf16 / norms         shadow copy for the tensor-core candidate pass (flat_tensor.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "wire.cuh"

namespace cm {

// ---- small stream pool so concurrent searches (RLock holders) do not serialise ---------------
static std::mutex g_stream_mu;
static std::map<int, std::vector<cudaStream_t>> g_stream_pool;
static std::map<cudaStream_t, int> g_stream_home;     // per device: a stream belongs to the device it was made on

int acquire_stream(cudaStream_t *out) {
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        std::vector<cudaStream_t> &pool = g_stream_pool[dev];
        if (!pool.empty()) {
            *out = pool.back();
            pool.pop_back();
            return CM_OK;
        }
    }
    CM_CUDA(cudaStreamCreateWithFlags(out, cudaStreamNonBlocking));
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_stream_home[*out] = dev;
    return CM_OK;
}
void release_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_stream_pool[g_stream_home[s]].push_back(s);
}

static std::set<int> g_pool_tuned;
static void tune_mempool() {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        if (!g_pool_tuned.insert(dev).second) return;
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thresh = ~0ull;   // keep freed workspace cached: searches reuse it
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
}
int ws_alloc(void **p, size_t bytes, cudaStream_t s) {
    static thread_local int tuned_dev = -1;   // the lock above is taken once per (thread, device change), not per call
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != tuned_dev) { tune_mempool(); tuned_dev = dev; }
    CM_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, s));
    return CM_OK;
}
void ws_free(void *p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// per-thread pinned staging for the zero-vector flags of the query batch in flight
static thread_local int *tl_zero_flags = nullptr;
static thread_local int64_t tl_zero_cap = 0, tl_zero_n = 0;
int *zero_flags_host(int64_t nq) {
    if (nq > tl_zero_cap) {
        if (tl_zero_flags) cudaFreeHost(tl_zero_flags);
        tl_zero_cap = std::max<int64_t>(nq, 4096);
        if (cudaHostAlloc((void **)&tl_zero_flags, (size_t)tl_zero_cap * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
            tl_zero_flags = nullptr; tl_zero_cap = 0;
        }
    }
    tl_zero_n = nq;
    return tl_zero_flags;
}
static int64_t first_zero_query() {
    for (int64_t i = 0; i < tl_zero_n; i++)
        if (tl_zero_flags[i]) return i;
    return -1;
}

int FlatIndex::rebuild_tmap() {
    if (!rows) return CM_OK;
    return make_tmap_2d(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, (uint64_t)ld, (uint64_t)cap,
                        (uint64_t)ld * 4, SCAN_CHUNK, SCAN_TILE_ROWS, CU_TENSOR_MAP_SWIZZLE_128B);
}

int FlatIndex::reserve(int64_t want) {
    if (want <= cap) return CM_OK;
    int64_t ncap = cap ? cap : 1024;
    while (ncap < want) ncap = ncap + ncap / 2 + 1024;
    ncap = (ncap + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS * SCAN_TILE_ROWS;
    float *nrows = nullptr;
    uint32_t *nids = nullptr;
    uint8_t *ndel = nullptr;
    CM_CUDA(cudaMalloc(&nrows, (size_t)ncap * ld * sizeof(float)));
    CM_CUDA(cudaMalloc(&nids, (size_t)ncap * sizeof(uint32_t)));
    CM_CUDA(cudaMalloc(&ndel, (size_t)ncap));
    CM_CUDA(cudaMemset(ndel, 0, (size_t)ncap));
    if (n > 0) {
        CM_CUDA(cudaMemcpy(nrows, rows, (size_t)n * ld * sizeof(float), cudaMemcpyDeviceToDevice));
        CM_CUDA(cudaMemcpy(nids, ids, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        CM_CUDA(cudaMemcpy(ndel, deleted, (size_t)n, cudaMemcpyDeviceToDevice));
    }
    // the memset / copies above ran on the legacy default stream, which the library's non-blocking streams do not
    // synchronise with: make them complete before anything else can touch the new buffers
    CM_CUDA(cudaDeviceSynchronize());
    cudaFree(rows); cudaFree(ids); cudaFree(deleted);
    rows = nrows; ids = nids; deleted = ndel; cap = ncap;
    shadow_rows = 0;   // tensor-path shadow must be rebuilt against the new buffers
    return rebuild_tmap();
}

FlatIndex::~FlatIndex() {
    cudaFree(rows); cudaFree(ids); cudaFree(deleted);
    cudaFree(staged_dev);
    cudaFree(id_sorted); cudaFree(pos_sorted);
    cudaFreeHost(staged_host);
    free_shadow();
}

// n successive Add()s.  src is a device buffer [n][dim] of RAW vectors; they are preprocessed into
// the index (reference order) and, when writeback_host != nullptr, copied back preprocessed.
int FlatIndex::add_from_device(const uint32_t *ids_host, const float *src_dev, int64_t n_add, float *writeback_host,
                               cudaStream_t st) {
    WsScope ws(st);
    if (n_add <= 0) return CM_OK;
    CM_TRY(reserve(n + n_add));
    bool fma = rounding_mode() == CM_ROUND_FMA;
    int *flags = nullptr;
    CM_TRY(ws.get(&flags, (size_t)n_add * sizeof(int)));
    const int pre_metric = raw_rows ? CM_L2 : metric;
    CM_TRY(launch_preprocess_rows(pre_metric, fma, src_dev, n_add, dim, dim, rows + (size_t)n * ld, ld, flags, st));
    int64_t good = n_add;
    if (pre_metric == CM_COSINE) {
        std::vector<int> hflags((size_t)n_add);
        CM_CUDA(cudaMemcpyAsync(hflags.data(), flags, (size_t)n_add * sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < n_add; i++)
            if (hflags[(size_t)i]) { good = i; break; }
    }
    if (good > 0) {
        std::vector<uint8_t> del((size_t)good, 0);
        bool any_del = false;
        for (int64_t i = 0; i < good; i++) {
            if (!deleted_ids.empty() && deleted_ids.count(ids_host[i])) { del[(size_t)i] = 1; any_del = true; }
        }
        CM_CUDA(cudaMemcpyAsync(ids + n, ids_host, (size_t)good * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (any_del) CM_CUDA(cudaMemcpyAsync(deleted + n, del.data(), (size_t)good, cudaMemcpyHostToDevice, st));
        else CM_CUDA(cudaMemsetAsync(deleted + n, 0, (size_t)good, st));
        if (writeback_host && pre_metric == CM_COSINE)
            CM_CUDA(cudaMemcpy2DAsync(writeback_host, (size_t)dim * 4, rows + (size_t)n * ld, (size_t)ld * 4,
                                      (size_t)dim * 4, (size_t)good, cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        ids_host_mirror.insert(ids_host_mirror.end(), ids_host, ids_host + good);
        for (int64_t i = 0; i < good; i++) n_deleted_rows += del[(size_t)i];
        n += good;
    }
    if (good < n_add)
        return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (row %lld of this Add batch)", (long long)good);
    return CM_OK;
}

int FlatIndex::remove(uint32_t id) {
    // flat_index.go:219-250: error when the ID is absent or already deleted
    std::vector<int64_t> hits;
    for (int64_t i = 0; i < n; i++)
        if (ids_host_mirror[(size_t)i] == id) hits.push_back(i);
    if (hits.empty()) return fail(CM_ERR_NOT_FOUND, "vector with ID %u not found", id);
    if (deleted_ids.count(id)) return fail(CM_ERR_NOT_FOUND, "vector with ID %u already deleted", id);
    deleted_ids.insert(id);
    uint8_t one = 1;
    for (int64_t i : hits) CM_CUDA(cudaMemcpy(deleted + i, &one, 1, cudaMemcpyHostToDevice));
    n_deleted_rows += (int64_t)hits.size();
    return CM_OK;
}

int FlatIndex::flush() {
    // flat_index.go:266-299: drop soft-deleted rows, keep scan order, clear the set
    if (deleted_ids.empty()) return CM_OK;
    std::vector<int64_t> keep;
    keep.reserve((size_t)n);
    for (int64_t i = 0; i < n; i++)
        if (!deleted_ids.count(ids_host_mirror[(size_t)i])) keep.push_back(i);
    int64_t m = (int64_t)keep.size();
    if (m > 0) {
        cudaStream_t st;
        CM_TRY(acquire_stream(&st));
        float *tmp = nullptr;
        int64_t *dpos = nullptr;
        int rc = ws_alloc((void **)&tmp, (size_t)m * ld * 4, st);
        if (rc == CM_OK) rc = ws_alloc((void **)&dpos, (size_t)m * 8, st);
        if (rc == CM_OK) {
            cudaMemcpyAsync(dpos, keep.data(), (size_t)m * 8, cudaMemcpyHostToDevice, st);
            rc = launch_gather_rows(rows, ld, ld, dpos, m, tmp, st);
            cudaMemcpyAsync(rows, tmp, (size_t)m * ld * 4, cudaMemcpyDeviceToDevice, st);
        }
        std::vector<uint32_t> nids((size_t)m);
        for (int64_t i = 0; i < m; i++) nids[(size_t)i] = ids_host_mirror[(size_t)keep[(size_t)i]];
        cudaMemcpyAsync(ids, nids.data(), (size_t)m * 4, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(deleted, 0, (size_t)cap, st);
        ws_free(tmp, st); ws_free(dpos, st);
        cudaError_t e = cudaStreamSynchronize(st);
        release_stream(st);
        if (rc != CM_OK) return rc;
        if (e != cudaSuccess) return fail(CM_ERR_CUDA, "flush: %s", cudaGetErrorString(e));
        ids_host_mirror.swap(nids);
    } else {
        ids_host_mirror.clear();
        CM_CUDA(cudaMemset(deleted, 0, (size_t)cap));
    }
    n = m;
    n_deleted_rows = 0;
    deleted_ids.clear();
    shadow_rows = 0;
    sorted_n = -1;
    return CM_OK;
}

// nq independent searchSingleQuery calls; everything on `st`, outputs to device pointers.
int FlatIndex::search_device(const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                             uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts,
                             cudaStream_t st, bool check_zero_queries) {
    WsScope ws(st);
    cm_flat_stats stats{};
    int64_t launches0 = g_kernel_launches.load();
    if (nq <= 0) return CM_OK;
    int64_t k_eff = p->k;
    if (k_eff <= 0 || k_eff > n) k_eff = n;                      // limiter.go:12-17 on len(idx.vectors)
    if (out_stride < k_eff)
        return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)k_eff);
    bool fma = rounding_mode() == CM_ROUND_FMA;
    if (n == 0 || k_eff == 0) {
        CM_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)nq * sizeof(int64_t), st));
        if (metric == CM_COSINE) {
            // the reference preprocesses the query before it looks at the index (flat_index_search.go:236): a zero
            // query fails with ErrZeroVector on an empty index too
            float *qp0 = nullptr;
            int *qf0 = nullptr;
            CM_TRY(ws.get(&qp0, (size_t)nq * ld * 4));
            CM_TRY(ws.get(&qf0, (size_t)nq * sizeof(int)));
            CM_TRY(launch_preprocess_rows(metric, fma, q_dev, nq, dim, dim, qp0, ld, qf0, st));
            if (check_zero_queries) CM_CUDA(cudaMemcpyAsync(zero_flags_host(nq), qf0, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
            else CM_TRY(launch_mark_zero_queries(qf0, nq, out_counts, st));
        }
        return CM_OK;
    }

    // 1. soft deletes + document filter (flat_index_search.go:255-263): the caller's IDs are sorted on the device
    //    (flat_filter.cu); a selective filter becomes an explicit candidate list, any other a per-row skip mask
    const uint8_t *skip = nullptr;
    uint8_t *skip_buf = nullptr;
    uint32_t *filt_dev = nullptr;
    const bool has_filter = p->filter_ids && p->nfilter > 0;
    bool gather = false;
    if (has_filter) {
        CM_TRY(ws.get(&filt_dev, (size_t)p->nfilter * 4));
        CM_TRY(upload_filter_ids(p->filter_ids, p->nfilter, filt_dev, st));
        CM_TRY(sort_u32_device(filt_dev, p->nfilter, ws, st));
        gather = p->nfilter <= n / 16 && k_eff <= 4096;
        if (const char *e = getenv("COMET_B200_FILTER_GATHER")) gather = atoi(e) != 0 && k_eff <= 4096;
        if (!gather) {
            CM_TRY(ws.get(&skip_buf, (size_t)n));
            CM_TRY(launch_build_skip(ids, deleted, n, filt_dev, p->nfilter, skip_buf, st));
            skip = skip_buf;
        }
    } else if (n_deleted_rows > 0) {
        skip = deleted;
    }

    // 2. pick the pipeline
    int path = p->path;
    if (path == CM_PATH_AUTO) {
        // a filter (or a deleted set) that leaves under half of the rows may leave the candidate pass short of k survivors
        // in its first samples: it would detect that and the host would redo the query exactly -- right, but slower
        const bool sparse = (has_filter && p->nfilter < n / 2) || n_deleted_rows > n / 2;
        path = tensor_path_eligible(nq, k_eff, sparse, p->threshold) ? CM_PATH_TENSOR : CM_PATH_EXACT;
    }

    // 3. Distance.Preprocess on every query (flat_index_search.go:236) into a zero-padded [nq_pad][ld] block (the
    //    tensor path folds it into its own query-preparation kernel), then the search proper
    int64_t nq_pad = (nq + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
    float *qp = nullptr;
    int *qflags = nullptr;
    CM_TRY(ws.get(&qp, (size_t)nq_pad * ld * 4));
    CM_TRY(ws.get(&qflags, (size_t)nq * sizeof(int)));
    int rc = CM_OK;
    if (gather) {
        // selective filter: look the IDs up, score exactly those rows (no pass over the corpus)
        if (nq_pad > nq) CM_CUDA(cudaMemsetAsync(qp + (size_t)nq * ld, 0, (size_t)(nq_pad - nq) * ld * 4, st));
        CM_TRY(launch_preprocess_rows(metric, fma, q_dev, nq, dim, dim, qp, ld, qflags, st));
        CM_TRY(ensure_id_sort(st));
        const int64_t cap = std::max<int64_t>(128, (int64_t)p->nfilter * max_id_run);
        uint32_t *cand_pos = nullptr;
        int *cand_cnt = nullptr;
        CM_TRY(ws.get(&cand_pos, (size_t)cap * 4));
        CM_TRY(ws.get(&cand_cnt, sizeof(int)));
        CM_TRY(filter_candidates(filt_dev, p->nfilter, cand_pos, cap, cand_cnt, ws, st));
        rc = gather_scan_topk(*this, qp, nq, cand_pos, cand_cnt, cap, p->threshold, k_eff, out_stride, out_ids, out_scores, out_pos,
                              out_counts, st);
        stats.path_used = CM_PATH_EXACT;
        stats.passes = 0;
    } else if (path == CM_PATH_TENSOR) {
        rc = search_tensor(q_dev, qp, qflags, nq, k_eff, skip, p->threshold, out_stride, out_ids, out_scores, out_pos, out_counts, st, &stats);
    } else {
        if (nq_pad > nq) CM_CUDA(cudaMemsetAsync(qp + (size_t)nq * ld, 0, (size_t)(nq_pad - nq) * ld * 4, st));
        CM_TRY(launch_preprocess_rows(metric, fma, q_dev, nq, dim, dim, qp, ld, qflags, st));
        rc = search_exact(qp, nq, nq_pad, k_eff, skip, p->threshold, out_stride, out_ids, out_scores, out_pos, out_counts, st, &stats);
    }
    if (rc == CM_OK && check_zero_queries && metric == CM_COSINE) {
        // the host entry point reads these flags after its final synchronise: a zero query is reported
        // then (its row of results is garbage by then, and discarded) without stalling the pipeline here
        CM_CUDA(cudaMemcpyAsync(zero_flags_host(nq), qflags, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
    } else if (rc == CM_OK && metric == CM_COSINE) {
        // device entry point: nobody reads the flags on the host -- a zero query is reported as count -2
        CM_TRY(launch_mark_zero_queries(qflags, nq, out_counts, st));
    }
    stats.kernel_launches = g_kernel_launches.load() - launches0;
    {
        std::lock_guard<std::mutex> lk(stats_mu);
        last_stats = stats;
    }
    return rc;
}

int FlatIndex::search_exact(const float *qp, int64_t nq, int64_t nq_pad, int64_t k_eff, const uint8_t *skip,
                            float threshold, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                            int64_t *out_pos, int64_t *out_counts, cudaStream_t st, cm_flat_stats *stats) {
    WsScope ws(st);
    if (k_eff > 4096) return search_exact_bigk(qp, nq, k_eff, skip, threshold, out_stride, out_ids, out_scores, out_pos, out_counts, st, stats);
    bool fma = rounding_mode() == CM_ROUND_FMA;
    ScanLaunch L;
    CM_TRY(plan_scan(metric, fma, (int)std::min<int64_t>(nq, SCAN_MAX_QB), ld, n, (int)k_eff, &L));
    int64_t nq_run = (nq + L.qb - 1) / L.qb * L.qb;   // <= nq_pad
    (void)nq_pad;
    // big corpus: one pass (launch) per query group streams all rows; small table (fewer tiles than the
    // persistent grid): many query groups share one launch, otherwise most SMs would idle
    const int64_t n_tiles = (n + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    const int64_t n_groups = nq_run / L.qb;
    const int64_t per_launch = n_tiles >= 2 * (int64_t)sm_count() ? 1 : std::min<int64_t>(n_groups, 32768);
    // ... and with many groups in the launch a group gets only as many CTAs as keeps the launch at about one wave: a
    // CTA that fetches its 8 queries for a single 128-row tile spends its time on the fetch and on the 32-way merge
    // behind it (IVF / IVFPQ coarse step: 512 queries x 4096 centroids went from 2048 one-tile CTAs to 256 of 8 tiles)
    if (per_launch > 1) L.grid = (int)std::max<int64_t>(1, std::min<int64_t>(L.grid, L.slots / std::min(per_launch, n_groups)));
    uint64_t *pk = nullptr;
    int *pc = nullptr;
    CM_TRY(ws.get(&pk, (size_t)nq_run * L.grid * L.K * 8));
    CM_TRY(ws.get(&pc, (size_t)nq_run * L.grid * sizeof(int)));
    int passes = 0;
    for (int64_t g0 = 0; g0 < n_groups; g0 += per_launch, passes++) {
        int64_t q0 = g0 * L.qb;
        CM_TRY(launch_flat_scan(L, tmap, qp + (size_t)q0 * ld, ld, n, skip, threshold, pk + (size_t)q0 * L.grid * L.K,
                                pc + (size_t)q0 * L.grid, (int)std::min(per_launch, n_groups - g0), st));
    }
    CM_TRY(launch_merge_topk(pk, pc, (int)nq, L.grid, L.K, (int)k_eff, ids, out_stride, out_ids, out_scores, out_pos,
                             out_counts, st));
    stats->path_used = CM_PATH_EXACT;
    stats->passes = passes;
    return CM_OK;
}


// stored (already preprocessed) rows appended as they are
int FlatIndex::load_stored_rows(const uint32_t *ids_h, const float *rows_h, int64_t n_add) {
    cudaStream_t st;
    CM_TRY(acquire_stream(&st));
    const int64_t slab = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)dim * 4));
    float *stage = nullptr;
    int rc = ws_alloc((void **)&stage, (size_t)std::min(slab, n_add) * dim * 4, st);
    if (rc == CM_OK) rc = reserve(n + n_add);
    const bool was_raw = raw_rows;
    raw_rows = true;
    for (int64_t i0 = 0; rc == CM_OK && i0 < n_add; i0 += slab) {
        int64_t m = std::min(slab, n_add - i0);
        cudaError_t e = cudaMemcpyAsync(stage, rows_h + (size_t)i0 * dim, (size_t)m * dim * 4, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = fail(CM_ERR_CUDA, "load_rows: %s", cudaGetErrorString(e)); break; }
        rc = add_from_device(ids_h + i0, stage, m, nullptr, st);
    }
    raw_rows = was_raw;
    ws_free(stage, st);
    cudaStreamSynchronize(st);
    release_stream(st);
    return rc;
}

// back to the state NewFlatIndex leaves: no vectors, nothing deleted (buffers are kept)
int FlatIndex::reset() {
    if (deleted && cap > 0) CM_CUDA(cudaMemset(deleted, 0, (size_t)cap));
    n = 0;
    n_deleted_rows = 0;
    ids_host_mirror.clear();
    deleted_ids.clear();
    shadow_rows = 0;
    sorted_n = -1;
    return CM_OK;
}

// stored rows [first, first + m) to a dense host buffer (m x dim)
int FlatIndex::read_rows(int64_t first, int64_t m, float *out) const {
    if (m <= 0) return CM_OK;
    CM_CUDA(cudaMemcpy2D(out, (size_t)dim * 4, rows + (size_t)first * ld, (size_t)ld * 4, (size_t)dim * 4, (size_t)m, cudaMemcpyDeviceToHost));
    return CM_OK;
}

// FlatIndex.WriteTo (flat_index.go:366-470): Flush, then magic "FLAT", version, dim, distance kind, vector count,
// per vector (ID, dimension, data), roaring blob of the (now empty) deleted set.
int FlatIndex::save(wire::Sink &s) {
    CM_TRY(flush());
    CM_TRY(wire::write_header(s, "FLAT", dim, metric));
    CM_WIRE_PUT(s.u32((uint32_t)n), "vector count");
    const int64_t slab = std::max<int64_t>(1, (int64_t)(64u << 20) / ((int64_t)dim * 4));
    std::vector<float> host((size_t)std::min(slab, std::max<int64_t>(n, 1)) * dim);
    std::vector<uint8_t> rec;
    for (int64_t i0 = 0; i0 < n; i0 += slab) {
        const int64_t m = std::min(slab, n - i0);
        CM_TRY(read_rows(i0, m, host.data()));
        const size_t per = 8 + (size_t)dim * 4;
        rec.resize((size_t)m * per);
        for (int64_t i = 0; i < m; i++) {
            uint8_t *r = rec.data() + (size_t)i * per;
            const uint32_t id = ids_host_mirror[(size_t)(i0 + i)], d = (uint32_t)dim;
            memcpy(r, &id, 4);
            memcpy(r + 4, &d, 4);
            memcpy(r + 8, &host[(size_t)i * dim], (size_t)dim * 4);
        }
        CM_WIRE_PUT(s.put(rec.data(), rec.size()), "vector data");
    }
    CM_WIRE_PUT(wire::write_empty_bitmap(s), "bitmap");
    return CM_OK;
}

// FlatIndex.ReadFrom (flat_index.go:488-614): the whole stream is decoded and validated first; only then does the
// index state change (vectors and deleted set are REPLACED, as in the reference).
int FlatIndex::load(wire::Source &s) {
    CM_TRY(wire::read_header(s, "FLAT", dim, metric));
    uint32_t count = 0;
    CM_WIRE_GET(s.u32(&count), "vector count");
    std::vector<uint32_t> ids_h(count);
    std::vector<float> rows_h((size_t)count * dim);
    for (uint32_t i = 0; i < count; i++) {
        uint32_t vd = 0;
        CM_WIRE_GET(s.u32(&ids_h[i]), "vector ID");
        CM_WIRE_GET(s.u32(&vd), "vector dimension");
        if ((int64_t)vd != dim) return fail(CM_ERR_DIM_MISMATCH, "vector %u has dimension %u, expected %d", i, vd, dim);
        CM_WIRE_GET(s.get(&rows_h[(size_t)i * dim], (size_t)dim * 4), "vector data");
    }
    std::vector<uint32_t> dead;
    CM_TRY(wire::read_bitmap(s, &dead));
    CM_TRY(reset());
    if (count > 0) CM_TRY(load_stored_rows(ids_h.data(), rows_h.data(), (int64_t)count));
    for (uint32_t id : dead)
        if (remove(id) != CM_OK) deleted_ids.insert(id);      // an ID the stream marks deleted without holding it: kept in the set
    return CM_OK;
}

}  // namespace cm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using cm::FlatIndex;

struct cm_flat {
    FlatIndex ix;
};

extern "C" {

int cm_distance_pairs(int metric, const float *a, const float *b, int64_t n, int dim, float *out) {
    CM_TRY(cm::ensure_device());
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    if (n <= 0) return CM_OK;
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    size_t bytes = (size_t)n * (dim > 0 ? dim : 1) * 4;
    int rc = cm::ws_alloc((void **)&da, bytes, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&db, bytes, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dout, (size_t)n * 4, st);
    if (rc == CM_OK) {
        if (dim > 0) {
            cudaMemcpyAsync(da, a, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(db, b, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st);
        }
        rc = cm::launch_distance_pairs(metric, cm::rounding_mode() == CM_ROUND_FMA, da, db, n, dim, dout, st);
        cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(da, st); cm::ws_free(db, st); cm::ws_free(dout, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "distance_pairs: %s", cudaGetErrorString(e));
    return rc;
}

int cm_preprocess_rows(int metric, float *rows, int64_t n, int dim, int64_t *bad_row) {
    CM_TRY(cm::ensure_device());
    if (bad_row) *bad_row = -1;
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    if (n <= 0 || metric != CM_COSINE) return CM_OK;   // no-op for l2 / l2_squared (distance.go:139-150)
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *d = nullptr;
    int *flags = nullptr;
    int rc = cm::ws_alloc((void **)&d, (size_t)n * dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&flags, (size_t)n * sizeof(int), st);
    std::vector<int> hf((size_t)n, 0);
    std::vector<float> tmp((size_t)n * dim);
    if (rc == CM_OK) {
        cudaMemcpyAsync(d, rows, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st);
        rc = cm::launch_preprocess_rows(metric, cm::rounding_mode() == CM_ROUND_FMA, d, n, dim, dim, d, dim, flags, st);
        cudaMemcpyAsync(tmp.data(), d, (size_t)n * dim * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(hf.data(), flags, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(d, st); cm::ws_free(flags, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc != CM_OK) return rc;
    if (e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "preprocess_rows: %s", cudaGetErrorString(e));
    int64_t good = n;
    for (int64_t i = 0; i < n; i++) if (hf[(size_t)i]) { good = i; break; }
    memcpy(rows, tmp.data(), (size_t)good * dim * 4);
    if (good < n) {
        if (bad_row) *bad_row = good;
        return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (row %lld)", (long long)good);
    }
    return CM_OK;
}

int cm_flat_create(int dim, int metric, cm_flat **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");           // flat_index.go:120-122
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind"); // distance.go:12
    CM_TRY(cm::ensure_device());
    cm_flat *h = new cm_flat();
    h->ix.dim = dim;
    h->ix.ld = (dim + cm::SCAN_CHUNK - 1) / cm::SCAN_CHUNK * cm::SCAN_CHUNK;
    h->ix.metric = metric;
    if (const char *cg = getenv("COMET_B200_CTA_GROUP")) h->ix.tensor_cta_group = atoi(cg) == 1 ? 1 : 2;
    cudaGetDevice(&h->ix.device);
    *out = h;
    return CM_OK;
}
int cm_flat_destroy(cm_flat *h) {
    delete h;
    return CM_OK;
}
int cm_flat_reserve(cm_flat *h, int64_t n_rows) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.reserve(n_rows);
}
int cm_flat_add(cm_flat *h, const uint32_t *ids, float *rows, int64_t n, int writeback) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    int rc = CM_OK;
    // upload in slabs so a 3 GB Add does not need a second 3 GB staging buffer
    const int64_t slab = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)h->ix.dim * 4));
    float *stage = nullptr;
    rc = cm::ws_alloc((void **)&stage, (size_t)std::min(slab, n) * h->ix.dim * 4, st);
    if (rc == CM_OK) rc = h->ix.reserve(h->ix.n + n);
    for (int64_t i0 = 0; rc == CM_OK && i0 < n; i0 += slab) {
        int64_t m = std::min(slab, n - i0);
        cudaError_t e = cudaMemcpyAsync(stage, rows + (size_t)i0 * h->ix.dim, (size_t)m * h->ix.dim * 4, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = cm::fail(CM_ERR_CUDA, "add: %s", cudaGetErrorString(e)); break; }
        rc = h->ix.add_from_device(ids + i0, stage, m, writeback ? rows + (size_t)i0 * h->ix.dim : nullptr, st);
    }
    cm::ws_free(stage, st);
    cudaStreamSynchronize(st);
    cm::release_stream(st);
    return rc;
}
// FlatIndex.ReadFrom (flat_index.go:488-614) restores vectors that were preprocessed when they were first
// added: they are stored as they are (normalising a unit vector again could move its last bit).
int cm_flat_load_rows(cm_flat *h, const uint32_t *ids, const float *rows, int64_t n) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.load_stored_rows(ids, rows, n);
}
// WriteTo / ReadFrom (flat_index.go:366-614) on caller bytes or a file (".gz" = gzip, storage_provider.go:163-166)
int cm_flat_save(cm_flat *h, uint8_t *buf, int64_t cap, int64_t *bytes) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_buffer([&](cm::wire::Sink &s) { return h->ix.save(s); }, buf, cap, bytes);
}
int cm_flat_load(cm_flat *h, const uint8_t *buf, int64_t len, int64_t *consumed) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_buffer([&](cm::wire::Source &s) { return h->ix.load(s); }, buf, len, consumed);
}
int cm_flat_save_file(cm_flat *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_file([&](cm::wire::Sink &s) { return h->ix.save(s); }, path);
}
int cm_flat_load_file(cm_flat *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_file([&](cm::wire::Source &s) { return h->ix.load(s); }, path);
}
int cm_flat_get_ids(const cm_flat *h, int64_t first, int64_t n, uint32_t *out) {
    if (!h || first < 0 || n < 0 || first + n > h->ix.n || (n > 0 && !out)) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n > 0) memcpy(out, h->ix.ids_host_mirror.data() + first, (size_t)n * 4);
    return CM_OK;
}
int cm_flat_add_device(cm_flat *h, const uint32_t *ids_host, const float *rows_dev, int64_t n, void *stream) {
    if (!h || (n > 0 && (!ids_host || !rows_dev))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.add_from_device(ids_host, rows_dev, n, nullptr, (cudaStream_t)stream);
}
int cm_flat_remove(cm_flat *h, uint32_t id) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.remove(id);
}
int cm_flat_flush(cm_flat *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.flush();
}
int64_t cm_flat_size(const cm_flat *h) { return h ? h->ix.n : 0; }
int cm_flat_dim(const cm_flat *h) { return h ? h->ix.dim : 0; }
int cm_flat_metric(const cm_flat *h) { return h ? h->ix.metric : -1; }

int cm_flat_get_rows(const cm_flat *h, const int64_t *positions, int64_t n, float *out) {
    if (!h || (n > 0 && (!positions || !out))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n <= 0) return CM_OK;
    for (int64_t i = 0; i < n; i++)
        if (positions[i] < 0 || positions[i] >= h->ix.n) return cm::fail(CM_ERR_NOT_FOUND, "position %lld out of range", (long long)positions[i]);
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    int64_t *dpos = nullptr;
    float *dout = nullptr;
    int rc = cm::ws_alloc((void **)&dpos, (size_t)n * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dout, (size_t)n * h->ix.dim * 4, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dpos, positions, (size_t)n * 8, cudaMemcpyHostToDevice, st);
        rc = cm::launch_gather_rows(h->ix.rows, h->ix.ld, h->ix.dim, dpos, n, dout, st);
        cudaMemcpyAsync(out, dout, (size_t)n * h->ix.dim * 4, cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(dpos, st); cm::ws_free(dout, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "get_rows: %s", cudaGetErrorString(e));
    return rc;
}

int cm_flat_get_vector(const cm_flat *h, uint32_t id, float *out) {
    // flat_index_search.go:171-196 lookupNodeVectors: first row with this ID; deleted -> not found
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    for (int64_t i = 0; i < h->ix.n; i++) {
        if (h->ix.ids_host_mirror[(size_t)i] == id) {
            if (h->ix.deleted_ids.count(id)) return cm::fail(CM_ERR_NOT_FOUND, "node ID %u not found in index (deleted)", id);
            return cm_flat_get_rows(h, &i, 1, out);
        }
    }
    return cm::fail(CM_ERR_NOT_FOUND, "node ID %u not found in index", id);
}

int cm_flat_search_device(cm_flat *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                          int64_t *out_counts_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim)
        return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.search_device(queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                               out_counts_dev, (cudaStream_t)stream, false);
}

int cm_flat_search(cm_flat *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                   int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim)
        return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    if (nq <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::tl_zero_n = 0;
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *dq = nullptr, *dsc = nullptr;
    uint32_t *dids = nullptr;
    int64_t *dpos = nullptr, *dcnt = nullptr;
    size_t no = (size_t)nq * (size_t)(out_stride > 0 ? out_stride : 1);
    int rc = cm::ws_alloc((void **)&dq, (size_t)nq * dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dids, no * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dsc, no * 4, st);
    if (rc == CM_OK && out_pos) rc = cm::ws_alloc((void **)&dpos, no * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dcnt, (size_t)nq * 8, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dq, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st);
        rc = h->ix.search_device(dq, nq, p, out_stride, dids, dsc, dpos, dcnt, st, true);
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(out_ids, dids, no * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_scores, dsc, no * 4, cudaMemcpyDeviceToHost, st);
        if (out_pos) cudaMemcpyAsync(out_pos, dpos, no * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, dcnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(dq, st); cm::ws_free(dids, st); cm::ws_free(dsc, st); cm::ws_free(dpos, st); cm::ws_free(dcnt, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "flat_search: %s", cudaGetErrorString(e));
    if (rc != CM_OK) return rc;
    if (h->ix.metric == CM_COSINE) {                          // distance.go:269-290: Preprocess fails on a zero query
        int64_t z = cm::first_zero_query();
        cm::tl_zero_n = 0;
        if (z >= 0) return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)z);
    }
    // tensor path: a query whose candidate list overflowed (count -1) is redone by the exact scan
    std::vector<int64_t> redo;
    for (int64_t q = 0; q < nq; q++)
        if (out_counts[q] < 0) redo.push_back(q);
    if (!redo.empty()) {
        cm_search_params pe = *p;
        pe.path = CM_PATH_EXACT;
        size_t m = redo.size();
        std::vector<float> rq(m * (size_t)dim);
        std::vector<uint32_t> rid(m * (size_t)out_stride);
        std::vector<float> rsc(m * (size_t)out_stride);
        std::vector<int64_t> rpos(out_pos ? m * (size_t)out_stride : 0), rcnt(m);
        for (size_t i = 0; i < m; i++) memcpy(&rq[i * dim], queries + (size_t)redo[i] * dim, (size_t)dim * 4);
        CM_TRY(cm_flat_search(h, rq.data(), (int64_t)m, dim, &pe, out_stride, rid.data(), rsc.data(),
                              out_pos ? rpos.data() : nullptr, rcnt.data()));
        for (size_t i = 0; i < m; i++) {
            size_t o = (size_t)redo[i] * out_stride;
            memcpy(out_ids + o, &rid[i * out_stride], (size_t)out_stride * 4);
            memcpy(out_scores + o, &rsc[i * out_stride], (size_t)out_stride * 4);
            if (out_pos) memcpy(out_pos + o, &rpos[i * out_stride], (size_t)out_stride * 8);
            out_counts[redo[i]] = rcnt[i];
        }
        std::lock_guard<std::mutex> lk(h->ix.stats_mu);
        h->ix.last_stats.path_used = CM_PATH_TENSOR;
        h->ix.last_stats.fallback_queries = (int64_t)m;
    }
    return CM_OK;
}

int cm_merge_shards_device(const uint32_t *ids_dev, const float *scores_dev, const int64_t *counts_dev, int world,
                           int64_t nq, int64_t in_stride, int64_t k, int64_t out_stride, uint32_t *out_ids_dev,
                           float *out_scores_dev, int64_t *out_counts_dev, void *stream) {
    CM_TRY(cm::ensure_device());
    if (!ids_dev || !scores_dev || !out_ids_dev || !out_scores_dev || world <= 0 || in_stride <= 0)
        return cm::fail(CM_ERR_INVALID_ARG, "bad argument");
    if (k <= 0 || k > (int64_t)world * in_stride) k = (int64_t)world * in_stride;
    if (out_stride < k) return cm::fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < k %lld", (long long)out_stride, (long long)k);
    return cm::launch_merge_shards(ids_dev, scores_dev, counts_dev, world, nq, in_stride, (int)k, out_stride, out_ids_dev,
                                   out_scores_dev, out_counts_dev, (cudaStream_t)stream);
}

int cm_flat_last_stats(const cm_flat *h, cm_flat_stats *out) {
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(const_cast<cm_flat *>(h)->ix.stats_mu);
    *out = h->ix.last_stats;
    if (out->candidates < 0) {       // tensor path: the count lives on the device until somebody asks
        unsigned long long c = 0;
        out->candidates = 0;
        if (h->ix.rescored_dev && cudaMemcpy(&c, h->ix.rescored_dev, 8, cudaMemcpyDeviceToHost) == cudaSuccess)
            out->candidates = (int64_t)c;
    }
    return CM_OK;
}

}  // extern "C"
