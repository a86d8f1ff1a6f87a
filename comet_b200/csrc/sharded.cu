// sharded.cu -- row-sharded FlatIndex over the GPUs of one box, driven by ONE process (the Go host is one
// process): cm_flat_sharded_*.
//
// Layout: shard r lives on device devices[r] and holds the rows with global scan positions
// [r * rows_per_shard, (r + 1) * rows_per_shard) -- rows fill shard 0 first, then shard 1, ... -- so the
// reference's result order (score, scan position) is (score, shard, rank within the shard's list) and the
// per-shard top-K lists merge without ever looking at a row again (flat_index_search.go:277-291).
//
// One search = the path's ONE exchange step.  There is no NCCL in it: a single process owns the devices, so the
// exchange runs over NVLink peer mappings, ordered by CUDA events:
//   leader stream : queries ready (event `start`)
//   shard r stream: wait(start) | cm_flat_search_device on shard r -- with peer access its query-preparation kernel READS
//                   the query block from devices[0] and its result kernels WRITE the [nq][K] ids / scores / counts
//                   straight into slot r of devices[0]'s gather buffers (no copy between search and merge; without
//                   peer access the same bytes move by cudaMemcpyPeerAsync) | record(done[r])
//   leader stream : wait(done[*]) | merge_shards_kernel | (host variant: results to the caller)
// The shards are enqueued concurrently by one persistent worker thread each.
// Exactness is shard-local (every shard re-scores its candidates in reference order before the gather); a query a
// shard could not answer on its tensor path (count -1) is redone on the exact path of every shard.
#include <algorithm>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "shard_worker.cuh"


// what a row shard has to offer: FlatIndex and PQIndex shard the same way (the candidate number of a result IS its store
// position, so (score, shard, rank within the shard's list) is the single index's order for both)
struct ShardOps {
    int64_t (*size)(const void *);
    int (*add)(void *, const uint32_t *, float *, int64_t, int);
    int (*remove)(void *, uint32_t);
    int (*flush)(void *);
    int (*search_device)(void *, const float *, int64_t, int, const cm_search_params *, int64_t, uint32_t *, float *, int64_t *, void *);
    int (*destroy)(void *);
};
static const ShardOps FLAT_OPS = {
    [](const void *h) { return cm_flat_size((const cm_flat *)h); },
    [](void *h, const uint32_t *ids, float *rows, int64_t n, int wb) { return cm_flat_add((cm_flat *)h, ids, rows, n, wb); },
    [](void *h, uint32_t id) { return cm_flat_remove((cm_flat *)h, id); },
    [](void *h) { return cm_flat_flush((cm_flat *)h); },
    [](void *h, const float *q, int64_t nq, int dim, const cm_search_params *p, int64_t stride, uint32_t *ids, float *sc, int64_t *cnt, void *st) {
        return cm_flat_search_device((cm_flat *)h, q, nq, dim, p, stride, ids, sc, nullptr, cnt, st);
    },
    [](void *h) { return cm_flat_destroy((cm_flat *)h); },
};
static const ShardOps PQ_OPS = {
    [](const void *h) { return cm_pq_size((const cm_pq *)h); },
    [](void *h, const uint32_t *ids, float *rows, int64_t n, int wb) { return cm_pq_add((cm_pq *)h, ids, rows, n, wb); },
    [](void *h, uint32_t id) { return cm_pq_remove((cm_pq *)h, id); },
    [](void *h) { return cm_pq_flush((cm_pq *)h); },
    [](void *h, const float *q, int64_t nq, int dim, const cm_search_params *p, int64_t stride, uint32_t *ids, float *sc, int64_t *cnt, void *st) {
        return cm_pq_search_device((cm_pq *)h, q, nq, dim, p, stride, ids, sc, nullptr, cnt, st);
    },
    [](void *h) { return cm_pq_destroy((cm_pq *)h); },
};

struct cm_flat_sharded {
    int dim = 0, metric = 0;
    int64_t rows_per_shard = 0, n = 0;
    int cur = 0;                           // shard that receives the next Add (scan order = shard order)
    const ShardOps *ops = &FLAT_OPS;
    int64_t pq_cb_floats = 0;              // PQ shards: floats of the codebooks (M x Ksub x dsub = 2^nbits x dim)
    std::vector<int> dev;
    std::vector<void *> shard;             // cm_flat * or cm_pq * (see ops)
    std::vector<cudaStream_t> st;          // one stream per shard, on its device
    std::vector<char> direct;              // shard r's kernels read the queries from and write their lists into devices[0]'s memory (peer access)
    std::vector<cudaEvent_t> done;         // shard r's results are in the leader's gather buffer
    std::vector<cudaEvent_t> t_begin, t_searched;   // timing: shard r's stream before / after its local search
    cudaEvent_t t_merge0 = nullptr, t_merge1 = nullptr;   // leader: around the merge kernel
    bool timed = false;                    // the events above hold a complete search
    cudaEvent_t start = nullptr;           // leader: queries are ready
    // per-search buffers, grown on demand (shard side: on its device; gather side: on the leader)
    struct Buf {
        float *q = nullptr; uint32_t *ids = nullptr; float *sc = nullptr; int64_t *cnt = nullptr;
        int64_t cap_q = 0, cap_o = 0, cap_n = 0;
    };
    std::vector<Buf> buf;
    uint32_t *g_ids = nullptr; float *g_sc = nullptr; int64_t *g_cnt = nullptr;      // [W][nq][K], [W][nq]
    uint32_t *m_ids = nullptr; float *m_sc = nullptr; int64_t *m_cnt = nullptr; float *q_lead = nullptr;
    int64_t cap_g = 0, cap_m = 0, cap_ql = 0, cap_gc = 0;
    int64_t bytes_exchanged = 0;           // peer bytes of the last search (queries out + lists back)
    std::mutex search_mu;                  // searches share the gather / merge buffers
    std::vector<std::unique_ptr<ShardWorker>> workers;   // [W] (entry 0 unused: shard 0 is enqueued by the caller); empty = single-threaded
};

namespace cm {

static int grow(void **p, int64_t *cap, int64_t want, size_t elem) {
    if (want <= *cap) return CM_OK;
    cudaFree(*p);
    *p = nullptr;
    CM_CUDA(cudaMalloc(p, (size_t)want * elem));
    *cap = want;
    return CM_OK;
}

}  // namespace cm

extern "C" {

static int row_shards_create(int dim, int metric, const int *devices, int n_devices, int64_t rows_per_shard, const ShardOps *ops,
                             const std::function<int(void **)> &make_shard, cm_flat_sharded **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    if (!devices || n_devices <= 0 || n_devices > 64 || rows_per_shard <= 0)
        return cm::fail(CM_ERR_INVALID_ARG, "need 1..64 devices and a positive shard size");
    CM_TRY(cm::ensure_device());
    int n_dev = 0;
    CM_CUDA(cudaGetDeviceCount(&n_dev));
    for (int r = 0; r < n_devices; r++)
        if (devices[r] < 0 || devices[r] >= n_dev) return cm::fail(CM_ERR_INVALID_ARG, "no CUDA device %d", devices[r]);
    int prev = 0;
    cudaGetDevice(&prev);
    cm_flat_sharded *h = new cm_flat_sharded();
    h->dim = dim; h->metric = metric; h->rows_per_shard = rows_per_shard; h->ops = ops;
    h->dev.assign(devices, devices + n_devices);
    h->shard.resize((size_t)n_devices, nullptr);
    h->st.resize((size_t)n_devices, nullptr);
    h->done.resize((size_t)n_devices, nullptr);
    h->t_begin.resize((size_t)n_devices, nullptr);
    h->t_searched.resize((size_t)n_devices, nullptr);
    h->buf.resize((size_t)n_devices);
    h->direct.assign((size_t)n_devices, 1);
    int rc = CM_OK;
    for (int r = 0; r < n_devices && rc == CM_OK; r++) {
        cudaSetDevice(devices[r]);
        rc = make_shard(&h->shard[(size_t)r]);
        if (rc != CM_OK) break;
        if (cudaStreamCreateWithFlags(&h->st[(size_t)r], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&h->done[(size_t)r]) != cudaSuccess || cudaEventCreate(&h->t_begin[(size_t)r]) != cudaSuccess ||
            cudaEventCreate(&h->t_searched[(size_t)r]) != cudaSuccess)
            rc = cm::fail(CM_ERR_CUDA, "stream / event creation on device %d failed", devices[r]);
        // the exchange is peer-to-peer: leader <-> every other device
        if (rc == CM_OK && devices[r] != devices[0]) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[r], devices[0]);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = cm::fail(CM_ERR_CUDA, "peer access %d -> %d: %s", devices[r], devices[0], cudaGetErrorString(e));
                cudaGetLastError();
                cudaSetDevice(devices[0]);
                e = cudaDeviceEnablePeerAccess(devices[r], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = cm::fail(CM_ERR_CUDA, "peer access %d -> %d: %s", devices[0], devices[r], cudaGetErrorString(e));
                cudaGetLastError();
            } else {
                h->direct[(size_t)r] = 0;   // no peer access: queries and lists move by cudaMemcpyPeerAsync (staged by the driver)
            }
            if (getenv("COMET_B200_SHARD_COPIES")) h->direct[(size_t)r] = 0;
        }
    }
    if (rc == CM_OK) {
        cudaSetDevice(devices[0]);
        if (cudaEventCreateWithFlags(&h->start, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&h->t_merge0) != cudaSuccess ||
            cudaEventCreate(&h->t_merge1) != cudaSuccess)
            rc = cm::fail(CM_ERR_CUDA, "event creation failed");
    }
    cudaSetDevice(prev);
    if (rc != CM_OK) { cm_flat_sharded_destroy(h); return rc; }
    const char *thr = getenv("COMET_B200_SHARD_THREADS");
    if (n_devices > 1 && !(thr && atoi(thr) == 0)) {
        h->workers.resize((size_t)n_devices);
        for (int r = 1; r < n_devices; r++) h->workers[(size_t)r].reset(new ShardWorker(devices[r]));
    }
    *out = h;
    return CM_OK;
}

int cm_flat_sharded_create(int dim, int metric, const int *devices, int n_devices, int64_t rows_per_shard,
                           cm_flat_sharded **out) {
    return row_shards_create(dim, metric, devices, n_devices, rows_per_shard, &FLAT_OPS,
                             [=](void **sh) { return cm_flat_create(dim, metric, (cm_flat **)sh); }, out);
}

int cm_flat_sharded_destroy(cm_flat_sharded *h) {
    if (!h) return CM_OK;
    h->workers.clear();            // joins the worker threads
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t r = 0; r < h->dev.size(); r++) {
        cudaSetDevice(h->dev[r]);
        if (h->st[r]) { cudaStreamSynchronize(h->st[r]); cudaStreamDestroy(h->st[r]); }
        if (h->done[r]) cudaEventDestroy(h->done[r]);
        if (h->t_begin[r]) cudaEventDestroy(h->t_begin[r]);
        if (h->t_searched[r]) cudaEventDestroy(h->t_searched[r]);
        cudaFree(h->buf[r].q); cudaFree(h->buf[r].ids); cudaFree(h->buf[r].sc); cudaFree(h->buf[r].cnt);
        if (h->shard[r]) h->ops->destroy(h->shard[r]);
    }
    if (!h->dev.empty()) {
        cudaSetDevice(h->dev[0]);
        if (h->start) cudaEventDestroy(h->start);
        if (h->t_merge0) cudaEventDestroy(h->t_merge0);
        if (h->t_merge1) cudaEventDestroy(h->t_merge1);
        cudaFree(h->g_ids); cudaFree(h->g_sc); cudaFree(h->g_cnt); cudaFree(h->m_ids); cudaFree(h->m_sc); cudaFree(h->m_cnt);
        cudaFree(h->q_lead);
    }
    cudaSetDevice(prev);
    delete h;
    return CM_OK;
}

int cm_flat_sharded_shards(const cm_flat_sharded *h) { return h ? (int)h->dev.size() : 0; }
int64_t cm_flat_sharded_size(const cm_flat_sharded *h) { return h ? h->n : 0; }
int64_t cm_flat_sharded_last_exchange_bytes(const cm_flat_sharded *h) { return h ? h->bytes_exchanged : 0; }

// Rows go to shard `cur` until it holds rows_per_shard rows, then to the next one: scan order == (shard, position
// within the shard), also after a Flush has shortened earlier shards.
static int sharded_room(cm_flat_sharded *h) {
    while (h->cur < (int)h->dev.size() && h->ops->size(h->shard[(size_t)h->cur]) >= h->rows_per_shard) h->cur++;
    return h->cur < (int)h->dev.size() ? CM_OK
                                       : cm::fail(CM_ERR_UNSUPPORTED, "sharded index is full: %lld rows in %d shards of %lld",
                                                  (long long)h->n, (int)h->dev.size(), (long long)h->rows_per_shard);
}

// n successive FlatIndex.Add calls (flat_index.go:169-189)
int cm_flat_sharded_add(cm_flat_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    int64_t done = 0;
    int rc = CM_OK;
    while (done < n && rc == CM_OK) {
        rc = sharded_room(h);
        if (rc != CM_OK) break;
        void *sh = h->shard[(size_t)h->cur];
        const int64_t before = h->ops->size(sh);
        const int64_t m = std::min(h->rows_per_shard - before, n - done);
        cudaSetDevice(h->dev[(size_t)h->cur]);
        rc = h->ops->add(sh, ids + done, rows + (size_t)done * h->dim, m, writeback);
        const int64_t added = h->ops->size(sh) - before;     // a zero vector under cosine stops the batch at that row
        h->n += added;
        done += added;
    }
    cudaSetDevice(prev);
    return rc;
}

// rows already resident on shard `r`'s device.  Shards may be filled in parallel (one device each) as long as every
// shard before the last non-empty one ends up full before the first search: scan order is shard order.
int cm_flat_sharded_add_device(cm_flat_sharded *h, int r, const uint32_t *ids_host, const float *rows_dev, int64_t n,
                               void *stream) {
    if (!h || r < 0 || r >= (int)h->dev.size()) return cm::fail(CM_ERR_INVALID_ARG, "bad shard");
    if (h->ops != &FLAT_OPS) return cm::fail(CM_ERR_UNSUPPORTED, "device-resident rows are a FlatIndex entry point");
    const int64_t have = cm_flat_size((cm_flat *)h->shard[(size_t)r]);
    if (have + n > h->rows_per_shard) return cm::fail(CM_ERR_UNSUPPORTED, "shard %d would exceed its %lld rows", r, (long long)h->rows_per_shard);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->dev[(size_t)r]);
    int rc = cm_flat_add_device((cm_flat *)h->shard[(size_t)r], ids_host, rows_dev, n, stream);
    cudaSetDevice(prev);
    if (rc == CM_OK) h->n += n;
    return rc;
}

int cm_flat_sharded_reserve(cm_flat_sharded *h, int64_t n_rows) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    for (size_t r = 0; r < h->dev.size() && rc == CM_OK && n_rows > 0; r++) {
        const int64_t m = std::min(n_rows, h->rows_per_shard);
        cudaSetDevice(h->dev[r]);
        if (h->ops == &FLAT_OPS) rc = cm_flat_reserve((cm_flat *)h->shard[r], m);
        n_rows -= m;
    }
    cudaSetDevice(prev);
    return rc;
}

// FlatIndex.Remove (flat_index.go:219-250): the first shard in scan order that holds the ID soft-deletes it
int cm_flat_sharded_remove(cm_flat_sharded *h, uint32_t id) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_ERR_NOT_FOUND;
    for (size_t r = 0; r < h->dev.size(); r++) {
        rc = h->ops->remove(h->shard[r], id);
        if (rc != CM_ERR_NOT_FOUND) break;
    }
    cudaSetDevice(prev);
    return rc;
}

// FlatIndex.Flush (flat_index.go:266-299): every shard compacts itself; relative scan order is unchanged
int cm_flat_sharded_flush(cm_flat_sharded *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    int64_t total = 0;
    for (size_t r = 0; r < h->dev.size() && rc == CM_OK; r++) {
        rc = h->ops->flush(h->shard[r]);
        total += h->ops->size(h->shard[r]);
    }
    if (rc == CM_OK) h->n = total;
    cudaSetDevice(prev);
    return rc;
}

int cm_flat_sharded_shard_size(const cm_flat_sharded *h, int r, int64_t *rows) {
    if (!h || r < 0 || r >= (int)h->dev.size() || !rows) return cm::fail(CM_ERR_INVALID_ARG, "bad shard");
    *rows = h->ops->size(h->shard[(size_t)r]);
    return CM_OK;
}

// statistics of the last search, summed over the shards (path_used: the tensor path if any shard used it)
int cm_flat_sharded_last_stats(const cm_flat_sharded *h, cm_flat_stats *out) {
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    for (size_t r = 0; r < h->dev.size(); r++) {
        cm_flat_stats s;
        if (h->ops != &FLAT_OPS || cm_flat_last_stats((cm_flat *)h->shard[r], &s) != CM_OK) continue;
        out->path_used = std::max(out->path_used, s.path_used);
        out->passes = std::max(out->passes, s.passes);
        out->candidates += s.candidates;
        out->fallback_queries += s.fallback_queries;
        out->kernel_launches += s.kernel_launches;
    }
    return CM_OK;
}

// Device-side durations of the last search (waits for it): the slowest shard's local search (queries in, search), the
// slowest shard's copy of its lists to devices[0], and the merge kernel.
int cm_flat_sharded_last_timing(cm_flat_sharded *h, double *search_ms_max, double *gather_ms_max, double *merge_ms) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->search_mu);
    if (!h->timed) return cm::fail(CM_ERR_INVALID_ARG, "no search has run on this handle");
    double s_max = 0.0, g_max = 0.0;
    float ms = 0.0f;
    CM_CUDA(cudaEventSynchronize(h->t_merge1));
    for (size_t r = 0; r < h->dev.size(); r++) {
        CM_CUDA(cudaEventElapsedTime(&ms, h->t_begin[r], h->t_searched[r]));
        s_max = std::max(s_max, (double)ms);
        CM_CUDA(cudaEventElapsedTime(&ms, h->t_searched[r], h->done[r]));
        g_max = std::max(g_max, (double)ms);
    }
    CM_CUDA(cudaEventElapsedTime(&ms, h->t_merge0, h->t_merge1));
    if (search_ms_max) *search_ms_max = s_max;
    if (gather_ms_max) *gather_ms_max = g_max;
    if (merge_ms) *merge_ms = (double)ms;
    return CM_OK;
}

// Everything shard r does for one search, enqueued on its stream (runs on the calling thread or on the shard's worker)
static int shard_enqueue(cm_flat_sharded *h, int r, const float *q_lead_dev, int64_t nq, const cm_search_params *p, int64_t K) {
    cudaSetDevice(h->dev[(size_t)r]);
    cm_flat_sharded::Buf &b = h->buf[(size_t)r];
    if (!h->direct[(size_t)r] && b.cap_q < nq * h->dim) {
        cudaFree(b.q); b.q = nullptr;
        CM_CUDA(cudaMalloc(&b.q, (size_t)nq * h->dim * 4));
        b.cap_q = nq * h->dim;
    }
    if (!h->direct[(size_t)r] && b.cap_o < nq * K) {
        cudaFree(b.ids); cudaFree(b.sc);
        b.ids = nullptr; b.sc = nullptr;
        b.cap_o = 0;
        CM_CUDA(cudaMalloc(&b.ids, (size_t)nq * K * 4));
        CM_CUDA(cudaMalloc(&b.sc, (size_t)nq * K * 4));
        b.cap_o = nq * K;
    }
    if (!h->direct[(size_t)r] && b.cap_n < nq) {
        cudaFree(b.cnt); b.cnt = nullptr;
        b.cap_n = 0;
        CM_CUDA(cudaMalloc(&b.cnt, (size_t)nq * 8));
        b.cap_n = nq;
    }
    cudaStream_t s = h->st[(size_t)r];
    CM_CUDA(cudaStreamWaitEvent(s, h->start, 0));
    CM_CUDA(cudaEventRecord(h->t_begin[(size_t)r], s));
    // With peer access the shard's kernels work on devices[0]'s memory themselves: the query preparation reads the
    // query block over NVLink and the result emit writes this shard's [nq][K] lists straight into slot r of the
    // gather buffers -- the exchange is part of the kernels, no copy sits between search and merge.
    const bool direct = h->direct[(size_t)r] != 0;
    const float *q_r = q_lead_dev;
    if (!direct) {
        CM_CUDA(cudaMemcpyPeerAsync(b.q, h->dev[(size_t)r], q_lead_dev, h->dev[0], (size_t)nq * h->dim * 4, s));
        q_r = b.q;
    }
    uint32_t *o_ids = direct ? h->g_ids + (size_t)r * nq * K : b.ids;
    float *o_sc = direct ? h->g_sc + (size_t)r * nq * K : b.sc;
    int64_t *o_cnt = direct ? h->g_cnt + (size_t)r * nq : b.cnt;
    const int64_t n_r = h->ops->size(h->shard[(size_t)r]);
    if (n_r == 0) {
        CM_TRY(cm::launch_fill_counts(o_cnt, nq, 0, s));
    } else {
        cm_search_params pr = *p;
        pr.k = std::min<int64_t>(K, n_r);          // a shard can contribute at most K rows to the global top-K
        CM_TRY(h->ops->search_device(h->shard[(size_t)r], q_r, nq, h->dim, &pr, K, o_ids, o_sc, o_cnt, (void *)s));
    }
    CM_CUDA(cudaEventRecord(h->t_searched[(size_t)r], s));
    if (!direct) {       // this shard's lists go to slot r of the leader's gather buffers
        CM_CUDA(cudaMemcpyPeerAsync(h->g_ids + (size_t)r * nq * K, h->dev[0], b.ids, h->dev[(size_t)r], (size_t)nq * K * 4, s));
        CM_CUDA(cudaMemcpyPeerAsync(h->g_sc + (size_t)r * nq * K, h->dev[0], b.sc, h->dev[(size_t)r], (size_t)nq * K * 4, s));
        CM_CUDA(cudaMemcpyPeerAsync(h->g_cnt + (size_t)r * nq, h->dev[0], b.cnt, h->dev[(size_t)r], (size_t)nq * 8, s));
    }
    CM_CUDA(cudaEventRecord(h->done[(size_t)r], s));
    return CM_OK;
}

static int sharded_search_impl(cm_flat_sharded *h, const float *q_lead_dev, int64_t nq, const cm_search_params *p,
                               int64_t K, cudaStream_t lead) {
    // q_lead_dev: queries on the leader device, ready in `lead`'s order.  Leaves the merged result in h->m_*.
    const int W = (int)h->dev.size();
    h->bytes_exchanged = 0;
    cudaSetDevice(h->dev[0]);
    if (h->cap_g < (int64_t)W * nq * K) {
        cudaFree(h->g_ids); cudaFree(h->g_sc);
        h->g_ids = nullptr; h->g_sc = nullptr;
        h->cap_g = 0;
        CM_CUDA(cudaMalloc(&h->g_ids, (size_t)W * nq * K * 4));
        CM_CUDA(cudaMalloc(&h->g_sc, (size_t)W * nq * K * 4));
        h->cap_g = (int64_t)W * nq * K;
    }
    if (h->cap_m < nq * K) {
        cudaFree(h->m_ids); cudaFree(h->m_sc);
        h->m_ids = nullptr; h->m_sc = nullptr;
        h->cap_m = 0;
        CM_CUDA(cudaMalloc(&h->m_ids, (size_t)nq * K * 4));
        CM_CUDA(cudaMalloc(&h->m_sc, (size_t)nq * K * 4));
        h->cap_m = nq * K;
    }
    if (h->cap_gc < nq) {                 // the per-query counts grow with the batch, whatever K is
        cudaFree(h->g_cnt); cudaFree(h->m_cnt);
        h->g_cnt = nullptr; h->m_cnt = nullptr;
        h->cap_gc = 0;
        CM_CUDA(cudaMalloc(&h->g_cnt, (size_t)W * nq * 8));
        CM_CUDA(cudaMalloc(&h->m_cnt, (size_t)nq * 8));
        h->cap_gc = nq;
    }
    CM_CUDA(cudaEventRecord(h->start, lead));
    // The shards are enqueued concurrently: one persistent worker thread per shard beyond the first (a search is ~25
    // launches per shard; one thread enqueueing 8 shards in turn would put 0.3 ms between the first and the last).
    const bool threaded = W > 1 && h->workers.size() == (size_t)W;
    for (int r = 1; r < W && threaded; r++) h->workers[(size_t)r]->post([=] { return shard_enqueue(h, r, q_lead_dev, nq, p, K); });
    int rc = CM_OK;
    for (int r = 0; r < (threaded ? 1 : W) && rc == CM_OK; r++) rc = shard_enqueue(h, r, q_lead_dev, nq, p, K);
    for (int r = 1; r < W && threaded; r++) {
        std::string msg;
        const int rc_r = h->workers[(size_t)r]->wait(&msg);
        if (rc_r != CM_OK && rc == CM_OK) rc = cm::fail(rc_r, "%s", msg.c_str());
    }
    CM_TRY(rc);
    for (int r = 0; r < W; r++)
        if (h->dev[(size_t)r] != h->dev[0]) h->bytes_exchanged += nq * h->dim * 4 + nq * K * 8 + nq * 8;
    cudaSetDevice(h->dev[0]);
    for (int r = 0; r < W; r++) CM_CUDA(cudaStreamWaitEvent(lead, h->done[(size_t)r], 0));
    CM_CUDA(cudaEventRecord(h->t_merge0, lead));
    CM_TRY(cm::launch_merge_shards(h->g_ids, h->g_sc, h->g_cnt, W, nq, K, (int)K, K, h->m_ids, h->m_sc, h->m_cnt, lead));
    CM_CUDA(cudaEventRecord(h->t_merge1, lead));
    h->timed = true;
    return CM_OK;
}

static int64_t sharded_k(const cm_flat_sharded *h, int64_t k) {
    return (k <= 0 || k > h->n) ? h->n : k;              // limiter.go:12-17 on the whole index
}

int cm_flat_sharded_search_device(cm_flat_sharded *h, const float *queries_dev, int64_t nq, int dim,
                                  const cm_search_params *p, int64_t out_stride, uint32_t *out_ids_dev,
                                  float *out_scores_dev, int64_t *out_counts_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->dim, dim);
    if (nq <= 0) return CM_OK;
    const int64_t K = sharded_k(h, p->k);
    if (out_stride < K) return cm::fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)K);
    std::lock_guard<std::mutex> lk(h->search_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaStream_t lead = (cudaStream_t)stream;
    int rc = CM_OK;
    if (K == 0) {
        // empty index: shard 0 (on devices[0]) answers like an empty FlatIndex
        rc = h->ops->search_device(h->shard[0], queries_dev, nq, dim, p, out_stride, out_ids_dev, out_scores_dev, out_counts_dev, stream);
    } else {
        rc = sharded_search_impl(h, queries_dev, nq, p, K, lead);
        if (rc == CM_OK) {
            cudaSetDevice(h->dev[0]);
            cudaMemcpy2DAsync(out_ids_dev, (size_t)out_stride * 4, h->m_ids, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToDevice, lead);
            cudaMemcpy2DAsync(out_scores_dev, (size_t)out_stride * 4, h->m_sc, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToDevice, lead);
            cudaMemcpyAsync(out_counts_dev, h->m_cnt, (size_t)nq * 8, cudaMemcpyDeviceToDevice, lead);
        }
    }
    cudaSetDevice(prev);
    return rc;
}

int cm_flat_sharded_search(cm_flat_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                           int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->dim, dim);
    if (nq <= 0) return CM_OK;
    const int64_t K = sharded_k(h, p->k);
    if (out_stride < K) return cm::fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)K);
    if (K == 0) {        // empty index: shard 0 answers like an empty index of its kind (FlatIndex: a zero query under cosine still fails)
        if (h->ops == &FLAT_OPS) return cm_flat_search((cm_flat *)h->shard[0], queries, nq, dim, p, out_stride, out_ids, out_scores, nullptr, out_counts);
        return cm_pq_search((cm_pq *)h->shard[0], queries, nq, dim, p, out_stride, out_ids, out_scores, nullptr, out_counts);
    }
    std::unique_lock<std::mutex> lk(h->search_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->dev[0]);
    cudaStream_t lead = h->st[0];
    int rc = cm::grow((void **)&h->q_lead, &h->cap_ql, nq * dim, 4);
    if (rc == CM_OK) {
        cudaMemcpyAsync(h->q_lead, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, lead);
        // the shard streams start behind `lead`; shard 0 shares it
        rc = sharded_search_impl(h, h->q_lead, nq, p, K, lead);
    }
    if (rc == CM_OK) {
        cudaSetDevice(h->dev[0]);
        cudaMemcpy2DAsync(out_ids, (size_t)out_stride * 4, h->m_ids, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToHost, lead);
        cudaMemcpy2DAsync(out_scores, (size_t)out_stride * 4, h->m_sc, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToHost, lead);
        cudaMemcpyAsync(out_counts, h->m_cnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, lead);
        cudaError_t e = cudaStreamSynchronize(lead);
        if (e != cudaSuccess) rc = cm::fail(CM_ERR_CUDA, "sharded search: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    lk.unlock();
    if (rc != CM_OK) return rc;
    // per-query conditions the device entry points can only report (see cm_flat_search_device)
    std::vector<int64_t> redo;
    for (int64_t q = 0; q < nq; q++) {
        if (out_counts[q] == -2) return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)q);
        if (out_counts[q] < 0) redo.push_back(q);
    }
    if (!redo.empty() && p->path != CM_PATH_EXACT) {       // candidate overflow on some shard: exact scan everywhere
        cm_search_params pe = *p;
        pe.path = CM_PATH_EXACT;
        const size_t m = redo.size();
        std::vector<float> rq(m * (size_t)dim);
        std::vector<uint32_t> rid(m * (size_t)out_stride);
        std::vector<float> rsc(m * (size_t)out_stride);
        std::vector<int64_t> rcnt(m);
        for (size_t i = 0; i < m; i++) memcpy(&rq[i * dim], queries + (size_t)redo[i] * dim, (size_t)dim * 4);
        CM_TRY(cm_flat_sharded_search(h, rq.data(), (int64_t)m, dim, &pe, out_stride, rid.data(), rsc.data(), rcnt.data()));
        for (size_t i = 0; i < m; i++) {
            const size_t o = (size_t)redo[i] * out_stride;
            memcpy(out_ids + o, &rid[i * out_stride], (size_t)out_stride * 4);
            memcpy(out_scores + o, &rsc[i * out_stride], (size_t)out_stride * 4);
            out_counts[redo[i]] = rcnt[i];
        }
    }
    return CM_OK;
}

// ---- PQIndex row shards: the same driver over cm_pq shards (codebooks replicated: they are M x Ksub x dsub floats) ------
int cm_pq_sharded_create(int dim, int metric, int M, int nbits, const int *devices, int n_devices, int64_t rows_per_shard,
                         cm_pq_sharded **out) {
    int rc = row_shards_create(dim, metric, devices, n_devices, rows_per_shard, &PQ_OPS,
                               [=](void **sh) { return cm_pq_create(dim, metric, M, nbits, (cm_pq **)sh); }, (cm_flat_sharded **)out);
    if (rc == CM_OK) (*(cm_flat_sharded **)out)->pq_cb_floats = ((int64_t)1 << nbits) * dim;
    return rc;
}
int cm_pq_sharded_destroy(cm_pq_sharded *h) { return cm_flat_sharded_destroy((cm_flat_sharded *)h); }
int cm_pq_sharded_shards(const cm_pq_sharded *h) { return cm_flat_sharded_shards((const cm_flat_sharded *)h); }
int64_t cm_pq_sharded_size(const cm_pq_sharded *h) { return cm_flat_sharded_size((const cm_flat_sharded *)h); }
int cm_pq_sharded_trained(const cm_pq_sharded *h) {
    const cm_flat_sharded *s = (const cm_flat_sharded *)h;
    return s && !s->shard.empty() && cm_pq_trained((const cm_pq *)s->shard[0]);
}
// PQIndex.Train (pq_index.go:193-247) on devices[0]; the codebooks then go to every shard
int cm_pq_sharded_train(cm_pq_sharded *h, const float *rows, int64_t n) {
    cm_flat_sharded *s = (cm_flat_sharded *)h;
    if (!s || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (s->n > 0) return cm::fail(CM_ERR_UNSUPPORTED, "retraining a non-empty index is not supported");
    int prev = 0;
    cudaGetDevice(&prev);
    cm_pq *lead = (cm_pq *)s->shard[0];
    int rc = cm_pq_train(lead, rows, n);
    std::vector<float> cb((size_t)s->pq_cb_floats);
    if (rc == CM_OK) rc = cm_pq_get_codebooks(lead, cb.data());
    for (size_t r = 1; r < s->shard.size() && rc == CM_OK; r++) rc = cm_pq_set_codebooks((cm_pq *)s->shard[r], cb.data());
    cudaSetDevice(prev);
    return rc;
}
int cm_pq_sharded_set_codebooks(cm_pq_sharded *h, const float *codebooks) {
    cm_flat_sharded *s = (cm_flat_sharded *)h;
    if (!s || !codebooks) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    for (size_t r = 0; r < s->shard.size() && rc == CM_OK; r++) rc = cm_pq_set_codebooks((cm_pq *)s->shard[r], codebooks);
    cudaSetDevice(prev);
    return rc;
}
int cm_pq_sharded_get_codebooks(const cm_pq_sharded *h, float *out) {
    const cm_flat_sharded *s = (const cm_flat_sharded *)h;
    if (!s || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    return cm_pq_get_codebooks((const cm_pq *)s->shard[0], out);
}
int cm_pq_sharded_add(cm_pq_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback) {
    if (!cm_pq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
    return cm_flat_sharded_add((cm_flat_sharded *)h, ids, rows, n, writeback);
}
int cm_pq_sharded_remove(cm_pq_sharded *h, uint32_t id) { return cm_flat_sharded_remove((cm_flat_sharded *)h, id); }
int cm_pq_sharded_flush(cm_pq_sharded *h) { return cm_flat_sharded_flush((cm_flat_sharded *)h); }
int cm_pq_sharded_search(cm_pq_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                         uint32_t *out_ids, float *out_scores, int64_t *out_counts) {
    if (!cm_pq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index not trained");
    const cm_flat_sharded *s = (const cm_flat_sharded *)h;
    if (s->metric == CM_COSINE && queries && dim == s->dim && s->n > 0)      // Distance.Preprocess fails on a zero query (distance.go:269-290);
        for (int64_t q = 0; q < nq; q++) {                                    // an empty PQ index answers before it (pq_index_search.go:232)
            float ss = 0.0f;
            for (int j = 0; j < dim; j++) ss += queries[(size_t)q * dim + j] * queries[(size_t)q * dim + j];
            if (ss == 0.0f) return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)q);
        }
    return cm_flat_sharded_search((cm_flat_sharded *)h, queries, nq, dim, p, out_stride, out_ids, out_scores, out_counts);
}
int cm_pq_sharded_search_device(cm_pq_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev, void *stream) {
    if (!cm_pq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index not trained");
    return cm_flat_sharded_search_device((cm_flat_sharded *)h, queries_dev, nq, dim, p, out_stride, out_ids_dev, out_scores_dev,
                                         out_counts_dev, stream);
}

}  // extern "C"
