// batcher.cu -- cross-call dynamic batching for the flat index.
//
// The reference's users issue ONE query per Execute() (flat_index_search.go:109-165; WithQuery(q1..qB)
// unions and sums, it is not a batch API -- SURVEY F6), often from many goroutines under the index's
// RLock.  On the device one query costs a full pass over the corpus (HBM bound, ~0.5 ms for 1M x 768)
// while 512 queries cost barely more (tensor path), so concurrent callers are worth coalescing:
// cm_flat_batcher_search() blocks its caller, a worker thread gathers the pending requests that share
// (k, threshold) until `max_batch` are waiting or the oldest has waited `max_wait_us`, runs ONE
// cm_flat_search for them, and hands every caller its own rows of the result.  Results are exactly
// those of a direct call (same kernels, same order); errors (zero query under cosine, CUDA failures)
// are reported per batch to every caller in it.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

struct cm_flat;   // flat.cu

struct BatchRequest {
    const float *query;
    int64_t k;
    float threshold;
    uint32_t *out_ids;
    float *out_scores;
    int64_t *out_count;
    int64_t stride;
    int rc = CM_OK;
    char err[256];
    bool done = false;
    std::chrono::steady_clock::time_point t0;
};

struct cm_flat_batcher {
    cm_flat *index = nullptr;
    int dim = 0;
    int max_batch = 512;
    int max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<BatchRequest *> pending;
    bool stop = false;
    std::thread worker;
    int64_t batches = 0, requests = 0;

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        std::vector<BatchRequest *> batch;
        std::vector<float> q;
        std::vector<uint32_t> ids;
        std::vector<float> sc;
        std::vector<int64_t> cnt;
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !pending.empty(); });
            if (stop && pending.empty()) return;
            // wait for more company until the batch is full or the oldest request is due
            auto due = pending.front()->t0 + std::chrono::microseconds(max_wait_us);
            while (!stop && (int)pending.size() < max_batch && std::chrono::steady_clock::now() < due)
                cv_work.wait_until(lk, due);
            // take the requests that share the head's (k, threshold)
            batch.clear();
            const int64_t k = pending.front()->k;
            const float thr = pending.front()->threshold;
            for (auto it = pending.begin(); it != pending.end() && (int)batch.size() < max_batch;) {
                if ((*it)->k == k && (*it)->threshold == thr) { batch.push_back(*it); it = pending.erase(it); }
                else ++it;
            }
            lk.unlock();
            const int64_t nq = (int64_t)batch.size();
            int64_t stride = batch[0]->stride;
            for (auto *r : batch) stride = std::min(stride, r->stride);
            q.resize((size_t)nq * dim);
            for (int64_t i = 0; i < nq; i++) memcpy(&q[(size_t)i * dim], batch[(size_t)i]->query, (size_t)dim * 4);
            ids.assign((size_t)(nq * stride), 0);
            sc.assign((size_t)(nq * stride), 0.0f);
            cnt.assign((size_t)nq, 0);
            cm_search_params p{};
            p.k = k; p.threshold = thr; p.path = CM_PATH_AUTO;
            int rc = cm_flat_search(index, q.data(), nq, dim, &p, stride, ids.data(), sc.data(), nullptr, cnt.data());
            const char *msg = rc == CM_OK ? "" : cm_last_error();
            for (int64_t i = 0; i < nq; i++) {
                BatchRequest *r = batch[(size_t)i];
                r->rc = rc;
                if (rc == CM_OK) {
                    int64_t m = cnt[(size_t)i];
                    memcpy(r->out_ids, &ids[(size_t)(i * stride)], (size_t)m * 4);
                    memcpy(r->out_scores, &sc[(size_t)(i * stride)], (size_t)m * 4);
                    *r->out_count = m;
                } else {
                    snprintf(r->err, sizeof(r->err), "%s", msg);
                }
            }
            lk.lock();
            batches++;
            requests += nq;
            for (auto *r : batch) r->done = true;
            cv_done.notify_all();
        }
    }
};

extern "C" {

int cm_flat_batcher_create(cm_flat *index, int max_batch, int max_wait_us, cm_flat_batcher **out) {
    if (!index || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    cm_flat_batcher *b = new cm_flat_batcher();
    b->index = index;
    b->dim = cm_flat_dim(index);
    if (max_batch > 0) b->max_batch = max_batch;
    if (max_wait_us >= 0) b->max_wait_us = max_wait_us;
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return CM_OK;
}

int cm_flat_batcher_destroy(cm_flat_batcher *b) {
    if (!b) return CM_OK;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
    return CM_OK;
}

// One searchSingleQuery (flat_index_search.go:221-294) on behalf of the calling thread; blocks until the
// batch it joined has been answered.  out_stride >= min(k, n) like cm_flat_search.
int cm_flat_batcher_search(cm_flat_batcher *b, const float *query, int dim, int64_t k, float threshold, int64_t out_stride,
                           uint32_t *out_ids, float *out_scores, int64_t *out_count) {
    if (!b || !query || !out_ids || !out_scores || !out_count) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != b->dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", b->dim, dim);
    {
        int64_t n = cm_flat_size(b->index), ke = (k <= 0 || k > n) ? n : k;
        if (out_stride < ke) return cm::fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)ke);
    }
    if (cm_flat_metric(b->index) == CM_COSINE) {
        // Distance.Preprocess fails on a zero query (distance.go:269-290) -- for THAT caller only: in the reference every
        // Execute() is on its own, so a zero query must not take the rest of its batch down with it.  norm == 0 exactly
        // when the float32 sum of squares is 0 (every term is >= 0), whatever the summation order or rounding mode.
        float ss = 0.0f;
        for (int j = 0; j < dim; j++) ss += query[j] * query[j];
        if (ss == 0.0f) return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector");
    }
    BatchRequest r;
    r.query = query; r.k = k; r.threshold = threshold; r.out_ids = out_ids; r.out_scores = out_scores; r.out_count = out_count;
    r.stride = out_stride;
    r.t0 = std::chrono::steady_clock::now();
    r.err[0] = 0;
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stop) return cm::fail(CM_ERR_INVALID_ARG, "batcher is shutting down");
    b->pending.push_back(&r);
    b->cv_work.notify_one();
    b->cv_done.wait(lk, [&] { return r.done; });
    lk.unlock();
    if (r.rc != CM_OK) return cm::fail(r.rc, "%s", r.err);
    return CM_OK;
}

int cm_flat_batcher_stats(cm_flat_batcher *b, int64_t *batches, int64_t *requests) {
    if (!b) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(b->mu);
    if (batches) *batches = b->batches;
    if (requests) *requests = b->requests;
    return CM_OK;
}

}  // extern "C"
