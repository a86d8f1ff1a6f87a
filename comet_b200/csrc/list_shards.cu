// list_shards.cu -- the search driver of the list-sharded IVF / IVFPQ indexes (see list_shards.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>

#include "flat_kernels.cuh"
#include "list_shards.cuh"

namespace cm {

int ListShards::init(int dim_, int nlist_, int metric_, const int *devices, int n_devices) {
    if (!devices || n_devices <= 0 || n_devices > 64) return fail(CM_ERR_INVALID_ARG, "need 1..64 devices");
    CM_TRY(ensure_device());
    int n_dev = 0;
    CM_CUDA(cudaGetDeviceCount(&n_dev));
    for (int r = 0; r < n_devices; r++)
        if (devices[r] < 0 || devices[r] >= n_dev) return fail(CM_ERR_INVALID_ARG, "no CUDA device %d", devices[r]);
    dim = dim_; nlist = nlist_; metric = metric_;
    dev.assign(devices, devices + n_devices);
    const size_t w = (size_t)n_devices;
    st.assign(w, nullptr); done.assign(w, nullptr); direct.assign(w, 1);
    glob_len_dev.assign(w, nullptr); buf.resize(w); last_scanned.assign(w, 0);
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    for (int r = 0; r < n_devices && rc == CM_OK; r++) {
        cudaSetDevice(devices[r]);
        if (cudaStreamCreateWithFlags(&st[(size_t)r], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[(size_t)r], cudaEventDisableTiming) != cudaSuccess ||
            cudaMalloc(&glob_len_dev[(size_t)r], (size_t)std::max(nlist, 1) * sizeof(long long)) != cudaSuccess)
            rc = fail(CM_ERR_CUDA, "stream / event / buffer creation on device %d failed", devices[r]);
        if (rc == CM_OK && devices[r] != devices[0]) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[r], devices[0]);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(CM_ERR_CUDA, "peer access %d -> %d: %s", devices[r], devices[0], cudaGetErrorString(e));
                cudaGetLastError();
                cudaSetDevice(devices[0]);
                e = cudaDeviceEnablePeerAccess(devices[r], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(CM_ERR_CUDA, "peer access %d -> %d: %s", devices[0], devices[r], cudaGetErrorString(e));
                cudaGetLastError();
            } else {
                direct[(size_t)r] = 0;
            }
            if (getenv("COMET_B200_SHARD_COPIES")) direct[(size_t)r] = 0;
        }
    }
    if (rc == CM_OK) {
        cudaSetDevice(devices[0]);
        if (cudaEventCreateWithFlags(&start, cudaEventDisableTiming) != cudaSuccess) rc = fail(CM_ERR_CUDA, "event creation failed");
    }
    cudaSetDevice(prev);
    if (rc != CM_OK) return rc;
    owner.resize((size_t)nlist);
    for (int l = 0; l < nlist; l++) owner[(size_t)l] = l % n_devices;
    glob_len.assign((size_t)nlist, 0);
    const char *thr = getenv("COMET_B200_SHARD_THREADS");
    if (n_devices > 1 && !(thr && atoi(thr) == 0)) {
        workers.resize(w);
        for (int r = 1; r < n_devices; r++) workers[(size_t)r].reset(new ShardWorker(devices[r]));
    }
    return CM_OK;
}

void ListShards::destroy() {
    workers.clear();
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t r = 0; r < dev.size(); r++) {
        cudaSetDevice(dev[r]);
        if (r < st.size() && st[r]) { cudaStreamSynchronize(st[r]); cudaStreamDestroy(st[r]); }
        if (r < done.size() && done[r]) cudaEventDestroy(done[r]);
        if (r < glob_len_dev.size()) cudaFree(glob_len_dev[r]);
        if (r < buf.size()) { cudaFree(buf[r].q); cudaFree(buf[r].ids); cudaFree(buf[r].gno); cudaFree(buf[r].sc); cudaFree(buf[r].cnt); }
    }
    if (!dev.empty()) {
        cudaSetDevice(dev[0]);
        if (start) cudaEventDestroy(start);
        cudaFree(g_ids); cudaFree(g_gno); cudaFree(g_sc); cudaFree(g_cnt);
        cudaFree(m_ids); cudaFree(m_sc); cudaFree(m_cnt); cudaFree(q_lead);
    }
    cudaSetDevice(prev);
    st.clear(); done.clear(); glob_len_dev.clear(); buf.clear(); dev.clear();
}

int64_t ListShards::effective_k(const cm_search_params *p) const {
    int nprobes = p->nprobes;
    if (nprobes <= 0 || nprobes > nlist) nprobes = nlist;
    std::vector<long long> len = glob_len;
    std::sort(len.begin(), len.end(), std::greater<long long>());
    int64_t bound = 0;
    for (int i = 0; i < nprobes; i++) bound += len[(size_t)i];
    return (p->k <= 0 || p->k > bound) ? bound : p->k;
}

std::vector<int> ListShards::greedy_plan() const {
    std::vector<int> order((size_t)nlist), want((size_t)nlist);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return glob_len[(size_t)a] > glob_len[(size_t)b]; });
    std::vector<long long> load((size_t)W(), 0);
    for (int l : order) {
        int best = 0;
        for (int r = 1; r < W(); r++) if (load[(size_t)r] < load[(size_t)best]) best = r;
        want[(size_t)l] = best;
        load[(size_t)best] += glob_len[(size_t)l];
    }
    return want;
}

// everything shard r does for one search, enqueued on its stream
static int list_shard_enqueue(ListShards *h, const ListShards::ShardSearch &fn, int r, const float *q_lead_dev, int64_t nq,
                              const cm_search_params *p, int64_t K) {
    cudaSetDevice(h->dev[(size_t)r]);
    ListShards::Buf &b = h->buf[(size_t)r];
    const bool direct = h->direct[(size_t)r] != 0;
    if (!direct && b.cap_q < nq * h->dim) {
        cudaFree(b.q); b.q = nullptr;
        CM_CUDA(cudaMalloc(&b.q, (size_t)nq * h->dim * 4));
        b.cap_q = nq * h->dim;
    }
    if (!direct && b.cap_o < nq * K) {
        cudaFree(b.ids); cudaFree(b.gno); cudaFree(b.sc);
        b.ids = b.gno = nullptr; b.sc = nullptr;
        b.cap_o = 0;
        CM_CUDA(cudaMalloc(&b.ids, (size_t)nq * K * 4));
        CM_CUDA(cudaMalloc(&b.gno, (size_t)nq * K * 4));
        CM_CUDA(cudaMalloc(&b.sc, (size_t)nq * K * 4));
        b.cap_o = nq * K;
    }
    if (!direct && b.cap_n < nq) {
        cudaFree(b.cnt); b.cnt = nullptr;
        b.cap_n = 0;
        CM_CUDA(cudaMalloc(&b.cnt, (size_t)nq * 8));
        b.cap_n = nq;
    }
    cudaStream_t s = h->st[(size_t)r];
    CM_CUDA(cudaStreamWaitEvent(s, h->start, 0));
    const float *q_r = q_lead_dev;
    if (!direct) {
        CM_CUDA(cudaMemcpyPeerAsync(b.q, h->dev[(size_t)r], q_lead_dev, h->dev[0], (size_t)nq * h->dim * 4, s));
        q_r = b.q;
    }
    const size_t slot = (size_t)r * nq * K;
    uint32_t *o_ids = direct ? h->g_ids + slot : b.ids, *o_gno = direct ? h->g_gno + slot : b.gno;
    float *o_sc = direct ? h->g_sc + slot : b.sc;
    int64_t *o_cnt = direct ? h->g_cnt + (size_t)r * nq : b.cnt;
    cm_search_params pr = *p;
    pr.k = K;                 // the shard clamps it to what its own lists can hold
    CM_TRY(fn(r, q_r, nq, &pr, K, o_ids, o_sc, o_cnt, s, h->glob_len_dev[(size_t)r], o_gno));
    if (!direct) {
        CM_CUDA(cudaMemcpyPeerAsync(h->g_ids + slot, h->dev[0], b.ids, h->dev[(size_t)r], (size_t)nq * K * 4, s));
        CM_CUDA(cudaMemcpyPeerAsync(h->g_gno + slot, h->dev[0], b.gno, h->dev[(size_t)r], (size_t)nq * K * 4, s));
        CM_CUDA(cudaMemcpyPeerAsync(h->g_sc + slot, h->dev[0], b.sc, h->dev[(size_t)r], (size_t)nq * K * 4, s));
        CM_CUDA(cudaMemcpyPeerAsync(h->g_cnt + (size_t)r * nq, h->dev[0], b.cnt, h->dev[(size_t)r], (size_t)nq * 8, s));
    }
    CM_CUDA(cudaEventRecord(h->done[(size_t)r], s));
    return CM_OK;
}

int ListShards::search_impl(const ShardSearch &fn, const float *q_lead_dev, int64_t nq, const cm_search_params *p, int64_t K,
                            cudaStream_t lead) {
    const int w = W();
    if ((size_t)w * (size_t)K * 8 > max_smem_optin())
        return fail(CM_ERR_UNSUPPORTED, "%d list shards x k=%lld too large for the shard merge", w, (long long)K);
    if (len_dirty) {
        for (int r = 0; r < w; r++) {
            cudaSetDevice(dev[(size_t)r]);
            CM_CUDA(cudaMemcpy(glob_len_dev[(size_t)r], glob_len.data(), (size_t)nlist * sizeof(long long), cudaMemcpyHostToDevice));
        }
        len_dirty = false;
    }
    cudaSetDevice(dev[0]);
    if (cap_g < (int64_t)w * nq * K) {
        cudaFree(g_ids); cudaFree(g_gno); cudaFree(g_sc);
        g_ids = g_gno = nullptr; g_sc = nullptr;
        cap_g = 0;
        CM_CUDA(cudaMalloc(&g_ids, (size_t)w * nq * K * 4));
        CM_CUDA(cudaMalloc(&g_gno, (size_t)w * nq * K * 4));
        CM_CUDA(cudaMalloc(&g_sc, (size_t)w * nq * K * 4));
        cap_g = (int64_t)w * nq * K;
    }
    if (cap_m < nq * K) {
        cudaFree(m_ids); cudaFree(m_sc);
        m_ids = nullptr; m_sc = nullptr;
        cap_m = 0;
        CM_CUDA(cudaMalloc(&m_ids, (size_t)nq * K * 4));
        CM_CUDA(cudaMalloc(&m_sc, (size_t)nq * K * 4));
        cap_m = nq * K;
    }
    if (cap_nq < nq) {                    // the per-query counts grow with the batch, whatever K is
        cudaFree(g_cnt); cudaFree(m_cnt);
        g_cnt = nullptr; m_cnt = nullptr;
        cap_nq = 0;
        CM_CUDA(cudaMalloc(&g_cnt, (size_t)w * nq * 8));
        CM_CUDA(cudaMalloc(&m_cnt, (size_t)nq * 8));
        cap_nq = nq;
    }
    CM_CUDA(cudaEventRecord(start, lead));
    const bool threaded = w > 1 && workers.size() == (size_t)w;
    ListShards *self = this;
    for (int r = 1; r < w && threaded; r++)
        workers[(size_t)r]->post([=, &fn] { return list_shard_enqueue(self, fn, r, q_lead_dev, nq, p, K); });
    int rc = CM_OK;
    for (int r = 0; r < (threaded ? 1 : w) && rc == CM_OK; r++) rc = list_shard_enqueue(this, fn, r, q_lead_dev, nq, p, K);
    for (int r = 1; r < w && threaded; r++) {
        std::string msg;
        const int rc_r = workers[(size_t)r]->wait(&msg);
        if (rc_r != CM_OK && rc == CM_OK) rc = fail(rc_r, "%s", msg.c_str());
    }
    CM_TRY(rc);
    cudaSetDevice(dev[0]);
    for (int r = 0; r < w; r++) CM_CUDA(cudaStreamWaitEvent(lead, done[(size_t)r], 0));
    return launch_merge_keyed_shards(g_ids, g_sc, g_gno, g_cnt, w, nq, K, (int)K, K, m_ids, m_sc, m_cnt, lead);
}

int ListShards::search_device(const ShardSearch &fn, const float *queries_dev, int64_t nq, const cm_search_params *p,
                              int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev,
                              cudaStream_t lead) {
    std::lock_guard<std::mutex> lk(search_mu);
    const int64_t K = effective_k(p);
    if (out_stride < K) return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)K);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dev[0]);
    int rc = CM_OK;
    if (K == 0) {
        rc = launch_fill_counts(out_counts_dev, nq, 0, lead);
    } else {
        rc = search_impl(fn, queries_dev, nq, p, K, lead);
        if (rc == CM_OK) {
            cudaSetDevice(dev[0]);
            cudaMemcpy2DAsync(out_ids_dev, (size_t)out_stride * 4, m_ids, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToDevice, lead);
            cudaMemcpy2DAsync(out_scores_dev, (size_t)out_stride * 4, m_sc, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToDevice, lead);
            cudaMemcpyAsync(out_counts_dev, m_cnt, (size_t)nq * 8, cudaMemcpyDeviceToDevice, lead);
        }
    }
    cudaSetDevice(prev);
    return rc;
}

int ListShards::search_host(const ShardSearch &fn, const float *queries, int64_t nq, const cm_search_params *p, int64_t out_stride,
                            uint32_t *out_ids, float *out_scores, int64_t *out_counts) {
    if (metric == CM_COSINE)                         // Distance.Preprocess fails on a zero query (distance.go:269-290)
        for (int64_t q = 0; q < nq; q++) {
            float ss = 0.0f;
            for (int j = 0; j < dim; j++) ss += queries[(size_t)q * dim + j] * queries[(size_t)q * dim + j];
            if (ss == 0.0f) return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)q);
        }
    std::lock_guard<std::mutex> lk(search_mu);
    const int64_t K = effective_k(p);
    if (out_stride < K) return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)K);
    if (K == 0) { memset(out_counts, 0, (size_t)nq * 8); return CM_OK; }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dev[0]);
    cudaStream_t lead = st[0];
    int rc = CM_OK;
    if (cap_ql < nq * dim) {
        cudaFree(q_lead); q_lead = nullptr;
        if (cudaMalloc(&q_lead, (size_t)nq * dim * 4) != cudaSuccess) rc = fail(CM_ERR_CUDA, "query buffer allocation failed");
        else cap_ql = nq * dim;
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(q_lead, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, lead);
        rc = search_impl(fn, q_lead, nq, p, K, lead);
    }
    if (rc == CM_OK) {
        cudaSetDevice(dev[0]);
        cudaMemcpy2DAsync(out_ids, (size_t)out_stride * 4, m_ids, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToHost, lead);
        cudaMemcpy2DAsync(out_scores, (size_t)out_stride * 4, m_sc, (size_t)K * 4, (size_t)K * 4, (size_t)nq, cudaMemcpyDeviceToHost, lead);
        cudaMemcpyAsync(out_counts, m_cnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, lead);
        cudaError_t e = cudaStreamSynchronize(lead);
        if (e != cudaSuccess) rc = fail(CM_ERR_CUDA, "list-sharded search: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    return rc;
}

}  // namespace cm
