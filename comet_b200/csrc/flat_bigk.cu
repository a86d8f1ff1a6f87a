// flat_bigk.cu -- flat search when the caller wants (nearly) everything back: WithK(0) or
// k > 4096 (limiter.go:12-17 turns k <= 0 into "all").  The reference sorts all N results anyway
// (flat_index_search.go:277); here every row's key goes to HBM and one radix sort per query orders
// them.  Not a hot path: a result of >4096 rows per query is dominated by returning it.
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "flat_index.cuh"
#include "flat_kernels.cuh"

namespace cm {

template <int METRIC, bool FMA>
__global__ void all_keys_kernel(const float *__restrict__ rows, int ld, long long n, const float *__restrict__ q,
                                const uint8_t *__restrict__ skip, float threshold, uint64_t *__restrict__ keys) {
    extern __shared__ float qs[];
    for (int j = threadIdx.x; j < ld; j += blockDim.x) qs[j] = q[j];
    __syncthreads();
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *x = reinterpret_cast<const float4 *>(rows + (size_t)i * ld);
    float acc = 0.0f;
    for (int j = 0; j < ld / 4; j++) {
        float4 v = x[j];
        acc = metric_step<METRIC, FMA>(acc, qs[4 * j + 0], v.x);
        acc = metric_step<METRIC, FMA>(acc, qs[4 * j + 1], v.y);
        acc = metric_step<METRIC, FMA>(acc, qs[4 * j + 2], v.z);
        acc = metric_step<METRIC, FMA>(acc, qs[4 * j + 3], v.w);
    }
    float dist = metric_finish<METRIC>(acc);
    bool pass = !(skip && skip[i]) && !(threshold > 0.0f && dist > threshold);
    keys[i] = pass ? make_key(dist, (uint32_t)i) : KEY_INF;
}

__global__ void emit_sorted_kernel(const uint64_t *__restrict__ keys, long long n, long long k,
                                   const uint32_t *__restrict__ row_ids, uint32_t *__restrict__ out_ids,
                                   float *__restrict__ out_scores, long long *__restrict__ out_pos,
                                   long long *__restrict__ out_count) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i == 0) {
        // number of valid keys = first index holding KEY_INF (binary search), capped at k
        long long lo = 0, hi = n;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (keys[mid] == KEY_INF) hi = mid; else lo = mid + 1;
        }
        *out_count = lo < k ? lo : k;
    }
    if (i >= k || i >= n) return;
    uint64_t key = keys[i];
    if (key == KEY_INF) return;
    uint32_t pos = key_pos(key);
    out_ids[i] = row_ids ? row_ids[pos] : pos;
    out_scores[i] = key_score(key);
    if (out_pos) out_pos[i] = pos;
}

template <int METRIC>
static void launch_all_keys(bool fma, const float *rows, int ld, int64_t n, const float *q, const uint8_t *skip,
                            float threshold, uint64_t *keys, cudaStream_t st) {
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (fma) all_keys_kernel<METRIC, true><<<blocks, 128, ld * 4, st>>>(rows, ld, n, q, skip, threshold, keys);
    else all_keys_kernel<METRIC, false><<<blocks, 128, ld * 4, st>>>(rows, ld, n, q, skip, threshold, keys);
    count_launch();
}

int FlatIndex::search_exact_bigk(const float *qp, int64_t nq, int64_t k_eff, const uint8_t *skip, float threshold,
                                 int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                                 int64_t *out_counts, cudaStream_t st, cm_flat_stats *stats) {
    WsScope ws(st);
    bool fma = rounding_mode() == CM_ROUND_FMA;
    uint64_t *keys = nullptr, *sorted = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, sorted, (int64_t)n, 0, 64, st);
    CM_TRY(ws.get(&keys, (size_t)n * 8));
    CM_TRY(ws.get(&sorted, (size_t)n * 8));
    CM_TRY(ws.get(&tmp, tmp_bytes));
    for (int64_t qi = 0; qi < nq; qi++) {
        const float *q = qp + (size_t)qi * ld;
        switch (metric) {
        case CM_L2: launch_all_keys<CM_L2>(fma, rows, ld, n, q, skip, threshold, keys, st); break;
        case CM_L2SQ: launch_all_keys<CM_L2SQ>(fma, rows, ld, n, q, skip, threshold, keys, st); break;
        default: launch_all_keys<CM_COSINE>(fma, rows, ld, n, q, skip, threshold, keys, st); break;
        }
        CM_CUDA(cudaGetLastError());
        CM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, sorted, (int64_t)n, 0, 64, st));
        count_launch(3);
        unsigned blocks = (unsigned)((k_eff + 255) / 256);
        emit_sorted_kernel<<<blocks, 256, 0, st>>>(sorted, n, k_eff, ids, out_ids + (size_t)qi * out_stride,
                                                  out_scores + (size_t)qi * out_stride,
                                                  out_pos ? (long long *)out_pos + (size_t)qi * out_stride : nullptr,
                                                  (long long *)out_counts + qi);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    stats->path_used = CM_PATH_EXACT;
    stats->passes = (int)nq;
    return CM_OK;
}

// ---- merge of per-part candidate lists when K + Kp does not fit the shared-memory merge (k <= 0 / huge k on IVF, PQ,
// IVFPQ: limiter.go:12-17 turns k <= 0 into "every candidate") ----------------------------------------------------------
__global__ void part_keys_kernel(const uint64_t *__restrict__ part_keys, const int *__restrict__ part_counts, int parts, int Kp,
                                 uint64_t *__restrict__ keys) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)parts * Kp) return;
    int part = (int)(i / Kp), j = (int)(i % Kp);
    keys[i] = j < part_counts[part] ? part_keys[i] : KEY_INF;
}

int merge_topk_bigk(const uint64_t *part_keys, const int *part_counts, int nq, int parts, int Kp, int64_t K,
                    const uint32_t *row_ids, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                    int64_t *out_counts, cudaStream_t st) {
    WsScope ws(st);
    const int64_t total = (int64_t)parts * Kp;
    uint64_t *keys = nullptr, *sorted = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, sorted, total, 0, 64, st);
    CM_TRY(ws.get(&keys, (size_t)total * 8));
    CM_TRY(ws.get(&sorted, (size_t)total * 8));
    CM_TRY(ws.get(&tmp, tmp_bytes));
    for (int q = 0; q < nq; q++) {
        part_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_keys + (size_t)q * total, part_counts + (size_t)q * parts,
                                                                         parts, Kp, keys);
        CM_CUDA(cudaGetLastError());
        CM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, sorted, total, 0, 64, st));
        emit_sorted_kernel<<<(unsigned)((std::min<int64_t>(K, total) + 255) / 256), 256, 0, st>>>(
            sorted, (long long)total, (long long)K, row_ids, out_ids + (size_t)q * out_stride, out_scores + (size_t)q * out_stride,
            out_pos ? (long long *)out_pos + (size_t)q * out_stride : nullptr, (long long *)out_counts + q);
        CM_CUDA(cudaGetLastError());
        count_launch(5);
    }
    return CM_OK;
}

// ---- shard merge when world x k does not fit the shared-memory merge (k <= 0 / huge k on a sharded index) ------
__global__ void merge_keys_kernel(const float *__restrict__ scores, const long long *__restrict__ counts, int world,
                                  long long nq, long long q, long long in_stride, uint64_t *__restrict__ keys) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= world * in_stride) return;
    long long r = i / in_stride, j = i % in_stride;
    long long c = counts ? counts[(size_t)r * nq + q] : in_stride;
    keys[i] = j < c ? make_key(scores[((size_t)r * nq + q) * in_stride + j], (uint32_t)i) : KEY_INF;
}
__global__ void merge_emit_kernel(const uint64_t *__restrict__ keys, long long total, long long k, const uint32_t *__restrict__ ids,
                                  const long long *__restrict__ counts, int world, long long nq, long long q, long long in_stride,
                                  uint32_t *__restrict__ out_ids, float *__restrict__ out_scores, long long *__restrict__ out_count) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i == 0 && out_count) {
        long long bad = 0, have = 0;
        for (int r = 0; r < world; r++) {
            long long c = counts ? counts[(size_t)r * nq + q] : in_stride;
            bad = min(bad, c);
            have += c < in_stride ? (c > 0 ? c : 0) : in_stride;
        }
        *out_count = bad < 0 ? bad : (have < k ? have : k);
    }
    if (i >= k || i >= total) return;
    uint64_t key = keys[i];
    if (key == KEY_INF) return;
    uint32_t src = key_pos(key);
    long long r = src / in_stride, j = src % in_stride;
    out_ids[i] = ids[((size_t)r * nq + q) * in_stride + j];
    out_scores[i] = key_score(key);
}

int merge_shards_bigk(const uint32_t *ids, const float *scores, const int64_t *counts, int world, int64_t nq,
                      int64_t in_stride, int64_t k, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                      int64_t *out_counts, cudaStream_t st) {
    WsScope ws(st);
    const int64_t total = (int64_t)world * in_stride;
    if (total >= (1ll << 32)) return fail(CM_ERR_UNSUPPORTED, "%d shards x k=%lld too large for the shard merge", world, (long long)in_stride);
    uint64_t *keys = nullptr, *sorted = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, sorted, total, 0, 64, st);
    CM_TRY(ws.get(&keys, (size_t)total * 8));
    CM_TRY(ws.get(&sorted, (size_t)total * 8));
    CM_TRY(ws.get(&tmp, tmp_bytes));
    for (int64_t q = 0; q < nq; q++) {
        merge_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scores, (const long long *)counts, world, (long long)nq,
                                                                          (long long)q, (long long)in_stride, keys);
        CM_CUDA(cudaGetLastError());
        CM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, sorted, total, 0, 64, st));
        merge_emit_kernel<<<(unsigned)((std::min(k, total) + 255) / 256), 256, 0, st>>>(
            sorted, (long long)total, (long long)k, ids, (const long long *)counts, world, (long long)nq, (long long)q,
            (long long)in_stride, out_ids + (size_t)q * out_stride, out_scores + (size_t)q * out_stride,
            out_counts ? (long long *)out_counts + q : nullptr);
        CM_CUDA(cudaGetLastError());
        count_launch(5);
    }
    return CM_OK;
}

}  // namespace cm
