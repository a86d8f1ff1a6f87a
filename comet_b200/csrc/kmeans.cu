// kmeans.cu -- the reference's deterministic k-means (clustering.go:119-239) on the device, bit for bit:
//   init       centroid c = vector[c * (n / k)]                                   (no RNG, :147-162)
//   assign     argmin over centroids of Distance.Calculate, first minimum wins     (:180-198)
//   update     per cluster, per dimension: sequential float32 sum of the members IN VECTOR ORDER,
//              divided by float32(count); empty clusters keep their centroid       (:213-239)
//   stop       when no assignment changed, or after maxIter = 20 iterations.
// Exactness comes from keeping the summation order: members are grouped by a STABLE radix sort of
// (cluster, vector index), one thread owns one (cluster, dimension) pair and walks its members in order.
// Full-dimensional assignment reuses the exact flat scan (k = 1 against the centroid table); the
// small sub-space problems of PQ training use a shared-memory kernel.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "kmeans.cuh"

namespace cm {

// ---- sub-space assignment: one thread per vector, centroids staged in shared memory tiles ----
template <bool FMA>
__global__ void kmeans_assign_sub_kernel(const float *__restrict__ x, long long n, long long ldx, int off, int d,
                                         const float *__restrict__ cent, int k, int tile, int *__restrict__ assign,
                                         int *__restrict__ changed) {
    extern __shared__ float cen_s[];
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const float *v = x + (size_t)(i < n ? i : 0) * ldx + off;
    float best = INFINITY;
    int best_c = 0;
    for (int c0 = 0; c0 < k; c0 += tile) {
        int m = min(tile, k - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < m * d; e += blockDim.x) cen_s[e] = cent[(size_t)c0 * d + e];
        __syncthreads();
        if (i < n) {
            for (int c = 0; c < m; c++) {
                float dist = 0.0f;
                for (int j = 0; j < d; j++) dist = l2_step<FMA>(dist, v[j], cen_s[c * d + j]);
                if (dist < best) { best = dist; best_c = c0 + c; }
            }
        }
    }
    if (i < n) {
        if (assign[i] != best_c) { assign[i] = best_c; *changed = 1; }
    }
}

__global__ void kmeans_compare_kernel(const long long *__restrict__ pos, long long n, int *__restrict__ assign,
                                      int *__restrict__ changed) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = (int)pos[i];
    if (assign[i] != c) { assign[i] = c; *changed = 1; }
}

__global__ void iota_kernel(int *__restrict__ v, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}

// seg[c] = first position in the sorted cluster keys holding a key >= c  (c = 0..k)
__global__ void kmeans_segments_kernel(const int *__restrict__ sorted_keys, long long n, int k, long long *__restrict__ seg) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > k) return;
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    seg[c] = lo;
}

// block per cluster, thread per dimension: sequential sums in vector order, then / float32(count)
__global__ void kmeans_update_kernel(const float *__restrict__ x, long long ldx, int off, int d, const int *__restrict__ members,
                                     const long long *__restrict__ seg, float *__restrict__ cent, long long ldc) {
    const int c = blockIdx.x;
    const long long s0 = seg[c], s1 = seg[c + 1];
    if (s1 == s0) return;                                  // empty cluster keeps its centroid
    const float cnt = (float)(s1 - s0);
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float sum = 0.0f;
        for (long long s = s0; s < s1; s++) sum = __fadd_rn(sum, x[(size_t)members[s] * ldx + off + j]);
        cent[(size_t)c * ldc + j] = __fdiv_rn(sum, cnt);
    }
}

__global__ void kmeans_init_kernel(const float *__restrict__ x, long long n, long long ldx, int off, int d, int k,
                                   long long step, float *__restrict__ cent, long long ldc) {
    const int c = blockIdx.x;
    long long vi = (long long)c * step;
    if (vi >= n) vi = n - 1;
    for (int j = threadIdx.x; j < ldc; j += blockDim.x) cent[(size_t)c * ldc + j] = j < d ? x[(size_t)vi * ldx + off + j] : 0.0f;
}

struct KMeansWork {
    int *assign = nullptr, *changed = nullptr, *keys_sorted = nullptr, *idx = nullptr, *members = nullptr;
    long long *seg = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    cudaStream_t st;
    int alloc(long long n, int k, cudaStream_t s) {
        st = s;
        CM_TRY(ws_alloc((void **)&assign, (size_t)n * 4, st));
        CM_TRY(ws_alloc((void **)&changed, 4, st));
        CM_TRY(ws_alloc((void **)&keys_sorted, (size_t)n * 4, st));
        CM_TRY(ws_alloc((void **)&idx, (size_t)n * 4, st));
        CM_TRY(ws_alloc((void **)&members, (size_t)n * 4, st));
        CM_TRY(ws_alloc((void **)&seg, (size_t)(k + 1) * 8, st));
        cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, assign, keys_sorted, idx, members, (int)n, 0, 32, st);
        CM_TRY(ws_alloc(&cub_tmp, cub_bytes, st));
        CM_CUDA(cudaMemsetAsync(assign, 0xff, (size_t)n * 4, st));          // UnassignedCluster = -1
        iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx, n);
        count_launch();
        return CM_OK;
    }
    void release() {
        ws_free(assign, st); ws_free(changed, st); ws_free(keys_sorted, st); ws_free(idx, st); ws_free(members, st);
        ws_free(seg, st); ws_free(cub_tmp, st);
    }
    // group vector indices by cluster, vector order inside a cluster (stable LSD radix sort)
    int group(long long n, int k) {
        int bits = 1;
        while ((1 << bits) < k + 1 && bits < 31) bits++;
        CM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, assign, keys_sorted, idx, members, (int)n, 0, bits, st));
        kmeans_segments_kernel<<<(unsigned)((k + 1 + 127) / 128), 128, 0, st>>>(keys_sorted, n, k, seg);
        count_launch(4);
        CM_CUDA(cudaGetLastError());
        return CM_OK;
    }
    int read_changed(int *out) {
        CM_CUDA(cudaMemcpyAsync(out, changed, 4, cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        return CM_OK;
    }
};

int kmeans_subspace(const float *x, int64_t n, int64_t ldx, int off, int d, int k, int max_iter, float *cent_out, cudaStream_t st) {
    if (n <= 0 || k <= 0) return fail(CM_ERR_INVALID_ARG, "k-means needs vectors and k > 0");
    if (k > n) k = (int)n;
    if (max_iter <= 0) max_iter = 20;
    bool fma = rounding_mode() == CM_ROUND_FMA;
    KMeansWork W;
    CM_TRY(W.alloc(n, k, st));
    long long step = n / k;
    if (step == 0) step = 1;
    kmeans_init_kernel<<<k, 32, 0, st>>>(x, n, ldx, off, d, k, step, cent_out, d);
    count_launch();
    int tile = std::max(1, std::min(k, (int)(48 * 1024 / (d * 4))));
    size_t smem = (size_t)tile * d * 4;
    for (int it = 0; it < max_iter; it++) {
        CM_CUDA(cudaMemsetAsync(W.changed, 0, 4, st));
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (fma) kmeans_assign_sub_kernel<true><<<blocks, 128, smem, st>>>(x, n, ldx, off, d, cent_out, k, tile, W.assign, W.changed);
        else kmeans_assign_sub_kernel<false><<<blocks, 128, smem, st>>>(x, n, ldx, off, d, cent_out, k, tile, W.assign, W.changed);
        count_launch();
        CM_CUDA(cudaGetLastError());
        int changed = 0;
        CM_TRY(W.read_changed(&changed));
        if (!changed) break;
        CM_TRY(W.group(n, k));
        kmeans_update_kernel<<<k, 64, 0, st>>>(x, ldx, off, d, W.members, W.seg, cent_out, d);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    W.release();
    return CM_OK;
}

// exact nearest-centroid assignment of n device rows ([n_pad][ld], n_pad multiple of 8) against `cent`
int assign_nearest(FlatIndex &cent, const float *rows, int64_t n, long long *pos_out, cudaStream_t st) {
    WsScope ws(st);
    const int64_t group = 32768;
    uint32_t *t_ids = nullptr;
    float *t_sc = nullptr;
    long long *t_cnt = nullptr;
    CM_TRY(ws.get(&t_ids, (size_t)std::min(group, n) * 4));
    CM_TRY(ws.get(&t_sc, (size_t)std::min(group, n) * 4));
    CM_TRY(ws.get(&t_cnt, (size_t)std::min(group, n) * 8));
    for (int64_t i0 = 0; i0 < n; i0 += group) {
        int64_t m = std::min(group, n - i0);
        int64_t mpad = (m + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
        cm_flat_stats cst{};
        CM_TRY(cent.search_exact(rows + (size_t)i0 * cent.ld, m, mpad, 1, nullptr, 0.0f, 1, t_ids, t_sc, (int64_t *)(pos_out + i0),
                                 (int64_t *)t_cnt, st, &cst));
    }
    return CM_OK;
}

int kmeans_full(FlatIndex &cent, const float *rows, int64_t n, int k, int max_iter, long long *final_assign, cudaStream_t st) {
    WsScope ws(st);
    if (n <= 0 || k <= 0) return fail(CM_ERR_INVALID_ARG, "k-means needs vectors and k > 0");
    if (k > n) k = (int)n;
    if (max_iter <= 0) max_iter = 20;
    const int d = cent.dim, ld = cent.ld;
    KMeansWork W;
    CM_TRY(W.alloc(n, k, st));
    long long *pos = nullptr;
    CM_TRY(ws.get(&pos, (size_t)n * 8));
    // centroid table: k raw rows in the FlatIndex (its TMA descriptor stays valid while rows are rewritten in place)
    cent.n = 0; cent.ids_host_mirror.clear();
    CM_TRY(cent.reserve(k));
    long long step = n / k;
    if (step == 0) step = 1;
    kmeans_init_kernel<<<k, 128, 0, st>>>(rows, n, ld, 0, d, k, step, cent.rows, ld);
    count_launch();
    {
        std::vector<uint32_t> ids((size_t)k);
        for (int i = 0; i < k; i++) ids[(size_t)i] = (uint32_t)i;
        CM_CUDA(cudaMemcpyAsync(cent.ids, ids.data(), (size_t)k * 4, cudaMemcpyHostToDevice, st));
        CM_CUDA(cudaMemsetAsync(cent.deleted, 0, (size_t)k, st));
        CM_CUDA(cudaStreamSynchronize(st));
        cent.ids_host_mirror = ids;
        cent.n = k;
    }
    for (int it = 0; it < max_iter; it++) {
        CM_CUDA(cudaMemsetAsync(W.changed, 0, 4, st));
        CM_TRY(assign_nearest(cent, rows, n, pos, st));
        kmeans_compare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, n, W.assign, W.changed);
        count_launch();
        int changed = 0;
        CM_TRY(W.read_changed(&changed));
        if (!changed) break;
        CM_TRY(W.group(n, k));
        kmeans_update_kernel<<<k, 128, 0, st>>>(rows, ld, 0, d, W.members, W.seg, cent.rows, ld);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    if (final_assign) CM_TRY(assign_nearest(cent, rows, n, final_assign, st));   // FindNearestCentroidIndex vs the FINAL centroids
    W.release();
    return CM_OK;
}

// residual[i] = row[i] - centroid[assign[i]]  (ivfpq_index.go:222-231), padding columns zero
__global__ void residual_kernel(const float *__restrict__ rows, long long n, int d, int ld, const float *__restrict__ cent,
                                const long long *__restrict__ assign, float *__restrict__ out) {
    long long i = blockIdx.x;
    if (i >= n) return;
    const float *c = cent + (size_t)assign[i] * ld;
    for (int j = threadIdx.x; j < ld; j += blockDim.x)
        out[(size_t)i * ld + j] = j < d ? __fsub_rn(rows[(size_t)i * ld + j], c[j]) : 0.0f;
}
int launch_residuals(const float *rows, int64_t n, int d, int ld, const float *cent, const long long *assign, float *out, cudaStream_t st) {
    if (n <= 0) return CM_OK;
    residual_kernel<<<(unsigned)n, 128, 0, st>>>(rows, n, d, ld, cent, assign, out);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// host rows [n][dim] -> device [n_pad][ld] zero padded (n_pad multiple of 8), no preprocessing (Train never preprocesses)
int upload_training_rows(const float *rows_host, int64_t n, int dim, int ld, float **out, cudaStream_t st) {
    int64_t npad = (n + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
    float *d = nullptr;
    CM_TRY(ws_alloc((void **)&d, (size_t)npad * ld * 4, st));       // ownership goes to the caller (*out)
    CM_CUDA(cudaMemsetAsync(d, 0, (size_t)npad * ld * 4, st));
    CM_CUDA(cudaMemcpy2DAsync(d, (size_t)ld * 4, rows_host, (size_t)dim * 4, (size_t)dim * 4, (size_t)n, cudaMemcpyHostToDevice, st));
    *out = d;
    return CM_OK;
}

}  // namespace cm
