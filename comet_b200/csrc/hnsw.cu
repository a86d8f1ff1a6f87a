// hnsw.cu -- HNSWIndex device state, warp-per-query graph search and the cm_hnsw_* entry points.
//
// Replaces hnswIndexSearch.searchSingleQuery (hnsw_index_search.go:248-354) and HNSWIndex.searchLayer
// (hnsw_index.go:565-629): greedy descent through the upper layers (K10), beam search on layer 0 with
// Go's container/heap semantics (K11: Push = append + sift-up, Pop = swap(0, n-1) + sift-down, strict
// comparisons -- the order of equal keys decides which node is expanded next, so the heaps are
// reproduced operation by operation), document filter and threshold applied after the traversal.
// Results are bit-identical to a sequential CPU replay of the reference on the SAME graph.  (The reference draws node levels
// from an unseeded global RNG, hnsw_index.go:474-484, so two reference builds never agree with each
// other either; parity is defined on a shared graph.)
//
// Device layout (replaces map[uint32]*hnswNode, hnsw_index.go:50-61, 111-155):
//   rows fp32 [n][ld], ids u32 [n], deleted u8 [n], levels i32 [n]   -- by slot (insertion order)
//   node_base i64 [n+1]  : first (node, layer) pair of a slot;  edge_off i64 [pairs+1];  edges u32 (slots)
// One warp per query: the 32 lanes evaluate up to 32 neighbour distances at once -- each lane walks
// its own row in the reference's sequential order -- and lane 0 replays the heap operations in edge
// order.  The visited set is one bit per node per in-flight query (global memory, atomicOr).
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"

namespace cm {

static constexpr int HNSW_WARPS = 4;

struct HNSWIndex {
    int dim = 0, ld = 0, metric = 0, m = 0, efc = 0, efs = 0, device = 0;
    int64_t n = 0;
    float *rows = nullptr;
    uint32_t *ids = nullptr;
    uint8_t *deleted = nullptr;
    int *levels = nullptr;
    long long *node_base = nullptr, *edge_off = nullptr;
    uint32_t *edges = nullptr;
    long long entry_slot = -1;
    int max_level = -1;
    std::unordered_map<uint32_t, int64_t> slot_of;
    std::unordered_set<uint32_t> deleted_ids;

    void free_dev() {
        cudaFree(rows); cudaFree(ids); cudaFree(deleted); cudaFree(levels); cudaFree(node_base); cudaFree(edge_off); cudaFree(edges);
        rows = nullptr; ids = nullptr; deleted = nullptr; levels = nullptr; node_base = nullptr; edge_off = nullptr; edges = nullptr;
    }
    ~HNSWIndex() { free_dev(); }
};

struct HCand { float d; uint32_t slot; };

// ---- Go container/heap on (distance, slot) pairs; IS_MAX: less(i, j) = d[i] > d[j], else d[i] < d[j] ----
template <bool IS_MAX>
__device__ __forceinline__ bool h_less(const HCand *a, int i, int j) {
    return IS_MAX ? (a[i].d > a[j].d) : (a[i].d < a[j].d);
}
template <bool IS_MAX>
__device__ __forceinline__ void h_push(HCand *a, int &n, HCand c) {
    a[n++] = c;
    int j = n - 1;
    for (;;) {                                   // heap.up
        int i = (j - 1) / 2;
        if (i == j || !h_less<IS_MAX>(a, j, i)) break;
        HCand t = a[i]; a[i] = a[j]; a[j] = t;
        j = i;
    }
}
template <bool IS_MAX>
__device__ __forceinline__ HCand h_pop(HCand *a, int &n) {
    int last = n - 1;
    HCand t = a[0]; a[0] = a[last]; a[last] = t;
    int i = 0;
    for (;;) {                                   // heap.down(0, last)
        int j1 = 2 * i + 1;
        if (j1 >= last || j1 < 0) break;
        int j = j1, j2 = j1 + 1;
        if (j2 < last && h_less<IS_MAX>(a, j2, j1)) j = j2;
        if (!h_less<IS_MAX>(a, j, i)) break;
        HCand u = a[i]; a[i] = a[j]; a[j] = u;
        i = j;
    }
    n = last;
    return a[last];
}

// Distance.Calculate(query, row) by ONE lane in the reference's order (distance.go loops)
template <int METRIC, bool FMA>
__device__ __forceinline__ float row_distance(const float *__restrict__ row, const float *__restrict__ q_s, int ld) {
    const float4 *x = reinterpret_cast<const float4 *>(row);
    const float4 *q = reinterpret_cast<const float4 *>(q_s);
    float acc = 0.0f;
#pragma unroll 4
    for (int j = 0; j < ld / 4; j++) {
        float4 xv = __ldg(x + j), qv = q[j];
        acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
        acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
        acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
        acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
    }
    return metric_finish<METRIC>(acc);
}

template <int METRIC, bool FMA>
__global__ void __launch_bounds__(HNSW_WARPS * 32) hnsw_search_kernel(
    const float *__restrict__ rows, int ld, const uint32_t *__restrict__ ids, const uint8_t *__restrict__ deleted,
    const int *__restrict__ levels, const long long *__restrict__ node_base, const long long *__restrict__ edge_off,
    const uint32_t *__restrict__ edges, long long entry_slot, int max_level, const float *__restrict__ queries, int nq, int ef,
    long long k_req, float threshold, const uint8_t *__restrict__ doc_skip, uint32_t *__restrict__ visited, long long vis_words,
    HCand *__restrict__ cand_heaps, int cand_cap, long long out_stride, uint32_t *__restrict__ out_ids,
    float *__restrict__ out_scores, long long *__restrict__ out_pos, long long *__restrict__ out_counts,
    long long *__restrict__ work /* [nq][2]: distance evaluations, expansions */) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * HNSW_WARPS + warp;
    if (q >= nq) return;
    const size_t per_warp = (size_t)ld * 4 + (size_t)(ef + 1) * sizeof(HCand) + 32 * 12;
    uint8_t *base = smem + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    float *q_s = reinterpret_cast<float *>(base);
    HCand *res = reinterpret_cast<HCand *>(q_s + ld);
    float *nb_d = reinterpret_cast<float *>(res + (ef + 1));
    uint32_t *nb_slot = reinterpret_cast<uint32_t *>(nb_d + 32);
    uint32_t *nb_new = nb_slot + 32;
    for (int j = lane; j < ld; j += 32) q_s[j] = queries[(size_t)q * ld + j];
    __syncwarp();
    HCand *cands = cand_heaps + (size_t)q * cand_cap;
    uint32_t *vis = visited + (size_t)q * vis_words;
    long long evals = 0, expansions = 0;

    // ---- phase 1: greedy descent, hnsw_index_search.go:271-296 ----
    long long curr = entry_slot;
    float curr_dist = 0.0f;
    if (lane == 0) curr_dist = row_distance<METRIC, FMA>(rows + (size_t)curr * ld, q_s, ld);
    curr_dist = __shfl_sync(0xffffffffu, curr_dist, 0);
    evals++;
    for (int lc = max_level; lc > 0; lc--) {
        bool changed = true;
        while (changed) {
            changed = false;
            if (lc > levels[curr]) break;                       // lc < len(node.Edges)
            const long long pair = node_base[curr] + lc;
            const long long e0 = edge_off[pair], deg = edge_off[pair + 1] - e0;
            const long long node = curr;
            (void)node;
            for (long long b = 0; b < deg; b += 32) {
                long long j = b + lane;
                uint32_t nb = 0;
                bool valid = false;
                float d = 0.0f;
                if (j < deg) {
                    nb = edges[e0 + j];
                    valid = deleted[nb] == 0;
                    if (valid) d = row_distance<METRIC, FMA>(rows + (size_t)nb * ld, q_s, ld);
                }
                int cnt = (int)min(32LL, deg - b);
                for (int t = 0; t < cnt; t++) {
                    float dt = __shfl_sync(0xffffffffu, d, t);
                    bool vt = __shfl_sync(0xffffffffu, (int)valid, t) != 0;
                    uint32_t nt = __shfl_sync(0xffffffffu, nb, t);
                    if (vt) {
                        evals++;
                        if (dt < curr_dist) { curr_dist = dt; curr = nt; changed = true; }
                    }
                }
            }
        }
    }

    // ---- phase 2: searchLayer(query, curr, ef, 0), hnsw_index.go:565-629 ----
    int n_c = 0, n_r = 0;
    bool overflow = false;
    if (deleted[curr] == 0) {
        float d = 0.0f;
        if (lane == 0) {
            d = row_distance<METRIC, FMA>(rows + (size_t)curr * ld, q_s, ld);
            HCand c{d, (uint32_t)curr};
            h_push<false>(cands, n_c, c);
            h_push<true>(res, n_r, c);
        }
        evals++;
    }
    if (lane == 0) atomicOr(&vis[curr >> 5], 1u << (curr & 31));
    __syncwarp();
    for (;;) {
        // pop the nearest candidate (lane 0) and broadcast the decision
        float cur_d = 0.0f;
        uint32_t cur_slot = 0;
        int go = 0;
        if (lane == 0) {
            if (n_c > 0) {
                HCand cur = h_pop<false>(cands, n_c);
                if (!(n_r >= ef && cur.d > res[0].d)) { go = 1; cur_d = cur.d; cur_slot = cur.slot; }
            }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) break;
        cur_slot = __shfl_sync(0xffffffffu, cur_slot, 0);
        (void)cur_d;
        expansions++;
        const long long pair = node_base[cur_slot];             // layer 0 always exists
        const long long e0 = edge_off[pair], deg = edge_off[pair + 1] - e0;
        for (long long b = 0; b < deg; b += 32) {
            long long j = b + lane;
            uint32_t nb = 0, isnew = 0;
            float d = 0.0f;
            if (j < deg) {
                nb = edges[e0 + j];
                if (deleted[nb] == 0) {
                    uint32_t bit = 1u << (nb & 31);
                    uint32_t old = atomicOr(&vis[nb >> 5], bit);
                    isnew = (old & bit) == 0u;
                }
                if (isnew) d = row_distance<METRIC, FMA>(rows + (size_t)nb * ld, q_s, ld);
            }
            nb_d[lane] = d; nb_slot[lane] = nb; nb_new[lane] = isnew;
            evals += __popc(__ballot_sync(0xffffffffu, isnew != 0u));
            __syncwarp();
            if (lane == 0) {
                int cnt = (int)min(32LL, deg - b);
                for (int t = 0; t < cnt; t++) {
                    if (!nb_new[t]) continue;
                    float dt = nb_d[t];
                    if (n_r < ef || dt < res[0].d) {
                        HCand c{dt, nb_slot[t]};
                        if (n_c >= cand_cap) { overflow = true; break; }
                        h_push<false>(cands, n_c, c);
                        h_push<true>(res, n_r, c);
                        if (n_r > ef) (void)h_pop<true>(res, n_r);
                    }
                }
            }
            __syncwarp();
        }
        if (__shfl_sync(0xffffffffu, (int)overflow, 0)) break;
    }

    // ---- results: pop the max-heap back to front (ascending), post-filter, first k ----
    if (lane == 0) {
        long long count = 0;
        if (overflow) {
            count = -1;
        } else {
            int n = n_r;
            HCand *sorted = cands;                              // the candidate heap is dead: reuse as scratch
            for (int i = n - 1; i >= 0; i--) sorted[i] = h_pop<true>(res, n_r);
            long long kept = 0;
            for (int i = 0; i < n; i++) {
                uint32_t s = sorted[i].slot;
                if (doc_skip != nullptr && doc_skip[s]) continue;                    // docFilter.ShouldSkip
                if (threshold > 0.0f && sorted[i].d > threshold) continue;
                sorted[kept++] = sorted[i];
            }
            long long k = (k_req <= 0 || k_req > kept) ? kept : k_req;               // sanitizeK(k, len(results))
            if (k > out_stride) k = out_stride;
            for (long long i = 0; i < k; i++) {
                size_t o = (size_t)q * out_stride + i;
                out_ids[o] = ids[sorted[i].slot];
                out_scores[o] = sorted[i].d;
                if (out_pos) out_pos[o] = sorted[i].slot;
            }
            count = k;
        }
        out_counts[q] = count;
        if (work) { work[(size_t)q * 2] = evals; work[(size_t)q * 2 + 1] = expansions; }
    }
}

static int hnsw_search_device(HNSWIndex &ix, const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                              uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, int64_t *work,
                              cudaStream_t st, bool check_zero) {
    if (nq <= 0) return CM_OK;
    if (ix.n == 0 || ix.max_level == -1) {                                           // hnsw_index_search.go:258-260
        CM_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)nq * sizeof(int64_t), st));
        return CM_OK;
    }
    int ef = p->ef_search > 0 ? p->ef_search : ix.efs;                               // :302-305
    int64_t k_bound = (p->k <= 0 || p->k > ef) ? ef : p->k;
    if (out_stride < std::min<int64_t>(k_bound, ix.n))
        return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < min(k, ef, n) = %lld", (long long)out_stride,
                    (long long)std::min<int64_t>(k_bound, ix.n));
    bool fma = rounding_mode() == CM_ROUND_FMA;
    const int ld = ix.ld;
    // Distance.Preprocess on the queries
    float *qp = nullptr;
    int *qflags = nullptr;
    CM_TRY(ws_alloc((void **)&qp, (size_t)nq * ld * 4, st));
    CM_TRY(ws_alloc((void **)&qflags, (size_t)nq * sizeof(int), st));
    CM_TRY(launch_preprocess_rows(ix.metric, fma, q_dev, nq, ix.dim, ix.dim, qp, ld, qflags, st));
    if (check_zero && ix.metric == CM_COSINE) {
        std::vector<int> hf((size_t)nq);
        CM_CUDA(cudaMemcpyAsync(hf.data(), qflags, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < nq; i++)
            if (hf[(size_t)i]) {
                ws_free(qp, st); ws_free(qflags, st);
                return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)i);
            }
    }
    // document filter -> per-slot mask (soft deletes are handled inside the traversal, like the reference)
    uint8_t *doc_skip = nullptr;
    uint32_t *filt_dev = nullptr;
    if (p->filter_ids && p->nfilter > 0) {
        std::vector<uint32_t> f(p->filter_ids, p->filter_ids + p->nfilter);
        std::sort(f.begin(), f.end());
        f.erase(std::unique(f.begin(), f.end()), f.end());
        CM_TRY(ws_alloc((void **)&filt_dev, f.size() * 4, st));
        CM_TRY(ws_alloc((void **)&doc_skip, (size_t)ix.n, st));
        CM_CUDA(cudaMemcpyAsync(filt_dev, f.data(), f.size() * 4, cudaMemcpyHostToDevice, st));
        CM_TRY(launch_build_skip(ix.ids, nullptr, ix.n, filt_dev, (int64_t)f.size(), doc_skip, st));
        CM_CUDA(cudaStreamSynchronize(st));
    }
    const long long vis_words = (ix.n + 31) / 32;
    const int cand_cap = (int)std::min<int64_t>(ix.n + 1, (int64_t)16 * ef + 4096);
    size_t per_warp = ((size_t)ld * 4 + (size_t)(ef + 1) * sizeof(HCand) + 32 * 12 + 15) & ~(size_t)15;
    size_t smem = per_warp * HNSW_WARPS;
    if (smem > max_smem_optin()) return fail(CM_ERR_UNSUPPORTED, "efSearch %d with dim %d does not fit shared memory", ef, ix.dim);
    // queries in groups so that the visited bitmaps stay bounded (<= 1 GiB)
    int64_t qgroup = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 30) / (vis_words * 4 + (int64_t)cand_cap * 8)));
    uint32_t *visited = nullptr;
    HCand *heaps = nullptr;
    CM_TRY(ws_alloc((void **)&visited, (size_t)qgroup * vis_words * 4, st));
    CM_TRY(ws_alloc((void **)&heaps, (size_t)qgroup * cand_cap * sizeof(HCand), st));
    for (int64_t q0 = 0; q0 < nq; q0 += qgroup) {
        int64_t m = std::min(qgroup, nq - q0);
        CM_CUDA(cudaMemsetAsync(visited, 0, (size_t)m * vis_words * 4, st));
        unsigned blocks = (unsigned)((m + HNSW_WARPS - 1) / HNSW_WARPS);
        ProfScope prof(CM_PROF_HNSW, st);
#define CM_HNSW_LAUNCH(M, F)                                                                                          \
    do {                                                                                                              \
        CM_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hnsw_search_kernel<M, F><<<blocks, HNSW_WARPS * 32, smem, st>>>(                                              \
            ix.rows, ld, ix.ids, ix.deleted, ix.levels, ix.node_base, ix.edge_off, ix.edges, ix.entry_slot, ix.max_level, \
            qp + (size_t)q0 * ld, (int)m, ef, (long long)p->k, p->threshold, doc_skip, visited, vis_words, heaps, cand_cap, \
            (long long)out_stride, out_ids + (size_t)q0 * out_stride, out_scores + (size_t)q0 * out_stride,          \
            out_pos ? (long long *)out_pos + (size_t)q0 * out_stride : nullptr, (long long *)out_counts + q0,         \
            work ? (long long *)work + (size_t)q0 * 2 : nullptr);                                                     \
    } while (0)
        switch (ix.metric) {
        case CM_L2: if (fma) CM_HNSW_LAUNCH(CM_L2, true); else CM_HNSW_LAUNCH(CM_L2, false); break;
        case CM_L2SQ: if (fma) CM_HNSW_LAUNCH(CM_L2SQ, true); else CM_HNSW_LAUNCH(CM_L2SQ, false); break;
        default: if (fma) CM_HNSW_LAUNCH(CM_COSINE, true); else CM_HNSW_LAUNCH(CM_COSINE, false); break;
        }
#undef CM_HNSW_LAUNCH
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    ws_free(qp, st); ws_free(qflags, st); ws_free(doc_skip, st); ws_free(filt_dev, st); ws_free(visited, st); ws_free(heaps, st);
    return CM_OK;
}

}  // namespace cm

struct cm_hnsw {
    cm::HNSWIndex ix;
};

extern "C" {

int cm_hnsw_create(int dim, int metric, int m, int ef_construction, int ef_search, cm_hnsw **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");                 // hnsw_index.go:173
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    CM_TRY(cm::ensure_device());
    cm_hnsw *h = new cm_hnsw();
    // hnsw_index.go:178-190 defaults: M 16, efConstruction 200, efSearch = efConstruction
    h->ix.dim = dim; h->ix.metric = metric;
    h->ix.ld = (dim + cm::SCAN_CHUNK - 1) / cm::SCAN_CHUNK * cm::SCAN_CHUNK;
    h->ix.m = m > 0 ? m : 16;
    h->ix.efc = ef_construction > 0 ? ef_construction : 200;
    h->ix.efs = ef_search > 0 ? ef_search : h->ix.efc;
    cudaGetDevice(&h->ix.device);
    *out = h;
    return CM_OK;
}
int cm_hnsw_destroy(cm_hnsw *h) {
    delete h;
    return CM_OK;
}
int64_t cm_hnsw_size(const cm_hnsw *h) { return h ? h->ix.n : 0; }
int cm_hnsw_ef_search(const cm_hnsw *h) { return h ? h->ix.efs : 0; }

// Upload a graph built by HNSWIndex.Add / insertNode (hnsw_index.go:228-288, 493-552) -- by the Go
// package's own builder (or any other host-side builder).  rows are the STORED vectors (already preprocessed by Add);
// slots are insertion order; edge_off has sum(levels[i] + 1) + 1 entries, pairs ordered by (slot, layer);
// edge_ids are neighbour node IDs.
int cm_hnsw_load_graph(cm_hnsw *h, int64_t n, const uint32_t *ids, const float *rows, const int32_t *levels,
                       const int64_t *edge_off, const uint32_t *edge_ids, uint32_t entry_id, int max_level) {
    if (!h || n < 0 || (n > 0 && (!ids || !rows || !levels || !edge_off || !edge_ids)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::HNSWIndex &ix = h->ix;
    ix.free_dev();
    ix.slot_of.clear(); ix.deleted_ids.clear();
    ix.n = 0; ix.entry_slot = -1; ix.max_level = -1;
    if (n == 0) return CM_OK;
    std::vector<long long> base((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        if (levels[i] < 0) return cm::fail(CM_ERR_INVALID_ARG, "negative level at slot %lld", (long long)i);
        base[(size_t)i + 1] = base[(size_t)i] + levels[i] + 1;
        ix.slot_of[ids[i]] = i;
    }
    long long pairs = base[(size_t)n], n_edges = edge_off[pairs];
    std::vector<uint32_t> eslots((size_t)std::max<long long>(n_edges, 1));
    for (long long e = 0; e < n_edges; e++) {
        auto it = ix.slot_of.find(edge_ids[e]);
        if (it == ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "edge %lld points to unknown node ID %u", e, edge_ids[e]);
        eslots[(size_t)e] = (uint32_t)it->second;
    }
    auto ent = ix.slot_of.find(entry_id);
    if (ent == ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "entry point ID %u not in the graph", entry_id);
    int ld = ix.ld;
    CM_CUDA(cudaMalloc(&ix.rows, (size_t)n * ld * 4));
    CM_CUDA(cudaMemset(ix.rows, 0, (size_t)n * ld * 4));
    CM_CUDA(cudaMemcpy2D(ix.rows, (size_t)ld * 4, rows, (size_t)ix.dim * 4, (size_t)ix.dim * 4, (size_t)n, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&ix.ids, (size_t)n * 4));
    CM_CUDA(cudaMemcpy(ix.ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&ix.deleted, (size_t)n));
    CM_CUDA(cudaMemset(ix.deleted, 0, (size_t)n));
    CM_CUDA(cudaMalloc(&ix.levels, (size_t)n * 4));
    CM_CUDA(cudaMemcpy(ix.levels, levels, (size_t)n * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&ix.node_base, (size_t)(n + 1) * 8));
    CM_CUDA(cudaMemcpy(ix.node_base, base.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&ix.edge_off, (size_t)(pairs + 1) * 8));
    CM_CUDA(cudaMemcpy(ix.edge_off, edge_off, (size_t)(pairs + 1) * 8, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&ix.edges, eslots.size() * 4));
    CM_CUDA(cudaMemcpy(ix.edges, eslots.data(), eslots.size() * 4, cudaMemcpyHostToDevice));
    ix.n = n;
    ix.entry_slot = ent->second;
    ix.max_level = max_level;
    return CM_OK;
}

int cm_hnsw_remove(cm_hnsw *h, uint32_t id) {     // hnsw_index.go:300-330 soft delete
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    auto it = h->ix.slot_of.find(id);
    if (it == h->ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "node %u not found", id);
    if (h->ix.deleted_ids.count(id)) return cm::fail(CM_ERR_NOT_FOUND, "node %u already deleted", id);
    h->ix.deleted_ids.insert(id);
    uint8_t one = 1;
    CM_CUDA(cudaMemcpy(h->ix.deleted + it->second, &one, 1, cudaMemcpyHostToDevice));
    return CM_OK;
}

int cm_hnsw_search_device(cm_hnsw *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                          int64_t *out_counts_dev, int64_t *work_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::hnsw_search_device(h->ix, queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                                  out_counts_dev, work_dev, (cudaStream_t)stream, false);
}

int cm_hnsw_search(cm_hnsw *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                   uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, int64_t *out_work) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    if (nq <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *dq = nullptr, *dsc = nullptr;
    uint32_t *dids = nullptr;
    int64_t *dpos = nullptr, *dcnt = nullptr, *dwork = nullptr;
    size_t no = (size_t)nq * (size_t)(out_stride > 0 ? out_stride : 1);
    int rc = cm::ws_alloc((void **)&dq, (size_t)nq * dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dids, no * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dsc, no * 4, st);
    if (rc == CM_OK && out_pos) rc = cm::ws_alloc((void **)&dpos, no * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dcnt, (size_t)nq * 8, st);
    if (rc == CM_OK && out_work) rc = cm::ws_alloc((void **)&dwork, (size_t)nq * 16, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dq, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st);
        rc = cm::hnsw_search_device(h->ix, dq, nq, p, out_stride, dids, dsc, dpos, dcnt, dwork, st, true);
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(out_ids, dids, no * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_scores, dsc, no * 4, cudaMemcpyDeviceToHost, st);
        if (out_pos) cudaMemcpyAsync(out_pos, dpos, no * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, dcnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, st);
        if (out_work) {
            if (h->ix.n == 0) memset(out_work, 0, (size_t)nq * 16);
            else cudaMemcpyAsync(out_work, dwork, (size_t)nq * 16, cudaMemcpyDeviceToHost, st);
        }
    }
    cm::ws_free(dq, st); cm::ws_free(dids, st); cm::ws_free(dsc, st); cm::ws_free(dpos, st); cm::ws_free(dcnt, st); cm::ws_free(dwork, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "hnsw_search: %s", cudaGetErrorString(e));
    if (rc == CM_OK)
        for (int64_t q = 0; q < nq; q++)
            if (out_counts[q] < 0) return cm::fail(CM_ERR_UNSUPPORTED, "candidate heap overflow on query %lld (efSearch too small a bound)", (long long)q);
    return rc;
}

}  // extern "C"
