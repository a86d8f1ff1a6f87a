// flat_index.cuh -- FlatIndex device object (see flat.cu for the layout).
#pragma once

#include <cuda_bf16.h>

#include <mutex>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace cm {

int acquire_stream(cudaStream_t *out);
void release_stream(cudaStream_t s);
int ws_alloc(void **p, size_t bytes, cudaStream_t s);
namespace wire { struct Sink; struct Source; }
void ws_free(void *p, cudaStream_t s);

// Stream-ordered workspace of one call: everything taken through it goes back to the pool when the scope ends,
// on every return path (an early CM_TRY return used to leak what had been allocated before it).
struct WsScope {
    cudaStream_t st;
    std::vector<void *> owned;
    explicit WsScope(cudaStream_t s) : st(s) {}
    WsScope(const WsScope &) = delete;
    WsScope &operator=(const WsScope &) = delete;
    ~WsScope() {
        for (void *p : owned) ws_free(p, st);
    }
    template <class T>
    int get(T **p, size_t bytes) {
        void *v = nullptr;
        CM_TRY(ws_alloc(&v, bytes, st));
        *p = static_cast<T *>(v);
        owned.push_back(v);
        return CM_OK;
    }
    void adopt(void *p) {           // a buffer a helper allocated with ws_alloc on the same stream
        if (p) owned.push_back(p);
    }
};

struct FlatIndex;
// flat_filter.cu
int upload_filter_ids(const uint32_t *ids, int64_t nf, uint32_t *dst_dev, cudaStream_t st);
int sort_u32_device(uint32_t *keys, int64_t n, WsScope &ws, cudaStream_t st);
// ivf.cu: top-k over an explicit candidate list (ascending scan positions, count on the device, at most `cap`)
int gather_scan_topk(FlatIndex &S, const float *qp, int64_t nq, const uint32_t *cand_pos, const int *cand_cnt_dev, int64_t cap,
                     float threshold, int64_t k_eff, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                     int64_t *out_counts, cudaStream_t st);

struct FlatIndex {
    int dim = 0, ld = 0, metric = 0, device = 0;
    int64_t n = 0, cap = 0;
    float *rows = nullptr;
    uint32_t *ids = nullptr;
    uint8_t *deleted = nullptr;
    CUtensorMap tmap;                       // fp32 rows, box 32 x 128, SWIZZLE_128B
    std::vector<uint32_t> ids_host_mirror;  // for Remove / lookupNodeVectors (O(N) ID scans stay on the host)
    std::unordered_set<uint32_t> deleted_ids;
    int64_t n_deleted_rows = 0;
    bool raw_rows = false;                  // true: Add stores rows as given (k-means centroids are never normalised)

    // tensor-core candidate pass state (flat_tensor.cu)
    __nv_bfloat16 *rows_bf16 = nullptr;     // [cap][ldb] bf16 shadow of rows, ldb = dim padded to 64
    float *row_h = nullptr;                 // [cap + 128] |x|^2 / 2 for L2 / L2^2, 0 for cosine (candidate key offset); +inf behind row n
    unsigned int *max_bits = nullptr;       // device [2]: bits of max_row |x| and max_row |x - bf16(x)| (error bound)
    int ldb = 0;
    int64_t shadow_rows = 0;                // rows [0, shadow_rows) of the shadow are current
    int64_t shadow_cap = 0;
    int tensor_cta_group = 2;               // 2: CTA pairs (UMMA M=256), 1: single-CTA UMMA M=128
    std::mutex shadow_mu;
    CUtensorMap tmap_bf16;                  // box 64 x 128 rows (both operands staged in shared memory)
    CUtensorMap tmap_bf16_ts;               // box 64 x 32 rows (query-resident pass, flat_gemm_ts.cu)

    // document filters (flat_filter.cu): node IDs sorted with their scan positions, rebuilt lazily after Add / Flush
    uint32_t *id_sorted = nullptr, *pos_sorted = nullptr;
    int64_t sorted_n = -1, sorted_cap = 0;
    int max_id_run = 1;                     // most rows that share one node ID (Add does not reject repeated IDs)

    std::mutex stats_mu;
    cm_flat_stats last_stats{};
    unsigned long long *rescored_dev = nullptr;   // device: candidates re-scored by the last tensor-path search (all queries)
    int *staged_dev = nullptr;              // device [8]: most keys any query staged in the select of phase p (this search)
    int *staged_host = nullptr;             // pinned copy of the previous search's values (-1: unknown): picks the select shape

    ~FlatIndex();
    int reserve(int64_t want);
    int rebuild_tmap();
    int add_from_device(const uint32_t *ids_host, const float *src_dev, int64_t n_add, float *writeback_host,
                        cudaStream_t st);
    int remove(uint32_t id);
    int flush();
    int reset();
    int load_stored_rows(const uint32_t *ids_host, const float *rows_host, int64_t n_add);
    int read_rows(int64_t first, int64_t m, float *out) const;
    int ensure_id_sort(cudaStream_t st);
    int filter_candidates(const uint32_t *filt_sorted, int64_t nf, uint32_t *cand_pos, int64_t cap, int *cand_cnt, WsScope &ws,
                          cudaStream_t st);
    int save(wire::Sink &s);       // FlatIndex.WriteTo
    int load(wire::Source &s);     // FlatIndex.ReadFrom
    int search_device(const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                      uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, cudaStream_t st,
                      bool check_zero_queries);
    int search_exact(const float *qp, int64_t nq, int64_t nq_pad, int64_t k_eff, const uint8_t *skip, float threshold,
                     int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts,
                     cudaStream_t st, cm_flat_stats *stats);
    int search_exact_bigk(const float *qp, int64_t nq, int64_t k_eff, const uint8_t *skip, float threshold,
                          int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                          int64_t *out_counts, cudaStream_t st, cm_flat_stats *stats);
    // flat_tensor.cu
    bool tensor_path_eligible(int64_t nq, int64_t k_eff, bool has_skip, float threshold) const;
    int search_tensor(const float *q_raw, float *qp, int *qflags, int64_t nq, int64_t k_eff, const uint8_t *skip,
                      float threshold, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                      int64_t *out_counts, cudaStream_t st, cm_flat_stats *stats);
    void free_shadow();
    int ensure_shadow(cudaStream_t st);
};

}  // namespace cm
