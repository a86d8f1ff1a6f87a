// flat_kernels.cuh -- launch wrappers of the flat-index device kernels (flat_kernels.cu).
#pragma once

#include "common.cuh"

namespace cm {

static constexpr int SCAN_TILE_ROWS = 128;   // rows per tile == consumer threads per CTA
static constexpr int SCAN_CHUNK = 32;        // floats per TMA box row (128 B, SWIZZLE_128B)
static constexpr int SCAN_STAGE_BYTES = SCAN_TILE_ROWS * SCAN_CHUNK * 4;
static constexpr int SCAN_THREADS = SCAN_TILE_ROWS + 32;  // 4 consumer warps + 1 TMA producer warp
static constexpr int SCAN_MAX_QB = 8;
static constexpr int MERGE_THREADS = 256;

struct ScanLaunch {
    int metric;
    bool fma;
    int qb;            // queries per pass (1,2,4,8)
    int stages;        // TMA ring depth
    int grid;          // persistent CTAs
    int slots;         // CTAs of this footprint the device holds at once
    int K, C;          // keep K, buffer capacity C (pow2, C >= K + SCAN_TILE_ROWS)
    size_t smem;
};

// Pick QB / stages / grid for a scan of `n_rows` rows with leading dimension ld floats.
int plan_scan(int metric, bool fma, int nq, int ld, int64_t n_rows, int K, ScanLaunch *out);

// One launch: n_groups groups of QB queries (device, [n_groups * qb][ld], preprocessed, zero padded)
// against all rows (blockIdx.y = group).  part_keys: [n_groups * qb][grid][K], part_counts likewise.
int launch_flat_scan(const ScanLaunch &L, const CUtensorMap &tmap, const float *queries, int ld,
                     int64_t n_rows, const uint8_t *skip, float threshold, uint64_t *part_keys,
                     int *part_counts, int n_groups, cudaStream_t stream);

// Merge `parts` partial lists per query into the final sorted top-K and materialise outputs.
// part_keys: [nq][parts][Kp]; outputs are [nq][out_stride].
int launch_merge_topk(const uint64_t *part_keys, const int *part_counts, int nq, int parts, int Kp, int K,
                      const uint32_t *row_ids, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                      int64_t *out_pos, int64_t *out_counts, cudaStream_t stream, bool pdl = false,
                      const int *overflow = nullptr, const int *stat_cnt = nullptr, unsigned long long *stat_sum = nullptr);

// Merge `world` per-shard sorted result lists per query ([world][nq][in_stride], counts [world][nq] or
// NULL = full) into the global top-K ordered by (score, shard, rank within the shard).
int launch_merge_shards(const uint32_t *ids, const float *scores, const int64_t *counts, int world, int64_t nq,
                        int64_t in_stride, int K, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                        int64_t *out_counts, cudaStream_t stream);
// Merge of per-shard lists ordered by (score, gno): list shards of IVF / IVFPQ, gno = the candidate's number in the
// reference's append loop over all probed lists (unique per query across shards)
int launch_merge_keyed_shards(const uint32_t *ids, const float *scores, const uint32_t *gno, const int64_t *counts, int world,
                              int64_t nq, int64_t in_stride, int K, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                              int64_t *out_counts, cudaStream_t stream);
// launch_merge_topk's fallback when K + Kp does not fit shared memory: one radix sort per query (flat_bigk.cu)
int merge_topk_bigk(const uint64_t *part_keys, const int *part_counts, int nq, int parts, int Kp, int64_t K,
                    const uint32_t *row_ids, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                    int64_t *out_counts, cudaStream_t stream);
// the same for world x in_stride beyond the shared-memory merge: one radix sort per query (flat_bigk.cu)
int merge_shards_bigk(const uint32_t *ids, const float *scores, const int64_t *counts, int world, int64_t nq,
                      int64_t in_stride, int64_t k, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                      int64_t *out_counts, cudaStream_t stream);

// Distance.Preprocess / PreprocessInPlace on n rows (one thread per row, reference order).
// dst rows have leading dimension ld_dst (zero padded beyond dim); flags[i] = 1 for a zero vector.
int launch_preprocess_rows(int metric, bool fma, const float *src, int64_t n, int dim, int ld_src, float *dst,
                           int ld_dst, int *zero_flags, cudaStream_t stream);

// Distance.Calculate over n pairs (one thread per pair, reference order).
int launch_distance_pairs(int metric, bool fma, const float *a, const float *b, int64_t n, int dim, float *out,
                          cudaStream_t stream);

// counts[q] = value (a kernel rather than a memset: the array may live on a peer device)
int launch_fill_counts(int64_t *counts, int64_t nq, int64_t value, cudaStream_t stream);
// out_counts[q] = -2 where flags[q] != 0 (zero query under cosine, reported by the device entry points)
int launch_mark_zero_queries(const int *flags, int64_t nq, int64_t *out_counts, cudaStream_t stream);

// skip[row] = deleted[row] | (filter given && id[row] not in sorted filter)
int launch_build_skip(const uint32_t *row_ids, const uint8_t *deleted, int64_t n, const uint32_t *filter_sorted,
                      int64_t nfilter, uint8_t *skip, cudaStream_t stream);

// gather rows by position into a dense [n][dim] buffer
int launch_gather_rows(const float *rows, int ld, int dim, const int64_t *pos, int64_t n, float *out,
                       cudaStream_t stream);

}  // namespace cm
