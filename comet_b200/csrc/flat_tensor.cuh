// flat_tensor.cuh -- pieces shared by the two candidate-pass kernels of the tensor path (flat_tensor.cu: both
// operands streamed through shared memory; flat_gemm_ts.cu: queries resident in tensor memory).
#pragma once

#include "common.cuh"

namespace cm {

static constexpr int MAX_PH = 6;           // phases of the candidate pass

struct GemmPhase {
    int cls;          // 0: tiles t % SA == 0; 1: t % SB == 0 && t % SA != 0; 2: t % SB != 0; 3: all tiles
    int SA, SB;
    int n_tiles;      // tiles in this class
    int dbg;          // debugging aid: 1 = epilogue skips the accumulator scan, 2 = loads TMEM but does not compare
    int dense;        // 1: (nearly) every score of this phase is a candidate (phase A): emit column by column;
                      // 2 (query-resident pass only): emit nothing, only derive a first bound from a row sample
};

__device__ __forceinline__ int phase_tile(const GemmPhase &p, int i) {
    switch (p.cls) {
    case 0: return i * p.SA;
    case 1: { int R = p.SA / p.SB; int j = i + i / (R - 1) + 1; return j * p.SB; }
    case 2: return i + i / (p.SB - 1) + 1;
    default: return i;
    }
}

// v[c] for a run-time c without spilling the register array: a 32-way switch (taken only on the rare
// candidate path)
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int c) {
    switch (c) {
#define CM_PICK(i) case i: return v[i];
        CM_PICK(0) CM_PICK(1) CM_PICK(2) CM_PICK(3) CM_PICK(4) CM_PICK(5) CM_PICK(6) CM_PICK(7)
        CM_PICK(8) CM_PICK(9) CM_PICK(10) CM_PICK(11) CM_PICK(12) CM_PICK(13) CM_PICK(14) CM_PICK(15)
        CM_PICK(16) CM_PICK(17) CM_PICK(18) CM_PICK(19) CM_PICK(20) CM_PICK(21) CM_PICK(22) CM_PICK(23)
        CM_PICK(24) CM_PICK(25) CM_PICK(26) CM_PICK(27) CM_PICK(28) CM_PICK(29) CM_PICK(30)
#undef CM_PICK
    default: return v[31];
    }
}

// The same pick as a 5-level select tree: 31 SELs, no divergence.  Used when many lanes of the warp hold a
// hit in the same pass (the branch tree above then runs once per DISTINCT column, up to 32 times).
__device__ __forceinline__ uint32_t pick32_sel(const uint32_t (&v)[32], int c) {
    uint32_t a[16], b[8], d[4], e[2];
    const bool p0 = (c & 1) != 0, p1 = (c & 2) != 0, p2 = (c & 4) != 0, p3 = (c & 8) != 0, p4 = (c & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = p0 ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = p1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = p2 ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; i++) e[i] = p3 ? d[2 * i + 1] : d[2 * i];
    return p4 ? e[1] : e[0];
}

// ---- query-resident candidate pass (flat_gemm_ts.cu) ----
static constexpr int TS_N = 64;            // corpus rows per work item (UMMA N; 32 per CTA of the pair)
static constexpr int TS_QBLK = 256;        // queries per CTA pair (UMMA M; 128 TMEM lanes per CTA)
static constexpr int TS_SLOTS = 256;       // candidate slots per (query, cluster, column half) region
static constexpr int TS_MAX_LDB = 768;     // bf16 elements of a query row that fit beside two accumulators in TMEM

// regions per query for a launch over n_qblk query blocks on n_clusters CTA pairs
inline int ts_regions(int n_clusters, int n_qblk) { return 2 * ((n_clusters + n_qblk - 1) / n_qblk); }

// Bound hand-over of the sampling phase (GemmPhase::dense == 2).  A query is served by n_threads = 2 x (CTA pairs
// of its query block) epilogue threads, each seeing its own rows.  Every thread keeps its j smallest keys,
// j = ceil(K / n_threads): the union of those lists holds >= K distinct rows, all with key <= max over threads of
// the j-th smallest, so that maximum bounds the K-th smallest key of the sample -- and of the corpus.  It is
// looser than an exact selection over the sample (the ~5 % quantile instead of 2 % at K = 100), but it needs no
// candidate lists, no selection kernel and no second launch boundary.
struct TsBound {
    int j;                        // 1..TS_BOUND_J; 0: this launch neither produces nor consumes a sampled bound
    int read_bits;                // this phase takes its bound from gmax_bits (the phase after the sampling phase)
    const float2 *q_norms;        // |q|, |q - bf16(q)| per query (error bound E_q, header of flat_tensor.cu)
    const unsigned int *max_bits; // max row norm, max row residual norm
    int dim;
    float e_scale;
    unsigned int *gmax_bits;      // [nq_pad] float_to_ordered(bound), combined with atomicMax; 0 = no bound (padding)
};
static constexpr int TS_BOUND_J = 4;

// One phase of the candidate pass over the tiles of `phase`: keys under the per-query bound go to
// cand[q][region][slot], fill counts to cand_cnt[q][region] (overwritten, not accumulated).
// has_h: row_h holds the key offsets [>= n_tiles_total * TS_N], +inf for rows that must not become candidates;
// !has_h (cosine, nothing masked): offsets are zero and rows >= n_rows are cut from the hit masks.
// tmap_q128: the bf16 queries [nq_pad][ldb], box 64 x 128 rows, SWIZZLE_128B.
int launch_gemm_ts(const CUtensorMap &tmap_x32, const CUtensorMap &tmap_q128, const GemmPhase &ph, int n_qblk, int ldb,
                   const float *row_h, bool has_h, int64_t n_rows, const float *g_bound, const TsBound &tsb, uint64_t *cand,
                   int *cand_cnt, cudaStream_t st);

// ---- per-query selection / fused finish for that pass (flat_finish.cu) ----
static constexpr int TS_RS_CAP = 2048;          // survivors (candidates re-scored) per query
static constexpr int TS_SEL_STAGE_CAP = 11264;  // keys (old survivors + new candidates) staged in shared memory (88 KB)

struct TsSelectArgs {
    int nq;
    const uint64_t *cand; int *cand_cnt; int n_reg, slots, K, dim;
    const float2 *q_norms; const unsigned int *max_bits; float *g; int *overflow;
    const uint64_t *surv_in; const int *surv_in_cnt; uint64_t *surv_out; int *surv_out_cnt;   // lists of TS_RS_CAP keys
    float e_scale; int *staged_max;
    // finish only
    const float *rows; int ld, ch; const float *queries; float threshold; const uint32_t *row_ids;
    int64_t out_stride; uint32_t *out_ids; float *out_scores; int64_t *out_pos; int64_t *out_counts;
    unsigned long long *rescored;
};
// finish = false: K-th smallest staged key -> g[q], survivors -> surv_out.  finish = true: the same selection, then
// the survivors are re-scored in reference order, sorted by (score, scan position) and written as the query's
// result; out_counts[q] = -1 when a candidate list overflowed (the caller redoes that query exactly).
int launch_ts_select(const TsSelectArgs &a, bool finish, int metric, bool fma, cudaStream_t st);

// Re-score of the last selection's survivors in reference order + per-query sort and output by the last CTA of
// each query.  keys2 [nq][TS_RS_CAP], keys2_cnt [nq] and done [nq] must be zero on entry.
struct RescoreFinishArgs {
    int nq, K;
    const float *rows; int ld; const float *queries;
    const uint64_t *rs; const int *rs_cnt;             // survivor lists of TS_RS_CAP keys
    float threshold;
    uint64_t *keys2; int *keys2_cnt; int *done; const int *overflow;
    const uint32_t *row_ids; int64_t out_stride; uint32_t *out_ids; float *out_scores; int64_t *out_pos; int64_t *out_counts;
    unsigned long long *rescored;
};
int launch_rescore_finish(const RescoreFinishArgs &a, int metric, bool fma, cudaStream_t st);

}  // namespace cm
