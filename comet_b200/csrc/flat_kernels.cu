// flat_kernels.cu -- exact (reference-order) flat scan for sm_100a, its top-K merge, and the
// small per-row kernels (preprocess, pair distances, skip mask, gather).
//
// flat_scan_kernel replaces the hot loop of flatIndexSearch.searchSingleQuery
// (flat_index_search.go:254-279): for every stored row, Distance.Calculate in the reference's
// exact float32 order (distance.go:114-121, 158-165, 201-216), soft-delete / document-filter /
// threshold tests, then "sort everything, keep k" expressed as a K-smallest selection.
//
// Data movement: the row-major [N][ld] fp32 matrix is streamed once from HBM by a dedicated TMA
// producer warp as 128-row x 32-float boxes (16 KB, SWIZZLE_128B) through an mbarrier ring.  Each
// of the 128 consumer threads owns one row of the tile and walks it left to right (the order is
// what makes scores bit-identical to the reference), reading 16-byte pieces whose XOR swizzle makes
// the 8 threads of every quarter-warp hit 8 distinct bank groups.  Up to QB=8 queries (staged once
// per CTA with a bulk copy, read as smem broadcasts) share each pass over the rows, which keeps the
// pass HBM-bound: 3 FP32 ops per element and query against 23 B/clk/SM of HBM.
#include "flat_kernels.cuh"
#include "select.cuh"

namespace cm {

// (packed fp32x2 arithmetic of a distance step: metric_step4_unfused, common.cuh)

// ------------------------------------------------------------------------------------------------
// exact scan
// ------------------------------------------------------------------------------------------------
template <int METRIC, bool FMA, int QB>
__global__ void __launch_bounds__(SCAN_THREADS) flat_scan_kernel(
    const __grid_constant__ CUtensorMap tmap, const float *__restrict__ queries, int ld, long long n_rows,
    int n_tiles, int stages, const uint8_t *__restrict__ skip, float threshold, int K, int C,
    uint64_t *__restrict__ part_keys, int *__restrict__ part_counts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: [stages x 16 KB][q: QB x ld f32][bufs: QB x C u64][cnt QB][tau QB][bars]
    uint8_t *stage_base = smem;
    float *q_s = reinterpret_cast<float *>(smem + (size_t)stages * SCAN_STAGE_BYTES);
    uint64_t *bufs = reinterpret_cast<uint64_t *>(q_s + (size_t)QB * ld);
    uint64_t *tau = bufs + (size_t)QB * C;
    uint64_t *full_bar = tau + QB;
    uint64_t *empty_bar = full_bar + stages;
    uint64_t *q_bar = empty_bar + stages;
    int *cnt = reinterpret_cast<int *>(q_bar + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_chunks = ld / SCAN_CHUNK;
    // blockIdx.y = query group (QB queries each): small tables are scanned for many groups in one launch
    queries += (size_t)blockIdx.y * QB * ld;
    part_keys += (size_t)blockIdx.y * QB * gridDim.x * K;
    part_counts += (size_t)blockIdx.y * QB * gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], SCAN_TILE_ROWS / 32);
        }
        mbar_init(q_bar, 1);
        fence_mbar_init();
    }
    if (tid < QB) { cnt[tid] = 0; tau[tid] = KEY_INF; }
    __syncthreads();

    if (warp == SCAN_TILE_ROWS / 32) {
        // ===== TMA producer warp =====
        if (lane == 0) {
            prefetch_tmap(&tmap);
            mbar_arrive_expect_tx(q_bar, (uint32_t)(QB * ld * sizeof(float)));
            bulk_load_1d(q_s, queries, (uint32_t)(QB * ld * sizeof(float)), q_bar);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int c = 0; c < n_chunks; c++, it++) {
                    int s = it % stages;
                    uint32_t ph = (it / stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], SCAN_STAGE_BYTES);
                    tma_load_2d(stage_base + (size_t)s * SCAN_STAGE_BYTES, &tmap, c * SCAN_CHUNK,
                                t * SCAN_TILE_ROWS, &full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumer warps: thread r owns row r of every tile =====
    const NamedBarrier bar{1, SCAN_TILE_ROWS};
    mbar_wait(q_bar, 0);
    const int r = tid;
    const uint32_t row_off = (uint32_t)r * (SCAN_CHUNK * 4);
    const uint32_t swz = (uint32_t)(r & 7);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        float acc[QB];
#pragma unroll
        for (int qi = 0; qi < QB; qi++) acc[qi] = 0.0f;
        for (int c = 0; c < n_chunks; c++, it++) {
            int s = it % stages;
            uint32_t ph = (it / stages) & 1;
            mbar_wait(&full_bar[s], ph);
            const uint8_t *sp = stage_base + (size_t)s * SCAN_STAGE_BYTES + row_off;
            float4 xv[SCAN_CHUNK / 4];
#pragma unroll
            for (int j = 0; j < SCAN_CHUNK / 4; j++)
                xv[j] = *reinterpret_cast<const float4 *>(sp + (((uint32_t)j ^ swz) << 4));
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);   // row data is in registers: free the slot
            const float *qc = q_s + c * SCAN_CHUNK;
#pragma unroll
            for (int j = 0; j < SCAN_CHUNK / 4; j++) {
#pragma unroll
                for (int qi = 0; qi < QB; qi++) {
                    float4 qv = *reinterpret_cast<const float4 *>(qc + (size_t)qi * ld + j * 4);
                    float a = acc[qi];
                    if (FMA) {
                        a = metric_step<METRIC, FMA>(a, qv.x, xv[j].x);
                        a = metric_step<METRIC, FMA>(a, qv.y, xv[j].y);
                        a = metric_step<METRIC, FMA>(a, qv.z, xv[j].z);
                        a = metric_step<METRIC, FMA>(a, qv.w, xv[j].w);
                    } else {
                        a = metric_step4_unfused<METRIC>(a, qv, xv[j]);
                    }
                    acc[qi] = a;
                }
            }
        }
        // ---- selection for this tile ----
        long long row = (long long)t * SCAN_TILE_ROWS + r;
        bool live = row < n_rows;
        if (live && skip != nullptr) live = skip[row] == 0;
        bar.sync();   // appends of the previous tile are visible
        unsigned need = 0;
#pragma unroll
        for (int qi = 0; qi < QB; qi++) need |= (cnt[qi] > C - SCAN_TILE_ROWS ? 1u : 0u) << qi;
        bar.sync();   // everyone has taken the same decision before any count moves again
#pragma unroll 1
        for (int qi = 0; qi < QB; qi++) {
            if (need & (1u << qi))
                compact_topk(bufs + (size_t)qi * C, C, K, &cnt[qi], &tau[qi], tid, SCAN_TILE_ROWS, bar);
        }
#pragma unroll
        for (int qi = 0; qi < QB; qi++) {
            float dist = metric_finish<METRIC>(acc[qi]);
            bool pass = live && !(threshold > 0.0f && dist > threshold);
            if (pass) {
                uint64_t key = make_key(dist, (uint32_t)row);
                if (key < tau[qi]) {
                    int slot = atomicAdd(&cnt[qi], 1);
                    bufs[(size_t)qi * C + slot] = key;
                }
            }
        }
    }
    // ---- CTA result: K smallest per query ----
#pragma unroll 1
    for (int qi = 0; qi < QB; qi++) {
        compact_topk(bufs + (size_t)qi * C, C, K, &cnt[qi], &tau[qi], tid, SCAN_TILE_ROWS, bar);
        int m = cnt[qi];
        uint64_t *dst = part_keys + ((size_t)qi * gridDim.x + blockIdx.x) * K;
        for (int i = tid; i < m; i += SCAN_TILE_ROWS) dst[i] = bufs[(size_t)qi * C + i];
        if (tid == 0) part_counts[(size_t)qi * gridDim.x + blockIdx.x] = m;
    }
}

static size_t scan_smem_bytes(int qb, int ld, int stages, int C) {
    return (size_t)stages * SCAN_STAGE_BYTES + (size_t)qb * ld * 4 + (size_t)qb * C * 8 + (size_t)qb * 8 +
           (size_t)(2 * stages + 1) * 8 + (size_t)qb * 4 + 16;
}

int plan_scan(int metric, bool fma, int nq, int ld, int64_t n_rows, int K, ScanLaunch *out) {
    int C = next_pow2(K + 2 * SCAN_TILE_ROWS);
    if (C < 512) C = 512;
    size_t cap = max_smem_optin();
    int qb = nq >= 8 ? 8 : (nq >= 4 ? 4 : (nq >= 2 ? 2 : 1));
    int stages = 0, ctas_per_sm = 1;
    // Each consumer thread walks one row sequentially, so a CTA issues from only 4 warps: the scan is
    // latency bound unless several CTAs share an SM.  Prefer the largest CTA count per SM whose footprint
    // (ring of 3-4 stages) fits; fall back to one CTA per SM with a deep ring, then to a smaller query block
    // (small query blocks leave room for three or four).
    for (int ctas = 4; ctas >= 2 && stages == 0; ctas--) {
        const size_t share = (cap + 1024) / ctas - 2048;      // per-CTA budget incl. the reserved kilobyte
        for (int s = 4; s >= 3; s--)
            if (scan_smem_bytes(qb, ld, s, C) <= share) { stages = s; ctas_per_sm = ctas; break; }
    }
    if (stages == 0) {
        for (;; qb >>= 1) {
            for (stages = 8; stages >= 3; stages--)
                if (scan_smem_bytes(qb, ld, stages, C) <= cap) break;
            if (stages >= 3 || qb == 1) break;
        }
    }
    if (stages < 3) return fail(CM_ERR_UNSUPPORTED, "k=%d with dim pad %d does not fit the scan kernel's shared memory", K, ld);
    int64_t n_tiles = (n_rows + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    int grid = sm_count() * ctas_per_sm;
    out->slots = grid;
    if (grid > n_tiles) grid = (int)(n_tiles > 0 ? n_tiles : 1);
    out->metric = metric; out->fma = fma; out->qb = qb; out->stages = stages; out->grid = grid;
    out->K = K; out->C = C; out->smem = scan_smem_bytes(qb, ld, stages, C);
    return CM_OK;
}

template <int METRIC, bool FMA, int QB>
static int launch_scan_t(const ScanLaunch &L, const CUtensorMap &tmap, const float *queries, int ld,
                         int64_t n_rows, const uint8_t *skip, float threshold, uint64_t *part_keys,
                         int *part_counts, int n_groups, cudaStream_t stream) {
    auto kern = flat_scan_kernel<METRIC, FMA, QB>;
    CM_TRY(set_dyn_smem((const void *)kern, L.smem));
    int n_tiles = (int)((n_rows + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS);
    ProfScope prof(CM_PROF_FLAT_SCAN, stream);
    kern<<<dim3((unsigned)L.grid, (unsigned)n_groups), SCAN_THREADS, L.smem, stream>>>(tmap, queries, ld, (long long)n_rows, n_tiles, L.stages, skip,
                                                   threshold, L.K, L.C, part_keys, part_counts);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

template <int METRIC, bool FMA>
static int launch_scan_q(const ScanLaunch &L, const CUtensorMap &tmap, const float *queries, int ld,
                         int64_t n_rows, const uint8_t *skip, float threshold, uint64_t *pk, int *pc, int n_groups,
                         cudaStream_t st) {
    switch (L.qb) {
    case 1: return launch_scan_t<METRIC, FMA, 1>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st);
    case 2: return launch_scan_t<METRIC, FMA, 2>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st);
    case 4: return launch_scan_t<METRIC, FMA, 4>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st);
    case 8: return launch_scan_t<METRIC, FMA, 8>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st);
    }
    return fail(CM_ERR_INVALID_ARG, "bad query block %d", L.qb);
}

int launch_flat_scan(const ScanLaunch &L, const CUtensorMap &tmap, const float *queries, int ld,
                     int64_t n_rows, const uint8_t *skip, float threshold, uint64_t *pk, int *pc, int n_groups,
                     cudaStream_t st) {
#define CM_SCAN_CASE(M)                                                                                  \
    case M:                                                                                              \
        return L.fma ? launch_scan_q<M, true>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st) \
                     : launch_scan_q<M, false>(L, tmap, queries, ld, n_rows, skip, threshold, pk, pc, n_groups, st);
    switch (L.metric) {
        CM_SCAN_CASE(CM_L2)
        CM_SCAN_CASE(CM_L2SQ)
        CM_SCAN_CASE(CM_COSINE)
    }
#undef CM_SCAN_CASE
    return fail(CM_ERR_INVALID_ARG, "unknown metric %d", L.metric);
}

// ------------------------------------------------------------------------------------------------
// merge of per-CTA partial lists -> final sorted top-K per query
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MERGE_THREADS) merge_topk_kernel(
    const uint64_t *part_keys, const int *part_counts, int parts, int Kp, int K, int C,    // (no __restrict__: see pdl_wait())
    const uint32_t *__restrict__ row_ids, long long out_stride, uint32_t *__restrict__ out_ids,
    float *__restrict__ out_scores, long long *__restrict__ out_pos, long long *__restrict__ out_counts,
    const int *overflow, const int *stat_cnt, unsigned long long *stat_sum) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *buf = reinterpret_cast<uint64_t *>(smem);
    __shared__ int cnt;
    __shared__ uint64_t tau;
    const CtaBarrier bar;
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    pdl_wait();             // no-op unless launched with the programmatic-dependent-launch attribute
    pdl_trigger();
    if (tid == 0) { cnt = 0; tau = KEY_INF; }
    __syncthreads();
    const uint64_t *pk = part_keys + (size_t)q * parts * Kp;
    const int *pc = part_counts + (size_t)q * parts;
    // walk the parts in groups whose worst-case appends fit the free space
    int p = 0;
    while (p < parts) {
        int room = C - cnt;
        uint64_t t = tau;
        __syncthreads();   // same decision in every thread before counts move
        int p_end = p, need = 0;
        while (p_end < parts && need + pc[p_end] <= room) { need += pc[p_end]; p_end++; }
        if (p_end == p) {   // no room for the next part: compact first
            compact_topk(buf, C, K, &cnt, &tau, tid, MERGE_THREADS, bar);
            continue;
        }
        // the group's parts as one index space (part, slot): independent loads, not a latency chain per part (an IVF
        // query has thousands of 128-slot parts)
        const int span = p_end - p == 1 ? pc[p] : (p_end - p) * Kp;      // a single part: just its keys
        for (int idx = tid; idx < span; idx += MERGE_THREADS) {
            const int dp = idx / Kp, i = idx - dp * Kp;
            if (i < pc[p + dp]) {
                uint64_t key = pk[(size_t)(p + dp) * Kp + i];
                if (key < t) buf[atomicAdd(&cnt, 1)] = key;
            }
        }
        __syncthreads();
        p = p_end;
    }
    compact_topk(buf, C, K, &cnt, &tau, tid, MERGE_THREADS, bar);
    int m = cnt;
    for (int i = tid; i < m; i += MERGE_THREADS) {
        uint64_t key = buf[i];
        uint32_t pos = key_pos(key);
        size_t o = (size_t)q * out_stride + i;
        out_ids[o] = row_ids ? row_ids[pos] : pos;
        out_scores[o] = key_score(key);
        if (out_pos) out_pos[o] = pos;
    }
    if (tid == 0) {
        // tensor path: a query whose candidate lists overflowed gets count -1 (the host entry point redoes it with the
        // exact scan); stat_sum accumulates the candidates that went through the exact re-score
        out_counts[q] = (overflow && ld_pdl_s32(overflow + q)) ? -1 : m;
        if (stat_sum && stat_cnt) { const int c = ld_pdl_s32(stat_cnt + q); if (c > 0) atomicAdd(stat_sum, (unsigned long long)c); }
    }
}

int launch_merge_topk(const uint64_t *part_keys, const int *part_counts, int nq, int parts, int Kp, int K,
                      const uint32_t *row_ids, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                      int64_t *out_pos, int64_t *out_counts, cudaStream_t stream, bool pdl, const int *overflow,
                      const int *stat_cnt, unsigned long long *stat_sum) {
    int C = next_pow2(K + Kp);
    if (C < 2048) C = 2048;
    size_t smem = (size_t)C * 8;
    if (smem > max_smem_optin()) {    // k <= 0 / huge k: one radix sort per query over the part lists (flat_bigk.cu)
        if (pdl || overflow || stat_cnt) return fail(CM_ERR_UNSUPPORTED, "k=%d too large for the merge kernel", K);
        return merge_topk_bigk(part_keys, part_counts, nq, parts, Kp, (int64_t)K, row_ids, out_stride, out_ids, out_scores, out_pos,
                               out_counts, stream);
    }
    CM_TRY(set_dyn_smem((const void *)merge_topk_kernel, smem));
    ProfScope prof(CM_PROF_SELECT, stream);
    PdlLaunch L(dim3((unsigned)nq), dim3(MERGE_THREADS), smem, stream, 0, pdl);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, merge_topk_kernel, part_keys, part_counts, parts, Kp, K, C, row_ids,
                               (long long)out_stride, out_ids, out_scores, (long long *)out_pos, (long long *)out_counts,
                               overflow, stat_cnt, stat_sum));
    count_launch();
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// merge of per-SHARD results (multi-GPU): shard r holds rows [r*shard, (r+1)*shard) of the corpus and
// returns its own sorted top-K; the global order (score, scan position) is (score, shard, rank in
// the shard's list), so the merge needs nothing but the gathered lists.
// ------------------------------------------------------------------------------------------------
// Every gathered entry finds its own place: its rank in the union is its rank in its own (sorted) list plus, for every
// other shard, the number of that shard's entries that order before it -- a binary search per shard on the ordered score
// bits, ties going to the lower shard.  No selection, no sort, no atomics; W x K threads-worth of independent work.
static constexpr int MERGE_RANK_THREADS = 256;
__global__ void __launch_bounds__(MERGE_RANK_THREADS) merge_shards_kernel(
    const uint32_t *__restrict__ ids, const float *__restrict__ scores, const long long *__restrict__ counts, int world,
    long long nq, long long in_stride, int K, long long out_stride, uint32_t *__restrict__ out_ids,
    float *__restrict__ out_scores, long long *__restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *ord = reinterpret_cast<uint32_t *>(smem);                 // [world][in_stride] ordered score bits
    __shared__ int cnt_s[64];
    __shared__ long long bad_s;
    const int tid = threadIdx.x;
    const long long q = blockIdx.x;
    // A shard that could not answer a query (tensor-path candidate overflow: count -1, see cm_flat_search_device)
    // makes the merged answer unknown too: the -1 is passed on, never silently treated as an empty list;
    // -2 (zero query under cosine: every non-empty shard says so) wins over -1.
    if (tid == 0) {
        long long bad = 0, total = 0;
        for (int r = 0; r < world; r++) {
            long long c = counts ? counts[(size_t)r * nq + q] : in_stride;
            bad = min(bad, c);
            c = max(0ll, min(c, in_stride));
            cnt_s[r] = (int)c;
            total += c;
        }
        bad_s = bad;
        if (out_counts) out_counts[q] = bad < 0 ? bad : min((long long)K, total);
    }
    __syncthreads();
    if (bad_s < 0) return;
    const long long n_all = (long long)world * in_stride;
    for (long long e = tid; e < n_all; e += MERGE_RANK_THREADS) {
        const int r = (int)(e / in_stride), j = (int)(e - (long long)r * in_stride);
        if (j < cnt_s[r]) ord[e] = float_to_ordered(scores[((size_t)r * nq + q) * in_stride + j]);
    }
    __syncthreads();
    for (long long e = tid; e < n_all; e += MERGE_RANK_THREADS) {
        const int r = (int)(e / in_stride), j = (int)(e - (long long)r * in_stride);
        if (j >= cnt_s[r]) continue;
        const uint32_t mine = ord[e];
        long long rank = j;
        for (int r2 = 0; r2 < world && rank < K; r2++) {
            if (r2 == r) continue;
            const uint32_t *o2 = ord + (size_t)r2 * in_stride;
            int lo = 0, hi = cnt_s[r2];
            if (r2 < r) { while (lo < hi) { int mid = (lo + hi) >> 1; if (o2[mid] <= mine) lo = mid + 1; else hi = mid; } }   // ties: lower shard first
            else        { while (lo < hi) { int mid = (lo + hi) >> 1; if (o2[mid] < mine) lo = mid + 1; else hi = mid; } }
            rank += lo;
        }
        if (rank < K) {
            const size_t src = ((size_t)r * nq + q) * in_stride + j;
            out_ids[(size_t)q * out_stride + rank] = ids[src];
            out_scores[(size_t)q * out_stride + rank] = scores[src];
        }
    }
}

int launch_merge_shards(const uint32_t *ids, const float *scores, const int64_t *counts, int world, int64_t nq,
                        int64_t in_stride, int K, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                        int64_t *out_counts, cudaStream_t stream) {
    if (nq <= 0) return CM_OK;
    size_t smem = (size_t)world * in_stride * 4;
    if (smem + 1024 > max_smem_optin() || world > 64)    // k <= 0 / huge k: one radix sort per query over the gathered lists (flat_bigk.cu)
        return merge_shards_bigk(ids, scores, counts, world, nq, in_stride, (int64_t)K, out_stride, out_ids, out_scores, out_counts, stream);
    CM_TRY(set_dyn_smem((const void *)merge_shards_kernel, smem));
    ProfScope prof(CM_PROF_SELECT, stream);
    merge_shards_kernel<<<(unsigned)nq, MERGE_RANK_THREADS, smem, stream>>>(ids, scores, (const long long *)counts, world,
                                                                           (long long)nq, (long long)in_stride, K,
                                                                           (long long)out_stride, out_ids, out_scores,
                                                                           (long long *)out_counts);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// merge of per-shard results when the shards hold LISTS (IVF / IVFPQ list shards): the reference orders ties by the
// candidate's number in its append loop over the probed lists (ivf_index_search.go:252-308), a number every shard
// reports next to its scores (`gno`, unique per query across shards).  Order = (score, gno); the id of a winner is found
// again by a binary search of its key in each shard's (sorted) list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MERGE_THREADS) merge_keyed_shards_kernel(
    const uint32_t *__restrict__ ids, const float *__restrict__ scores, const uint32_t *__restrict__ gno,
    const long long *__restrict__ counts, int world, long long nq, long long in_stride, int K, int C, long long out_stride,
    uint32_t *__restrict__ out_ids, float *__restrict__ out_scores, long long *__restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *buf = reinterpret_cast<uint64_t *>(smem);
    __shared__ int cnt;
    __shared__ uint64_t tau;
    const CtaBarrier bar;
    const int tid = threadIdx.x;
    const long long q = blockIdx.x;
    if (tid == 0) { cnt = 0; tau = KEY_INF; }
    __syncthreads();
    long long bad = 0;
    for (int r = 0; r < world; r++) bad = min(bad, counts[(size_t)r * nq + q]);
    if (bad < 0) {
        if (tid == 0) out_counts[q] = bad;
        return;
    }
    for (int r = 0; r < world; r++) {
        long long c = min(counts[(size_t)r * nq + q], in_stride);
        const size_t base = ((size_t)r * nq + q) * in_stride;
        for (int j = tid; j < c; j += MERGE_THREADS) buf[atomicAdd(&cnt, 1)] = make_key(scores[base + j], gno[base + j]);
    }
    compact_topk(buf, C, K, &cnt, &tau, tid, MERGE_THREADS, bar);
    const int m = cnt;
    for (int i = tid; i < m; i += MERGE_THREADS) {
        const uint64_t key = buf[i];
        uint32_t id = 0;
        for (int r = 0; r < world; r++) {
            const long long c = min(counts[(size_t)r * nq + q], in_stride);
            const size_t base = ((size_t)r * nq + q) * in_stride;
            long long lo = 0, hi = c;                 // first j with key(j) >= key
            while (lo < hi) {
                long long mid = (lo + hi) >> 1;
                if (make_key(scores[base + mid], gno[base + mid]) < key) lo = mid + 1; else hi = mid;
            }
            if (lo < c && make_key(scores[base + lo], gno[base + lo]) == key) { id = ids[base + lo]; break; }
        }
        out_ids[(size_t)q * out_stride + i] = id;
        out_scores[(size_t)q * out_stride + i] = key_score(key);
    }
    if (tid == 0) out_counts[q] = m;
}

int launch_merge_keyed_shards(const uint32_t *ids, const float *scores, const uint32_t *gno, const int64_t *counts, int world,
                              int64_t nq, int64_t in_stride, int K, int64_t out_stride, uint32_t *out_ids, float *out_scores,
                              int64_t *out_counts, cudaStream_t stream) {
    if (nq <= 0) return CM_OK;
    int C = next_pow2((int)(world * in_stride));
    if (C < 512) C = 512;
    size_t smem = (size_t)C * 8;
    if (smem > max_smem_optin())
        return fail(CM_ERR_UNSUPPORTED, "%d list shards x k=%lld too large for the shard merge", world, (long long)in_stride);
    CM_TRY(set_dyn_smem((const void *)merge_keyed_shards_kernel, smem));
    ProfScope prof(CM_PROF_SELECT, stream);
    merge_keyed_shards_kernel<<<(unsigned)nq, MERGE_THREADS, smem, stream>>>(ids, scores, gno, (const long long *)counts, world,
                                                                            (long long)nq, (long long)in_stride, K, C,
                                                                            (long long)out_stride, out_ids, out_scores,
                                                                            (long long *)out_counts);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// per-row helpers
// ------------------------------------------------------------------------------------------------
// distance.go:244-264 PreprocessInPlace / :269-290 Preprocess.  One thread per row, sequential.
template <bool FMA>
__global__ void preprocess_rows_kernel(int metric, const float *__restrict__ src, long long n, int dim, int ld_src,
                                       float *__restrict__ dst, int ld_dst, int *__restrict__ zero_flags) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *s = src + (size_t)i * ld_src;
    float *d = dst + (size_t)i * ld_dst;
    float scale = 1.0f;
    int zero = 0;
    if (metric == CM_COSINE) {
        float sum = 0.0f;
        for (int j = 0; j < dim; j++) sum = dot_step<FMA>(sum, s[j], s[j]);
        float norm = __fsqrt_rn(sum);
        if (norm == 0.0f) zero = 1;
        scale = __fdiv_rn(1.0f, norm);
    }
    if (zero_flags) zero_flags[i] = zero;
    if (metric == CM_COSINE && !zero) {
        for (int j = 0; j < dim; j++) d[j] = __fmul_rn(s[j], scale);
    } else if (d != s) {
        for (int j = 0; j < dim; j++) d[j] = s[j];
    }
    if (d != s || ld_dst > dim)
        for (int j = dim; j < ld_dst; j++) d[j] = 0.0f;
}

// Same arithmetic for a SMALL number of rows (a query batch): one warp per row stages the row in
// shared memory with coalesced loads, lane 0 runs the reference's sequential sum, all lanes scale.
// (One thread per row is latency-bound there: 768 dependent strided loads per query.)
template <bool FMA>
__global__ void preprocess_rows_warp_kernel(int metric, const float *__restrict__ src, long long n, int dim, int ld_src,
                                            float *__restrict__ dst, int ld_dst, int *__restrict__ zero_flags) {
    extern __shared__ float prow[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long i = blockIdx.x * (long long)(blockDim.x >> 5) + w;
    if (i >= n) return;
    float *r = prow + (size_t)w * dim;
    const float *s = src + (size_t)i * ld_src;
    float *d = dst + (size_t)i * ld_dst;
    for (int j = lane; j < dim; j += 32) r[j] = s[j];
    __syncwarp();
    float scale = 1.0f;
    int zero = 0;
    if (metric == CM_COSINE) {
        if (lane == 0) {
            float sum = 0.0f;
            for (int j = 0; j < dim; j++) sum = dot_step<FMA>(sum, r[j], r[j]);
            float norm = __fsqrt_rn(sum);
            zero = norm == 0.0f;
            scale = __fdiv_rn(1.0f, norm);
        }
        scale = __shfl_sync(0xffffffffu, scale, 0);
        zero = __shfl_sync(0xffffffffu, zero, 0);
    }
    if (zero_flags && lane == 0) zero_flags[i] = zero;
    bool do_scale = metric == CM_COSINE && !zero;
    for (int j = lane; j < ld_dst; j += 32) {
        float v = 0.0f;
        if (j < dim) v = do_scale ? __fmul_rn(r[j], scale) : r[j];
        d[j] = v;
    }
}

int launch_preprocess_rows(int metric, bool fma, const float *src, int64_t n, int dim, int ld_src, float *dst,
                           int ld_dst, int *zero_flags, cudaStream_t stream) {
    if (n <= 0) return CM_OK;
    if (n <= 16384 && src != dst && (size_t)dim * 4 * 4 <= 48 * 1024) {
        int warps = 4;
        size_t smem = (size_t)warps * dim * 4;
        unsigned blocks = (unsigned)((n + warps - 1) / warps);
        if (fma)
            preprocess_rows_warp_kernel<true><<<blocks, warps * 32, smem, stream>>>(metric, src, n, dim, ld_src, dst, ld_dst, zero_flags);
        else
            preprocess_rows_warp_kernel<false><<<blocks, warps * 32, smem, stream>>>(metric, src, n, dim, ld_src, dst, ld_dst, zero_flags);
        count_launch();
        CM_CUDA(cudaGetLastError());
        return CM_OK;
    }
    int threads = 128;
    long long blocks = (n + threads - 1) / threads;
    if (fma)
        preprocess_rows_kernel<true><<<(unsigned)blocks, threads, 0, stream>>>(metric, src, n, dim, ld_src, dst, ld_dst, zero_flags);
    else
        preprocess_rows_kernel<false><<<(unsigned)blocks, threads, 0, stream>>>(metric, src, n, dim, ld_src, dst, ld_dst, zero_flags);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

template <int METRIC, bool FMA>
__global__ void distance_pairs_kernel(const float *__restrict__ a, const float *__restrict__ b, long long n, int dim,
                                      float *__restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *x = a + (size_t)i * dim, *y = b + (size_t)i * dim;
    float acc = 0.0f;
    for (int j = 0; j < dim; j++) acc = metric_step<METRIC, FMA>(acc, x[j], y[j]);
    out[i] = metric_finish<METRIC>(acc);
}

int launch_distance_pairs(int metric, bool fma, const float *a, const float *b, int64_t n, int dim, float *out,
                          cudaStream_t stream) {
    if (n <= 0) return CM_OK;
    int threads = 128;
    unsigned blocks = (unsigned)((n + threads - 1) / threads);
#define CM_PAIR_CASE(M)                                                                           \
    case M:                                                                                       \
        if (fma) distance_pairs_kernel<M, true><<<blocks, threads, 0, stream>>>(a, b, n, dim, out); \
        else distance_pairs_kernel<M, false><<<blocks, threads, 0, stream>>>(a, b, n, dim, out);    \
        break;
    switch (metric) {
        CM_PAIR_CASE(CM_L2)
        CM_PAIR_CASE(CM_L2SQ)
        CM_PAIR_CASE(CM_COSINE)
    default: return fail(CM_ERR_INVALID_ARG, "unknown metric %d", metric);
    }
#undef CM_PAIR_CASE
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

__global__ void fill_counts_kernel(long long *__restrict__ counts, long long nq, long long v) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q < nq) counts[q] = v;
}
int launch_fill_counts(int64_t *counts, int64_t nq, int64_t value, cudaStream_t stream) {
    if (nq <= 0) return CM_OK;
    fill_counts_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>((long long *)counts, (long long)nq, (long long)value);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// device entry points report a zero query under cosine (ErrZeroVector, distance.go:269-290) as count -2
__global__ void mark_zero_queries_kernel(const int *flags, long long nq, long long *out_counts) {
    pdl_wait();                 // a link of the step's programmatic-dependent-launch chain: behind the kernels that write the counts
    pdl_trigger();
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q < nq && flags[q]) out_counts[q] = -2;
}
int launch_mark_zero_queries(const int *flags, int64_t nq, int64_t *out_counts, cudaStream_t stream) {
    if (nq <= 0) return CM_OK;
    PdlLaunch L(dim3((unsigned)((nq + 255) / 256)), dim3(256), 0, stream);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, mark_zero_queries_kernel, flags, (long long)nq, (long long *)out_counts));
    count_launch();
    return CM_OK;
}

__global__ void build_skip_kernel(const uint32_t *__restrict__ row_ids, const uint8_t *__restrict__ deleted, long long n,
                                  const uint32_t *__restrict__ filt, long long nf, uint8_t *__restrict__ skip) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t s = deleted ? deleted[i] : 0;
    if (!s && nf > 0) {
        uint32_t id = row_ids[i];
        long long lo = 0, hi = nf;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (filt[mid] < id) lo = mid + 1; else hi = mid;
        }
        s = !(lo < nf && filt[lo] == id);
    }
    skip[i] = s;
}

int launch_build_skip(const uint32_t *row_ids, const uint8_t *deleted, int64_t n, const uint32_t *filter_sorted,
                      int64_t nfilter, uint8_t *skip, cudaStream_t stream) {
    if (n <= 0) return CM_OK;
    int threads = 256;
    unsigned blocks = (unsigned)((n + threads - 1) / threads);
    build_skip_kernel<<<blocks, threads, 0, stream>>>(row_ids, deleted, n, filter_sorted, nfilter, skip);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

__global__ void gather_rows_kernel(const float *__restrict__ rows, int ld, int dim, const long long *__restrict__ pos,
                                   long long n, float *__restrict__ out) {
    long long i = blockIdx.x;
    if (i >= n) return;
    const float *s = rows + (size_t)pos[i] * ld;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) out[(size_t)i * dim + j] = s[j];
}

int launch_gather_rows(const float *rows, int ld, int dim, const int64_t *pos, int64_t n, float *out,
                       cudaStream_t stream) {
    if (n <= 0) return CM_OK;
    gather_rows_kernel<<<(unsigned)n, 128, 0, stream>>>(rows, ld, dim, (const long long *)pos, n, out);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

}  // namespace cm
