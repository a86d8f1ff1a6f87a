// flat_finish.cu -- per-query selection for the query-resident candidate pass (flat_gemm_ts.cu), and the fused
// "finish" of a batched flat search: last selection + reference-order re-score + final sort in ONE kernel.
//
// One CTA per query.  The query's new candidates sit in n_reg regions (one per (cluster, column half) that served
// its query block), the survivors of earlier phases in a dense list.  Both are staged in shared memory; a radix
// select on the ordered score bits finds the K-th smallest key tau_K; everything under tau_K + 2 E_q (E_q: the
// proved bound on |tensor key - reference score|, header of flat_tensor.cu) survives.
//   * phases A, B (FINISH = false): survivors go back to global memory, g[q] = -(bound) feeds the next pass;
//   * last phase (FINISH = true): the survivors stay in shared memory, their fp32 rows are gathered with
//     cp.async (256-byte pieces, double buffered) and every row is walked by ONE thread in the reference's
//     summation order (distance.go:114-121 / 158-165 / 201-216) -- the scores are the reference's bit for bit --
//     then the (score, scan position) keys are sorted in place and the first K are written out: what
//     flatIndexSearch.searchSingleQuery returns (flat_index_search.go:277-291).
// Compared with select + re-score + merge + mark as four launches this keeps the survivor list on chip and lets
// the selection of one query overlap the row gathers of another on the same SM.
#include <algorithm>

#include "flat_tensor.cuh"
#include "select.cuh"

namespace cm {

static constexpr int FIN_THREADS = 512;              // 2 CTAs per SM at <= 64 registers (768 threads spill in the gather loop)
static constexpr int FIN_ROWS = 128;                    // candidates re-scored per round (one compute thread each)
static constexpr int FIN_TILE_ROWS = 64;                // rows per gather tile
static constexpr int FIN_RING = 5;                      // gather tiles in shared memory (FIN_RING - 1 in flight)
static constexpr int FIN_MAX_REG = 160;                 // candidate regions per query (2 x CTA pairs per query block)

template <bool FINISH, int METRIC, bool FMA>
__global__ void __launch_bounds__(FIN_THREADS, 2) ts_select_kernel(
    // no __restrict__ / __ldg on anything the preceding kernels write: see pdl_wait() in common.cuh
    const uint64_t *cand, int *cand_cnt, int n_reg, int slots, int K, int dim,
    const float2 *q_norms, const unsigned int *max_bits, float *g,
    int *overflow, const uint64_t *surv_in, const int *surv_in_cnt,
    uint64_t *surv_out, int *surv_out_cnt, int surv_cap, int stage_cap, float e_scale,
    int *staged_max,
    // FINISH only
    const float *rows, int ld, int ch, const float *queries, float threshold,
    const uint32_t *row_ids, long long out_stride, uint32_t *out_ids,
    float *out_scores, long long *out_pos, long long *out_counts,
    unsigned long long *rescored) {
    constexpr int NT = FIN_THREADS, NW = NT / 32;
    extern __shared__ __align__(16) uint8_t fin_smem[];
    uint64_t *key_s = reinterpret_cast<uint64_t *>(fin_smem);                            // [stage_cap]
    uint64_t *surv_s = key_s + stage_cap;                                                // [surv_cap] (FINISH)
    __shared__ int hist[256];
    __shared__ int cnt_s[FIN_MAX_REG], off_s[FIN_MAX_REG];
    __shared__ int warp_tot[8];
    __shared__ uint32_t s_prefix, s_lo, s_hi;
    __shared__ int s_rank, s_out, s_ovf, s_total;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();             // everything read below comes from the candidate pass that precedes this launch
    pdl_trigger();
    const int n_old = surv_in_cnt ? ld_pdl_s32(surv_in_cnt + q) : 0;
    if (tid == 0) { s_ovf = ld_pdl_s32(overflow + q); s_out = 0; s_lo = 0xFFFFFFFFu; s_hi = 0u; s_prefix = 0; s_rank = K; }
    __syncthreads();
    // ---- region fill counts -> offsets behind the old survivors; counters reset for the next phase ----
    {
        int c = 0;
        if (tid < n_reg) {
            c = ld_pdl_s32(cand_cnt + (size_t)q * n_reg + tid);
            cand_cnt[(size_t)q * n_reg + tid] = 0;
            if (c > slots) s_ovf = 1;
        }
        int v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (warp < 8 && lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (tid < n_reg) {
            int before = n_old;
            for (int w = 0; w < warp; w++) before += warp_tot[w];
            cnt_s[tid] = c;
            off_s[tid] = before + v - c;
            if (tid == n_reg - 1) s_total = before + v;
        }
        if (n_reg == 0 && tid == 0) s_total = n_old;
        __syncthreads();
    }
    const int total = s_total;
    if (staged_max && tid == 0) atomicMax(staged_max, total);
    if (s_ovf || total > stage_cap) {
        if (tid == 0) {
            overflow[q] = 1; g[q] = INFINITY;
            if (surv_out_cnt) surv_out_cnt[q] = 0;
            if (FINISH) out_counts[q] = -1;             // the host entry point redoes this query with the exact scan
        }
        return;
    }
    // ---- stage: old survivors, then the regions (a warp per region, two keys per lane and load) ----
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < n_old; i += NT) {
        uint64_t v = surv_in[(size_t)q * surv_cap + i];
        key_s[i] = v;
        uint32_t h = (uint32_t)(v >> 32); lo = min(lo, h); hi = max(hi, h);
    }
    for (int r = warp; r < n_reg; r += NW) {
        const int cc = cnt_s[r];
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(cand + ((size_t)q * n_reg + r) * slots);
        uint64_t *dst = key_s + off_s[r];
        for (int i = 2 * lane; i < cc; i += 64) {
            const ulonglong2 v = src[i >> 1];
            dst[i] = v.x;
            uint32_t h = (uint32_t)(v.x >> 32); lo = min(lo, h); hi = max(hi, h);
            if (i + 1 < cc) { dst[i + 1] = v.y; h = (uint32_t)(v.y >> 32); lo = min(lo, h); hi = max(hi, h); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
    __syncthreads();
    float bound = INFINITY;
    if (total >= K) {
        // radix select: 8-bit windows counted from the HIGHEST SIGNIFICANT bit of the key range; two windows are
        // enough -- the bound only has to be >= the K-th smallest key, the undecided low bits are rounded UP
        const uint32_t base = s_lo, range = s_hi - s_lo;
        int bits_left = range == 0 ? 0 : 32 - __clz(range);
        const int stop_at = max(0, bits_left - 16);
        while (bits_left > stop_at) {
            const int width = min(8, bits_left);
            const int shift = bits_left - width;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = (shift + width) >= 32 ? 0u : (0xFFFFFFFFu << (shift + width));
            const uint32_t dmask = (1u << width) - 1u;
            for (int i = tid; i < total; i += NT) {
                uint32_t v = (uint32_t)(key_s[i] >> 32) - base;
                if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & dmask], 1);
            }
            __syncthreads();
            if (warp == 0) {
                int h[8], sum = 0;
                const int r = s_rank;
#pragma unroll
                for (int u = 0; u < 8; u++) { h[u] = hist[lane * 8 + u]; sum += h[u]; }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                int before = inc - sum;
                __syncwarp();            // every lane has read s_rank before the one that owns the bucket rewrites it
                if (r > before && r <= inc) {
                    int rr = r - before, bb = 0;
                    for (; bb < 7; bb++) {
                        if (rr <= h[bb]) break;
                        rr -= h[bb];
                    }
                    s_rank = rr;
                    s_prefix = prefix | ((uint32_t)(lane * 8 + bb) << shift);
                }
            }
            __syncthreads();
            bits_left = shift;
        }
        float tau = ordered_to_float(s_prefix + ((1u << bits_left) - 1u) + base);
        // E_q: header of flat_tensor.cu.  X, Dx: max row norm / max bf16 residual norm; nq, dq: the query's.
        float2 qn = q_norms[q];
        float nq = qn.x, dq = qn.y, X = __uint_as_float(max_bits[0]), Dx = __uint_as_float(max_bits[1]);
        float E = 1.001f * (nq * Dx + dq * X + dq * Dx) +
                  1.01f * (float)(dim + 8) * 1.1920929e-07f * (nq * X + 0.5f * (nq + X) * (nq + X));
        bound = tau + 2.0f * E * e_scale;
        bound = bound + fabsf(bound) * 1e-6f;
    }
    if (tid == 0) g[q] = -bound;
    // ---- survivors: every staged key under the new bound (one shared-memory atomic per warp) ----
    const uint32_t bound_hi = float_to_ordered(bound);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint64_t *sdst = FINISH ? surv_s : surv_out + (size_t)q * surv_cap;
    for (int i0 = 0; i0 < total; i0 += NT) {
        const int i = i0 + tid;
        const uint64_t key = i < total ? key_s[i] : 0ull;
        const bool keep = i < total && (uint32_t)(key >> 32) <= bound_hi;
        const uint32_t b = __ballot_sync(0xffffffffu, keep);
        if (b != 0u) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_out, __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = base + __popc(b & lt_mask);
            if (keep && slot < surv_cap) sdst[slot] = key;
        }
    }
    __syncthreads();
    const int m = s_out;
    if (m > surv_cap) {
        if (tid == 0) {
            overflow[q] = 1; g[q] = INFINITY;
            if (surv_out_cnt) surv_out_cnt[q] = 0;
            if (FINISH) out_counts[q] = -1;
        }
        return;
    }
    if (!FINISH) {
        if (tid == 0) surv_out_cnt[q] = m;
        return;
    }

    // =============================== finish: re-score, sort, write ===============================
    if (tid == 0 && rescored && m > 0) atomicAdd(rescored, (unsigned long long)m);
    // The staging area is free now: a ring of FIN_RING tiles + the query.  A tile = FIN_TILE_ROWS rows x ch floats
    // (256-byte pieces when ld % 64 == 0: longer DRAM bursts per gathered row); a round of FIN_ROWS candidates is
    // walked k-chunk by k-chunk, row half by row half, with FIN_RING - 1 tiles in flight behind the one being
    // summed -- the CTA is a gather stream that has to keep ~64 KB outstanding to get its share of HBM.
    const int pcs = ch >> 2;                             // 16-byte pieces per row and step (8 or 16)
    const int pcs_shift = pcs == 16 ? 4 : 3;
    const int row_b = ch * 4, tile_b = FIN_TILE_ROWS * row_b;
    uint8_t *ring = fin_smem;
    float *q_s = reinterpret_cast<float *>(fin_smem + FIN_RING * tile_b);
    for (int j = tid; j < ld; j += NT) q_s[j] = queries[(size_t)q * ld + j];
    const int n_chunks = ld / ch;
    constexpr int HALVES = FIN_ROWS / FIN_TILE_ROWS;
    for (int base = 0; base < m; base += FIN_ROWS) {
        const int nrow = min(FIN_ROWS, m - base);
        const int n_tiles = n_chunks * HALVES;           // tile t: k-chunk t / HALVES, row half t % HALVES
        __syncthreads();            // q_s staged; previous round's readers of the ring are done
        auto issue = [&](int t) {
            if (t < n_tiles) {
                const int c = t / HALVES, r0 = (t % HALVES) * FIN_TILE_ROWS;
                uint8_t *dst = ring + (size_t)(t % FIN_RING) * tile_b;
                const int rows_here = min(FIN_TILE_ROWS, nrow - r0);
                for (int idx = tid; idx < (rows_here << pcs_shift); idx += NT) {
                    const int r = idx >> pcs_shift, piece = idx & (pcs - 1);
                    const uint32_t pos = key_pos(surv_s[base + r0 + r]);
                    const float *src = rows + (size_t)pos * ld + c * ch + piece * 4;
                    const uint32_t d = smem_u32(dst + r * row_b + ((piece ^ (r & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");     // (possibly empty: keeps the group count uniform)
        };
#pragma unroll 1
        for (int t = 0; t < FIN_RING - 1; t++) issue(t);
        float acc = 0.0f;
        const int my_half = tid / FIN_TILE_ROWS, my_r = tid % FIN_TILE_ROWS;     // compute threads: tid < FIN_ROWS
#pragma unroll 1
        for (int t = 0; t < n_tiles; t++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(FIN_RING - 2) : "memory");
            __syncthreads();        // tile t has landed for everybody; everybody is done with tile t - 1
            issue(t + FIN_RING - 1);                                 // into the slot tile t - 1 occupied
            if (tid < FIN_ROWS && my_half == t % HALVES && tid < nrow) {
                const uint8_t *sp = ring + (size_t)(t % FIN_RING) * tile_b + my_r * row_b;
                const float *qc = q_s + (t / HALVES) * ch;
                for (int j = 0; j < pcs; j++) {
                    float4 xv = *reinterpret_cast<const float4 *>(sp + ((j ^ (my_r & 7)) << 4));
                    float4 qv = *reinterpret_cast<const float4 *>(qc + j * 4);
                    acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
                    acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
                    acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
                    acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (tid < nrow) {
            const float dist = metric_finish<METRIC>(acc);
            const uint32_t pos = key_pos(surv_s[base + tid]);
            // flat_index_search.go:265-271: a threshold > 0 drops rows farther than it
            surv_s[base + tid] = (threshold > 0.0f && dist > threshold) ? KEY_INF : make_key(dist, pos);
        }
    }
    __syncthreads();
    // ---- sort the exact (score, scan position) keys, keep the first K ----
    int P = 64;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += NT) surv_s[i] = KEY_INF;
    __syncthreads();
    bitonic_sort_smem(surv_s, P, tid, NT, CtaBarrier());
    const int want = min(K, m);
    int kept = 0;
    for (int i = tid; i < want; i += NT) {
        const uint64_t key = surv_s[i];
        if (key != KEY_INF) {
            const uint32_t pos = key_pos(key);
            const size_t o = (size_t)q * out_stride + i;
            out_ids[o] = row_ids ? row_ids[pos] : pos;
            out_scores[o] = key_score(key);
            if (out_pos) out_pos[o] = pos;
            kept++;
        }
    }
    // count = keys that are not KEY_INF among the first `want` (they sort last)
    kept = __reduce_add_sync(0xffffffffu, kept);
    if (tid == 0) s_out = 0;
    __syncthreads();
    if (lane == 0 && kept) atomicAdd(&s_out, kept);
    __syncthreads();
    if (tid == 0) out_counts[q] = s_out;
}

// ------------------------------------------------------------------------------------------------
// Split tail: the survivors of the last selection are re-scored by RSF_GRID_X small CTAs per query (the row
// gathers of a 512-query batch are spread over 1024 CTAs, three per SM, 32 KB in flight each), and the CTA
// that finishes LAST for a query sorts that query's exact keys and writes its result -- no merge kernel, no
// overflow-marking kernel, no launch boundaries behind the gathers.
// ------------------------------------------------------------------------------------------------
static constexpr int RSF_GRID_X = 2;
template <int METRIC, bool FMA, int CH>
__global__ void __launch_bounds__(128) rescore_finish_kernel(
    const float *__restrict__ rows, int ld, const float *queries, const uint64_t *rs, const int *rs_cnt, int rs_cap,
    float threshold, uint64_t *keys2, int *keys2_cnt, int *done, const int *overflow, int K,
    const uint32_t *__restrict__ row_ids, long long out_stride, uint32_t *__restrict__ out_ids,
    float *__restrict__ out_scores, long long *__restrict__ out_pos, long long *__restrict__ out_counts,
    unsigned long long *rescored) {
    constexpr int PCS = CH / 4;                      // 16-byte pieces per row per step
    constexpr int ROW_B = CH * 4;                    // bytes of a row in a stage
    constexpr int STAGE_B = 128 * ROW_B;
    extern __shared__ __align__(16) uint8_t rsf_smem[];
    float *q_s = reinterpret_cast<float *>(rsf_smem);                    // [ld]
    uint8_t *stage = rsf_smem + (size_t)ld * 4;                          // [2][128 rows][ROW_B], 16-byte pieces XOR-swizzled
    __shared__ uint32_t pos_s[128];
    __shared__ int s_last;
    const int q = blockIdx.y, tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int cnt = min(ld_pdl_s32(rs_cnt + q), rs_cap);
    for (int j = tid; j < ld; j += 128) q_s[j] = queries[(size_t)q * ld + j];
    const int n_chunks = ld / CH;
    for (int base = blockIdx.x * 128; base < cnt; base += gridDim.x * 128) {
        const int mine = base + tid;
        const bool live = mine < cnt;
        const uint32_t pos = key_pos(rs[(size_t)q * rs_cap + (live ? mine : base)]);
        __syncthreads();            // previous chunk's readers of pos_s / stage are done; q_s is staged
        pos_s[tid] = pos;
        __syncthreads();
        auto issue = [&](int c) {
            uint8_t *dst = stage + (size_t)(c & 1) * STAGE_B;
#pragma unroll
            for (int p = 0; p < PCS; p++) {
                int idx = p * 128 + tid;
                int r = idx / PCS, piece = idx % PCS;
                const float *src = rows + (size_t)pos_s[r] * ld + c * CH + piece * 4;
                uint32_t d = smem_u32(dst + r * ROW_B + ((piece ^ (r & 7)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
        float acc = 0.0f;
        for (int c = 0; c < n_chunks; c++) {
            if (c + 1 < n_chunks) {
                issue(c + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            const uint8_t *sp = stage + (size_t)(c & 1) * STAGE_B + tid * ROW_B;
            const float *qc = q_s + c * CH;
#pragma unroll
            for (int j = 0; j < PCS; j++) {
                float4 xv = *reinterpret_cast<const float4 *>(sp + ((j ^ (tid & 7)) << 4));
                float4 qv = *reinterpret_cast<const float4 *>(qc + j * 4);
                acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
                acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
                acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
                acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
            }
            __syncthreads();
        }
        const float dist = metric_finish<METRIC>(acc);
        // flat_index_search.go:265-271: a threshold > 0 drops rows farther than it
        if (live && !(threshold > 0.0f && dist > threshold)) {
            const int slot = atomicAdd(&keys2_cnt[q], 1);
            keys2[(size_t)q * rs_cap + slot] = make_key(dist, pos);
        }
    }
    // ---- the last CTA of this query to get here sorts and writes the result ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&done[q], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (ld_pdl_s32(overflow + q)) {              // a candidate list overflowed in some phase: the caller redoes the query
        if (tid == 0) out_counts[q] = -1;
        return;
    }
    if (tid == 0 && rescored && cnt > 0) atomicAdd(rescored, (unsigned long long)cnt);
    const int m = min(ld_pdl_s32(keys2_cnt + q), rs_cap);
    uint64_t *sort_s = reinterpret_cast<uint64_t *>(rsf_smem);          // the gather stages are free now
    int P = 64;
    while (P < m) P <<= 1;
    __syncthreads();
    for (int i = tid; i < P; i += 128) {
        uint64_t v = KEY_INF;
        if (i < m) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(keys2 + (size_t)q * rs_cap + i) : "memory");
        sort_s[i] = v;
    }
    __syncthreads();
    bitonic_sort_smem(sort_s, P, tid, 128, CtaBarrier());
    const int want = min(K, m);
    for (int i = tid; i < want; i += 128) {
        const uint64_t key = sort_s[i];
        const uint32_t pos = key_pos(key);
        const size_t o = (size_t)q * out_stride + i;
        out_ids[o] = row_ids ? row_ids[pos] : pos;
        out_scores[o] = key_score(key);
        if (out_pos) out_pos[o] = pos;
    }
    if (tid == 0) out_counts[q] = want;
}

template <int METRIC, bool FMA, int CH>
static int launch_rescore_finish_t(const RescoreFinishArgs &a, cudaStream_t st) {
    const size_t smem = std::max((size_t)a.ld * 4 + 2 * 128 * (size_t)CH * 4, (size_t)TS_RS_CAP * 8);
    auto kern = rescore_finish_kernel<METRIC, FMA, CH>;
    CM_TRY(set_dyn_smem((const void *)kern, smem));
    PdlLaunch L(dim3(RSF_GRID_X, (unsigned)a.nq), dim3(128), smem, st);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, kern, a.rows, a.ld, a.queries, a.rs, a.rs_cnt, (int)TS_RS_CAP, a.threshold, a.keys2,
                               a.keys2_cnt, a.done, a.overflow, a.K, a.row_ids, (long long)a.out_stride, a.out_ids,
                               a.out_scores, (long long *)a.out_pos, (long long *)a.out_counts, a.rescored));
    count_launch();
    return CM_OK;
}

int launch_rescore_finish(const RescoreFinishArgs &a, int metric, bool fma, cudaStream_t st) {
    const bool wide = a.ld % 64 == 0;
#define CM_RSF_CASE(M)                                                                                                \
    case M:                                                                                                           \
        if (wide) return fma ? launch_rescore_finish_t<M, true, 64>(a, st) : launch_rescore_finish_t<M, false, 64>(a, st); \
        return fma ? launch_rescore_finish_t<M, true, 32>(a, st) : launch_rescore_finish_t<M, false, 32>(a, st);
    switch (metric) {
        CM_RSF_CASE(CM_L2)
        CM_RSF_CASE(CM_L2SQ)
        CM_RSF_CASE(CM_COSINE)
    default: return fail(CM_ERR_INVALID_ARG, "unknown metric %d", metric);
    }
#undef CM_RSF_CASE
}

size_t ts_select_smem(bool finish) {
    return (size_t)TS_SEL_STAGE_CAP * 8 + (finish ? (size_t)TS_RS_CAP * 8 : 0);
}

template <bool FINISH, int METRIC, bool FMA>
static int launch_ts_select_t(const TsSelectArgs &a, cudaStream_t st) {
    auto kern = ts_select_kernel<FINISH, METRIC, FMA>;
    const size_t smem = ts_select_smem(FINISH);
    CM_TRY(set_dyn_smem((const void *)kern, smem));
    PdlLaunch L(dim3((unsigned)a.nq), dim3(FIN_THREADS), smem, st);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, kern, a.cand, a.cand_cnt, a.n_reg, a.slots, a.K, a.dim, a.q_norms, a.max_bits, a.g,
                               a.overflow, a.surv_in, a.surv_in_cnt, a.surv_out, a.surv_out_cnt, (int)TS_RS_CAP,
                               (int)TS_SEL_STAGE_CAP, a.e_scale, a.staged_max, a.rows, a.ld, a.ch, a.queries,
                               a.threshold, a.row_ids, (long long)a.out_stride, a.out_ids, a.out_scores,
                               (long long *)a.out_pos, (long long *)a.out_counts, a.rescored));
    count_launch();
    return CM_OK;
}

int launch_ts_select(const TsSelectArgs &a, bool finish, int metric, bool fma, cudaStream_t st) {
    if (a.n_reg > FIN_MAX_REG) return fail(CM_ERR_UNSUPPORTED, "%d candidate regions per query (max %d)", a.n_reg, FIN_MAX_REG);
    if (!finish) return launch_ts_select_t<false, CM_L2SQ, false>(a, st);
    if ((size_t)FIN_RING * FIN_TILE_ROWS * a.ch * 4 + (size_t)a.ld * 4 > (size_t)TS_SEL_STAGE_CAP * 8)
        return fail(CM_ERR_UNSUPPORTED, "finish kernel: rows of %d floats do not fit its staging area", a.ld);
#define CM_FIN_CASE(M)                                                                  \
    case M:                                                                             \
        return fma ? launch_ts_select_t<true, M, true>(a, st) : launch_ts_select_t<true, M, false>(a, st);
    switch (metric) {
        CM_FIN_CASE(CM_L2)
        CM_FIN_CASE(CM_L2SQ)
        CM_FIN_CASE(CM_COSINE)
    default: return fail(CM_ERR_INVALID_ARG, "unknown metric %d", metric);
    }
#undef CM_FIN_CASE
}

}  // namespace cm
