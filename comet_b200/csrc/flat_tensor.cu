// flat_tensor.cu -- batched flat search on the 5th-gen tensor cores: a bf16 tcgen05 Q x X^T
// candidate pass with a PROVED error bound, then the reference-order fp32 re-score of the few
// surviving candidates, so results stay bit-identical to flatIndexSearch.searchSingleQuery
// (flat_index_search.go:221-294) while the corpus is streamed once per batch instead of once per
// 8 queries.
//
// Why the result is exact.  Let s(q,x) be the reference score (distance.go:114-121/158-165/201-216,
// sequential fp32) and a~(q,x) = h_x - dot_bf16(q,x) the tensor-core key (h_x = |x|^2/2 for L2/L2^2,
// 0 for cosine; smaller = closer).  Write q = q~ + dq, x = x~ + dx with q~, x~ the bf16 roundings;
// the residuals are exactly representable in fp32, so |dq| and |dx| are KNOWN, not estimated.  Then
// |q.x - q~.x~| <= |q~||dx| + |dq||x~| + |dq||dx| (Cauchy-Schwarz), and with the fp32 effects (tensor
// accumulation, the reference's own sequential rounding, clamp / sqrt plateaus) a~ differs from the
// quantity the reference orders by by at most
//     E_q = 1.001 (|q| Dx + |dq| X + |dq| Dx) + 1.01 (d + 8) 2^-23 (|q| X + (|q| + X)^2 / 2),
// X = max_row |x|, Dx = max_row |dx| (both maintained on the device as rows are added).
// If tau is the K-th smallest a~ over ANY subset of the rows, every row of the reference's top-K
// has a~ <= tau + 2 E_q.  The pass therefore runs in up to three phases over disjoint row samples
// (A: everything is a candidate; B, C: only keys under the bound from the previous phases), a final
// selection keeps the keys <= tau_final + 2 E_q, and those rows (a few hundred per query on
// N(0,1) data) are re-scored in reference order and sorted by (score, scan position) -- the same
// order the exact scan produces.  Candidate-list overflow (adversarial ties) is detected and those
// queries are redone by the exact scan: never a silent approximation.
//
// Kernel shape (flat_gemm_kernel): persistent, warp-specialised, one CTA per SM or one CTA PAIR
// (cta_group::2, UMMA M=256) per two SMs.  Per CTA: 128 corpus rows x 256 queries per accumulator,
// two accumulators in TMEM (512 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
// operands arrive by TMA (SWIZZLE_128B, 64 bf16 per row) through an mbarrier ring.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "flat_tensor.cuh"
#include "select.cuh"
#include "tcgen05.cuh"

namespace cm {

static constexpr int GT_ROWS = 128;        // corpus rows per CTA per tile (TMEM lanes)
static constexpr int GT_QBLK = 256;        // queries per accumulator (UMMA N)
static constexpr int GT_BK = 64;           // bf16 per k-step: one 128-byte swizzle atom
static constexpr int GT_THREADS = 384;     // warps 0-7 epilogue, 8 TMEM alloc, 9 idle, 10 TMA, 11 MMA.  The issue arbiter
                                           // favours the highest warp id of a scheduler, so the two single-lane
                                           // driver warps sit above the (often spinning) epilogue warps.
static constexpr int GT_WARP_ALLOC = 8, GT_WARP_TMA = 10, GT_WARP_MMA = 11;
static constexpr int GT_A_BYTES = GT_ROWS * GT_BK * 2;
static constexpr int GT_MAX_NQ = 1024;     // queries per launch (bounds in smem)
static constexpr int CAND_SLOTS = 128;     // candidate slots per (query, CTA, lane quadrant) region
static constexpr int RS_CAP = TS_RS_CAP;   // candidates re-scored per query
static constexpr int RS_GRID_X = 2;        // re-score CTAs per query (each loops over its 128-candidate chunks)
static constexpr int SEL_STAGE_CAP = 12288; // keys (old survivors + new candidates) staged in smem by the select kernel (96 KB)

template <int CG, bool HAS_H>
__global__ void __launch_bounds__(GT_THREADS, 1) flat_gemm_kernel(
    const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_q, GemmPhase phase,
    int n_qblk, int k_blocks, int stages, long long n_rows, const float *__restrict__ row_h,
    const uint8_t *__restrict__ skip, const float *__restrict__ g_bound, int nq_pad, int cand_slots,
    int n_cta_total, uint64_t *__restrict__ cand, int *__restrict__ cand_cnt) {
    constexpr int B_ROWS = GT_QBLK / CG;                 // query rows this CTA stages per k-step
    constexpr int B_BYTES = B_ROWS * GT_BK * 2;
    constexpr int STAGE_BYTES = GT_A_BYTES + B_BYTES;
    constexpr uint32_t IDESC = tc::make_idesc_bf16(GT_ROWS * CG, GT_QBLK);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *stage_base = smem;
    float *g_s = reinterpret_cast<float *>(smem + (size_t)stages * STAGE_BYTES);
    int *cnt_s = reinterpret_cast<int *>(g_s + GT_MAX_NQ);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(cnt_s + 4 * GT_MAX_NQ);
    uint64_t *empty_bar = full_bar + stages;
    uint64_t *tfull_bar = empty_bar + stages;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta_rank = (CG == 2) ? tc::cluster_ctarank() : 0u;
    const int cluster = (CG == 2) ? (int)tc::cluster_id_x() : (int)blockIdx.x;
    const int n_clusters = (CG == 2) ? (int)tc::nclusters_x() : (int)gridDim.x;
    const int n_work = phase.n_tiles * n_qblk;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8 * CG); }
        fence_mbar_init();
    }
    // fill counts live as [query][region] (region = CTA x lane quadrant): this CTA's four are one int4
    for (int i = tid; i < nq_pad; i += GT_THREADS) {
        g_s[i] = g_bound[i];
        const int4 c4 = *reinterpret_cast<const int4 *>(cand_cnt + ((size_t)i * n_cta_total + blockIdx.x) * 4);
        cnt_s[0 * GT_MAX_NQ + i] = c4.x; cnt_s[1 * GT_MAX_NQ + i] = c4.y;
        cnt_s[2 * GT_MAX_NQ + i] = c4.z; cnt_s[3 * GT_MAX_NQ + i] = c4.w;
    }
    if (warp == GT_WARP_ALLOC) tc::tmem_alloc<CG>(smem_u32(tmem_slot), 512);
    tc::fence_before_thread_sync();
    if (CG == 2) tc::cluster_sync_all(); else __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == GT_WARP_TMA) {
        // ===== TMA producer (one lane) =====
        if (lane == 0) {
            prefetch_tmap(&tmap_x);
            prefetch_tmap(&tmap_q);
            // ring position and phase are loop-carried (no division in the issue loop: this one thread has to
            // stay ahead of the tensor pipe, 512 cycles per k-block)
            int s = 0;
            uint32_t ph = 0;
            const uint32_t stage0 = smem_u32(stage_base), full0 = smem_u32(&full_bar[0]);
            const uint32_t full0_leader = (CG == 2) ? tc::mapa(full0, 0) : full0;
            // timing probes (results are garbage): dbg bit 64 / 128 = stop loading the corpus / query operand
            // once the ring has been filled once -- what the mainloop costs with half (or none) of its L2 traffic
            int fills = 0;
            for (int w = cluster; w < n_work; w += n_clusters) {
                int t = phase_tile(phase, w / n_qblk), nb = w % n_qblk;
                int row0 = t * (GT_ROWS * CG) + (int)cta_rank * GT_ROWS;
                int q0 = nb * GT_QBLK + (int)cta_rank * B_ROWS;
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait_parked(&empty_bar[s], ph ^ 1);
                    const uint32_t bar = full0_leader + (uint32_t)s * 8u;
                    const bool ld_a = !(phase.dbg & 64) || fills < stages, ld_b = !(phase.dbg & 128) || fills < stages;
                    if (fills < stages) fills++;
                    if (cta_rank == 0)
                        tc::mbar_arrive_expect_tx_addr(full0 + (uint32_t)s * 8u, ((ld_a ? GT_A_BYTES : 0) + (ld_b ? B_BYTES : 0)) * CG);
                    const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)STAGE_BYTES;
                    if (ld_a) tc::tma_load_2d_cg<CG>(sa, &tmap_x, kb * GT_BK, row0, bar);
                    if (ld_b) tc::tma_load_2d_cg<CG>(sa + GT_A_BYTES, &tmap_q, kb * GT_BK, q0, bar);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == GT_WARP_MMA) {
        // ===== MMA issuer (one lane of the leader CTA) =====
        if (cta_rank == 0 && lane == 0) {
            uint32_t wi = 0;
            int s = 0;
            uint32_t ph = 0;
            const uint32_t stage0 = smem_u32(stage_base), empty0 = smem_u32(&empty_bar[0]);
            for (int w = cluster; w < n_work; w += n_clusters, wi++) {
                uint32_t acc = wi & 1, aph = (wi >> 1) & 1;
                mbar_wait_parked(&tempty_bar[acc], aph ^ 1);
                tc::fence_after_thread_sync();
                uint32_t d_tmem = tmem_base + acc * GT_QBLK;
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait_parked(&full_bar[s], ph);
                    tc::fence_after_thread_sync();
                    const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)STAGE_BYTES;
                    uint64_t adesc = tc::make_smem_desc_sw128(sa);
                    uint64_t bdesc = tc::make_smem_desc_sw128(sa + GT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < GT_BK / 16; k++) {
                        // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in 16-byte units
                        tc::mma_bf16<CG>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC,
                                         (uint32_t)((kb | k) != 0));
                    }
                    tc::mma_commit<CG>(empty0 + (uint32_t)s * 8u);     // frees the smem slot (both CTAs)
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                tc::mma_commit<CG>(smem_u32(&tfull_bar[acc]));         // accumulator ready (both CTAs)
            }
        }
        __syncwarp();
    } else if (warp < 8) {
        // ===== epilogue: 8 warps; thread = one corpus row (TMEM lane) x 128 of the 256 query columns =====
        // Candidates of query q found by lane quadrant `ew` of this CTA go to the private region
        // (q, CTA, ew): its fill count is only ever touched by the warps of that quadrant, one query
        // half each, so a warp ballot hands out the slots -- no atomics, no divergence.
        const int ew = warp & 3, half = warp >> 2;
        uint32_t wi = 0;
        uint32_t tempty0 = smem_u32(&tempty_bar[0]), tempty1 = smem_u32(&tempty_bar[1]);
        if (CG == 2) { tempty0 = tc::mapa(tempty0, 0); tempty1 = tc::mapa(tempty1, 0); }
        uint64_t *my_cand = cand + ((size_t)blockIdx.x * 4 + ew) * cand_slots;   // + q * q_stride
        const size_t q_stride = (size_t)n_cta_total * 4 * cand_slots;
        int *cnt_w = cnt_s + ew * GT_MAX_NQ;
        const uint32_t lt_mask = (1u << lane) - 1u;
        const int row_in_tile = (int)cta_rank * GT_ROWS + ew * 32 + lane;
        // per-row data of the first item; the next item's is fetched while this one is processed
        long long row = -1;
        float hx = 0.0f;
        bool row_ok = false;
        auto fetch_row = [&](int w) {
            if (w < n_work) {
                row = (long long)phase_tile(phase, w / n_qblk) * (GT_ROWS * CG) + row_in_tile;
                row_ok = row < n_rows;
                hx = (HAS_H && row_ok) ? __ldg(row_h + row) : 0.0f;
                if (row_ok && skip != nullptr) row_ok = __ldg(skip + row) == 0;
            }
        };
        fetch_row(cluster);
        for (int w = cluster; w < n_work; w += n_clusters, wi++) {
            const uint32_t acc = wi & 1, aph = (wi >> 1) & 1;
            const int nb = w % n_qblk;
            const uint32_t cur_row = (uint32_t)row;
            const float cur_hx = hx;
            const bool cur_ok = row_ok;
            fetch_row(w + n_clusters);
            const int qbase = nb * GT_QBLK + half * (GT_QBLK / 2);
            const float *g = g_s + qbase;
            if (phase.dbg & 16) { while (!mbar_try_wait(&tfull_bar[acc], aph)) __nanosleep(128); }
            else if (phase.dbg & 32) mbar_wait(&tfull_bar[acc], aph);
            else mbar_wait_parked(&tfull_bar[acc], aph);
            tc::fence_after_thread_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * GT_QBLK + half * (GT_QBLK / 2);
#pragma unroll 1
            for (int cb = 0; cb < GT_QBLK / 64; cb++) {
                if ((phase.dbg & 15) == 1) break;
                uint32_t vv[32];
                tc::tmem_ld_32x32(taddr + cb * 32, vv);
                tc::tmem_ld_wait();
                if ((phase.dbg & 15) == 2) { if ((vv[0] ^ vv[13] ^ vv[31]) == 0x12345678u) cnt_w[0] = 1; continue; }
                // Scan: t = dot - g_q (- h_x); the sign bit of t is shifted into a per-row hit mask --
                // 2 (cosine) or 3 (L2) ops per accumulator value and no branches.  The kernel runs at
                // the board's power cap, so every instruction here costs tensor clocks.
                uint32_t mask = 0;
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                    float4 gg = *reinterpret_cast<const float4 *>(g + cb * 32 + c4 * 4);
                    float t0 = __uint_as_float(vv[c4 * 4 + 0]) - gg.x, t1 = __uint_as_float(vv[c4 * 4 + 1]) - gg.y;
                    float t2 = __uint_as_float(vv[c4 * 4 + 2]) - gg.z, t3 = __uint_as_float(vv[c4 * 4 + 3]) - gg.w;
                    if (HAS_H) { t0 -= cur_hx; t1 -= cur_hx; t2 -= cur_hx; t3 -= cur_hx; }
                    mask = __funnelshift_l(__float_as_uint(t0), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(t1), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(t2), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(t3), mask, 1);
                }
                uint32_t hits = cur_ok ? ~mask : 0u;       // bit (31 - c) set <=> column c is a candidate
                if (phase.dense) {
                    // Phase A: every live (row, query) pair is a candidate.  Column by column: the fill
                    // count of (query, CTA, quadrant) belongs to this warp alone, so a ballot hands out the
                    // slots -- the value is a named register (no pick), no atomics, no divergence.
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const bool h = (hits >> (31 - c)) & 1u;
                        const uint32_t b = __ballot_sync(0xffffffffu, h);
                        if (b != 0u) {
                            const int q = qbase + cb * 32 + c;
                            // only lane 0 ever touches the counter (read here, written below): no lane can
                            // overtake another one's read under independent thread scheduling
                            const int base = __shfl_sync(0xffffffffu, lane == 0 ? cnt_w[q] : 0, 0);
                            const int slot = base + __popc(b & lt_mask);
                            if (h && slot < cand_slots)
                                my_cand[(size_t)q * q_stride + slot] = make_key(cur_hx - __uint_as_float(vv[c]), cur_row);
                            if (lane == 0) cnt_w[q] = base + __popc(b);
                        }
                    }
                } else if (__popc(__ballot_sync(0xffffffffu, hits != 0u)) > 3) {
                    // several lanes hold hits: one pass of this loop serves one hit of EVERY such lane
                    while (hits != 0u) {
                        const int c = __clz(hits);
                        hits &= ~(0x80000000u >> c);
                        const float dot = __uint_as_float(pick32_sel(vv, c));
                        const int q = qbase + cb * 32 + c;
                        const int slot = atomicAdd(&cnt_w[q], 1);   // counter private to this lane quadrant
                        if (slot < cand_slots) my_cand[(size_t)q * q_stride + slot] = make_key(cur_hx - dot, cur_row);
                    }
                } else {
                    while (hits != 0u) {                        // rare; usually one iteration for one or two lanes
                        const int c = __clz(hits);
                        hits &= ~(0x80000000u >> c);
                        const float dot = __uint_as_float(pick32(vv, c));
                        const int q = qbase + cb * 32 + c;
                        const int slot = atomicAdd(&cnt_w[q], 1);   // counter private to this lane quadrant
                        if (slot < cand_slots) my_cand[(size_t)q * q_stride + slot] = make_key(cur_hx - dot, cur_row);
                    }
                }
                __syncwarp();
            }
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(acc ? tempty1 : tempty0);
        }
    }

    // ---- teardown: nobody exits (or frees TMEM) while the peer may still touch this CTA ----
    tc::fence_before_thread_sync();
    if (CG == 2) tc::cluster_sync_all(); else __syncthreads();
    for (int i = tid; i < nq_pad; i += GT_THREADS)
        *reinterpret_cast<int4 *>(cand_cnt + ((size_t)i * n_cta_total + blockIdx.x) * 4) =
            make_int4(cnt_s[0 * GT_MAX_NQ + i], cnt_s[1 * GT_MAX_NQ + i], cnt_s[2 * GT_MAX_NQ + i], cnt_s[3 * GT_MAX_NQ + i]);
    if (warp == GT_WARP_ALLOC) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc<CG>(tmem_base, 512);
    }
}

static size_t gemm_smem_bytes(int cg, int stages) {
    size_t stage = GT_A_BYTES + (size_t)(GT_QBLK / cg) * GT_BK * 2;
    return (size_t)stages * stage + (size_t)GT_MAX_NQ * 4 * 5 + (size_t)(2 * stages + 4) * 8 + 16;
}

template <int CG, bool HAS_H>
static int launch_gemm_t(const CUtensorMap &tx, const CUtensorMap &tq, const GemmPhase &ph, int n_qblk, int k_blocks,
                         int64_t n_rows, const float *row_h, const uint8_t *skip, const float *g_bound, int nq_pad,
                         uint64_t *cand, int *cand_cnt, cudaStream_t st) {
    int stages = CG == 2 ? 6 : 4;
    size_t smem = gemm_smem_bytes(CG, stages);
    auto kern = flat_gemm_kernel<CG, HAS_H>;
    CM_TRY(set_dyn_smem((const void *)kern, smem));
    int n_work = ph.n_tiles * n_qblk;
    if (n_work <= 0) return CM_OK;
    int clusters = std::min(sm_count() / CG, n_work);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CG));
    cfg.blockDim = dim3(GT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(CM_PROF_FLAT_GEMM, st);
    CM_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tq, ph, n_qblk, k_blocks, stages, (long long)n_rows, row_h, skip,
                               g_bound, nq_pad, (int)CAND_SLOTS, sm_count(), cand, cand_cnt));
    count_launch();
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// shadow copies: rows -> bf16 + h_x + max norm; queries -> bf16 + |q|
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}

// one warp per row; src [n][ld_src] fp32 (ld_src % 4 == 0), dst [n][ldb] bf16 zero padded.
// Besides the bf16 copy it produces what the error bound needs, all rounded UP a little:
//   |x| and |x - bf16(x)| (the rounding residual is exactly representable in fp32),
//   per row into out_norms[row] = {|x|, |dx|} (queries) and/or as running maxima in max_bits[0..1] (corpus).
__global__ void to_bf16_rows_kernel(const float *__restrict__ src, long long n, int dim, int ld_src,
                                    __nv_bfloat16 *__restrict__ dst, int ldb, float h_scale,
                                    float *__restrict__ out_h, float2 *__restrict__ out_norms,
                                    unsigned int *__restrict__ max_bits) {
    long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float *s = src + (size_t)row * ld_src;
    uint2 *d = reinterpret_cast<uint2 *>(dst + (size_t)row * ldb);
    float sq = 0.0f, rsq = 0.0f;
    for (int j = lane * 4; j < ldb; j += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j + 3 < dim) v = *reinterpret_cast<const float4 *>(s + j);
        else {
            if (j < dim) v.x = s[j];
            if (j + 1 < dim) v.y = s[j + 1];
            if (j + 2 < dim) v.z = s[j + 2];
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        float2 r0 = __bfloat1622float2(p0), r1 = __bfloat1622float2(p1);
        float ex = v.x - r0.x, ey = v.y - r0.y, ez = v.z - r1.x, ew = v.w - r1.y;
        sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        rsq += ex * ex + ey * ey + ez * ez + ew * ew;
        d[j >> 2] = make_uint2(*reinterpret_cast<uint32_t *>(&p0), *reinterpret_cast<uint32_t *>(&p1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        rsq += __shfl_xor_sync(0xffffffffu, rsq, o);
    }
    if (lane == 0) {
        float nrm = sqrtf(sq) * 1.0001f, rn = sqrtf(rsq) * 1.0001f;
        if (out_h) out_h[row] = h_scale * sq;
        if (out_norms) out_norms[row] = make_float2(nrm, rn);
        if (max_bits) {
            atomicMax(&max_bits[0], __float_as_uint(nrm));
            atomicMax(&max_bits[1], __float_as_uint(rn));
        }
    }
}

static int launch_to_bf16(const float *src, int64_t n, int dim, int ld_src, __nv_bfloat16 *dst, int ldb, float h_scale,
                          float *out_h, float2 *out_norms, unsigned int *max_bits, cudaStream_t st) {
    if (n <= 0) return CM_OK;
    int threads = 256;
    long long blocks = (n * 32 + threads - 1) / threads;
    to_bf16_rows_kernel<<<(unsigned)blocks, threads, 0, st>>>(src, n, dim, ld_src, dst, ldb, h_scale, out_h, out_norms,
                                                              max_bits);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// candidate selection: K-th smallest key (radix select on the ordered score bits) -> next bound
// ------------------------------------------------------------------------------------------------
// block per query.  The query's NEW candidates live in n_reg regions of `slots` keys (one per GEMM CTA
// and lane quadrant); the survivors of the previous phases (keys under the previous bound) live in a
// dense list.  Both are staged once in shared memory (stage_cap keys); the K-th smallest is found by
// a radix select on (key - min), starting at the highest significant byte; the keys under the new
// bound tau_K + 2E become the survivors for the next phase (ping-pong list) and the region counters
// are reset.  After the last phase the survivors ARE the rows to re-score.
// Two launch shapes: <768, 2> stages up to SEL_STAGE_CAP keys (96 KB, two CTAs per SM: 296 CTAs in flight, so
// 512 queries take two rounds) and <512, 4> stages up to SEL_STAGE_CAP_SMALL keys (48 KB, four CTAs per SM: all
// 512 queries in ONE round).  The host picks the small shape for a phase when the keys it staged on the previous
// call (and, for phase A, the rows it is about to emit) fit with margin.
static constexpr int SEL_THREADS = 768;     // >= candidate regions per query (4 x SMs); 2 CTAs per SM at <= 42 registers
static constexpr int SEL_THREADS_SMALL = 512;
static constexpr int SEL_STAGE_CAP_SMALL = 6144;
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) cand_select_kernel(
    const uint64_t *__restrict__ cand, int *__restrict__ cand_cnt, int nq_pad, int n_reg, int slots, int K, int dim,
    const float2 *__restrict__ q_norms, const unsigned int *__restrict__ max_bits, float *__restrict__ g,
    int *__restrict__ overflow, const uint64_t *__restrict__ surv_in, const int *__restrict__ surv_in_cnt,
    uint64_t *__restrict__ surv_out, int *__restrict__ surv_out_cnt, int surv_cap, int stage_cap, float e_scale,
    int *__restrict__ staged_max) {
    constexpr int NW = NT / 32;
    constexpr int NRB = NT >= 768 ? 1 : 2;     // region blocks of NT regions each (n_reg <= NRB * NT, checked on the host)
    extern __shared__ __align__(16) uint8_t sel_smem_raw[];
    uint64_t *key_s = reinterpret_cast<uint64_t *>(sel_smem_raw);                 // [stage_cap]
    __shared__ int hist[256];
    __shared__ int warp_tot[NRB][32];
    __shared__ uint32_t s_prefix, s_lo, s_hi;
    __shared__ int s_rank, s_out, s_ovf, s_total;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_old = surv_in_cnt ? surv_in_cnt[q] : 0;
    if (tid == 0) { s_ovf = overflow[q]; s_out = 0; s_lo = 0xFFFFFFFFu; s_hi = 0u; s_prefix = 0; s_rank = K; }
    // ---- region counts -> offsets behind the old survivors (block scan), counters reset for the next phase ----
    int c[NRB], incl[NRB];
#pragma unroll
    for (int rb = 0; rb < NRB; rb++) {
        const int r = rb * NT + tid;
        c[rb] = 0;
        if (r < n_reg) {
            c[rb] = cand_cnt[(size_t)q * n_reg + r];
            cand_cnt[(size_t)q * n_reg + r] = 0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int rb = 0; rb < NRB; rb++) {
        if (c[rb] > slots) s_ovf = 1;
        int v = c[rb];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        incl[rb] = v;
        if (lane == 31) warp_tot[rb][warp] = v;
    }
    __syncthreads();
    {   // exclusive scan of the warp totals of both region blocks by every warp (shuffle scan)
        int carry = n_old;
#pragma unroll
        for (int rb = 0; rb < NRB; rb++) {
            int wt = lane < NW ? warp_tot[rb][lane] : 0, wi = wt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            incl[rb] += carry + __shfl_sync(0xffffffffu, wi - wt, warp);
            carry += __shfl_sync(0xffffffffu, wi, 31);
        }
        if (tid == 0) s_total = carry;
    }
    __syncthreads();
    const int total = s_total;
    if (staged_max && tid == 0) atomicMax(staged_max, total);
    if (s_ovf || total > stage_cap) {
        if (tid == 0) { overflow[q] = 1; g[q] = INFINITY; surv_out_cnt[q] = 0; }
        return;
    }
    // ---- stage: old survivors (all threads), then one region per thread: its keys are contiguous, so all of a
    // thread's 16-byte loads are in flight together (the regions hold ~15 keys each in phases B and C, 32 in A).
    // The range of the keys (radix base) is folded into the copy.  (Key-by-key 8-byte cp.async, which needs no
    // registers and keeps every key in flight, measured slower in phases B and C: 24.0 / 28.4 us against 21.7 /
    // 25.4 us.)
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < n_old; i += NT) {
        uint64_t v = surv_in[(size_t)q * surv_cap + i];
        key_s[i] = v;
        uint32_t h = (uint32_t)(v >> 32); lo = min(lo, h); hi = max(hi, h);
    }
#pragma unroll
    for (int rb = 0; rb < NRB; rb++) {
        const int r = rb * NT + tid, cc = c[rb];
        if (r < n_reg && cc > 0) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(cand + ((size_t)q * n_reg + r) * slots);
            uint64_t *dst = key_s + (incl[rb] - cc);
            constexpr int VB = MINB >= 4 ? 2 : 4;       // 16-byte loads in flight per thread (register budget)
            for (int i0 = 0; i0 < cc; i0 += 2 * VB) {
                ulonglong2 v[VB];
#pragma unroll
                for (int u = 0; u < VB; u++)
                    if (i0 + 2 * u < cc) v[u] = __ldg(src + (i0 >> 1) + u);
#pragma unroll
                for (int u = 0; u < VB; u++) {
                    const int i = i0 + 2 * u;
                    if (i < cc) { dst[i] = v[u].x; uint32_t h = (uint32_t)(v[u].x >> 32); lo = min(lo, h); hi = max(hi, h); }
                    if (i + 1 < cc) { dst[i + 1] = v[u].y; uint32_t h = (uint32_t)(v[u].y >> 32); lo = min(lo, h); hi = max(hi, h); }
                }
            }
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
        // digits are 8-bit windows counted from the HIGHEST SIGNIFICANT bit of the range, so the first
        // pass already spreads the keys over up to 256 bins (byte-aligned windows would put them all
        // into the one or two bins of the range's top byte and serialise the shared-memory atomics)
        const uint32_t base = s_lo, range = s_hi - s_lo;
        // Two windows (16 significant bits) are enough: the bound only has to be >= the K-th smallest key,
        // so the undecided low bits are rounded UP (all ones) -- at most 2^-16 of the key range looser.
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
                // bin holding rank s_rank: warp scan over 8 bins per lane
                int h[8], sum = 0;
                const int r = s_rank;              // read before the shuffles: they order it against the write below
#pragma unroll
                for (int u = 0; u < 8; u++) { h[u] = hist[lane * 8 + u]; sum += h[u]; }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                int before = inc - sum;
                bool mine = r > before && r <= inc;
                if (mine) {
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
        // E_q: see the header comment.  X, Dx: max row norm / max bf16 residual norm; nq, dq: the query's.
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
            if (keep && slot < surv_cap) surv_out[(size_t)q * surv_cap + slot] = key;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_out > surv_cap) { overflow[q] = 1; surv_out_cnt[q] = 0; g[q] = INFINITY; }
        else surv_out_cnt[q] = s_out;
    }
}

// ------------------------------------------------------------------------------------------------
// exact re-score of the candidates in reference order (distance.go loops), 128 candidates per block
// ------------------------------------------------------------------------------------------------
// CH = floats of a row fetched per step: 32 (one 128-byte line) or 64 (two consecutive lines, when ld % 64 == 0 --
// longer DRAM bursts for the same bytes in flight).
template <int METRIC, bool FMA, int CH>
__global__ void __launch_bounds__(128) rescore_kernel(const float *__restrict__ rows, int ld, const float *queries,
                                                      const uint64_t *rs, const int *rs_cnt,      // (no __restrict__: see pdl_wait())
                                                      int rs_cap, float threshold, uint64_t *out_keys, int *out_cnt) {
    constexpr int PCS = CH / 4;                      // 16-byte pieces per row per step
    constexpr int ROW_B = CH * 4;                    // bytes of a row in a stage
    constexpr int STAGE_B = 128 * ROW_B;
    extern __shared__ __align__(16) uint8_t smem[];
    float *q_s = reinterpret_cast<float *>(smem);                    // [ld]
    uint8_t *stage = smem + (size_t)ld * 4;                          // [2][128 rows][ROW_B], 16-byte pieces XOR-swizzled
    __shared__ uint32_t pos_s[128];
    const int q = blockIdx.y, tid = threadIdx.x;
    pdl_wait();             // no-op unless launched with the programmatic-dependent-launch attribute
    pdl_trigger();
    const int cnt = min(ld_pdl_s32(rs_cnt + q), rs_cap);
    for (int j = tid; j < ld; j += 128) q_s[j] = queries[(size_t)q * ld + j];
    const int n_chunks = ld / CH;
    // a few CTAs per query, each walking its 128-candidate chunks (a (RS_CAP / 128) x nq grid would launch
    // mostly empty CTAs: ~230 candidates per query survive on the headline workload)
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
        float dist = metric_finish<METRIC>(acc);
        if (live && !(threshold > 0.0f && dist > threshold)) {
            int slot = atomicAdd(&out_cnt[q], 1);
            out_keys[(size_t)q * rs_cap + slot] = make_key(dist, pos);
        }
    }
}

template <int METRIC, bool FMA, int CH>
static int launch_rescore_t(const float *rows, int ld, const float *queries, int nq, const uint64_t *rs, const int *rs_cnt,
                            float threshold, uint64_t *out_keys, int *out_cnt, cudaStream_t st, bool pdl) {
    dim3 grid(RS_GRID_X, (unsigned)nq);
    size_t smem = (size_t)ld * 4 + 2 * 128 * (size_t)CH * 4;
    auto kern = rescore_kernel<METRIC, FMA, CH>;
    CM_TRY(set_dyn_smem((const void *)kern, smem));
    PdlLaunch L(grid, dim3(128), smem, st, 0, pdl);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, kern, rows, ld, queries, rs, rs_cnt, (int)RS_CAP, threshold, out_keys, out_cnt));
    return CM_OK;
}

static int launch_rescore(int metric, bool fma, const float *rows, int ld, const float *queries, int nq,
                          const uint64_t *rs, const int *rs_cnt, float threshold, uint64_t *out_keys, int *out_cnt,
                          cudaStream_t st, bool pdl = false) {
    int ch = (ld % 64 == 0) ? 64 : 32;
    if (const char *e = getenv("COMET_B200_RS_CH")) ch = (atoi(e) == 64 && ld % 64 == 0) ? 64 : 32;
    ProfScope prof(CM_PROF_RESCORE, st);
#define CM_RS_CASE(M)                                                                                                  \
    case M:                                                                                                            \
        if (ch == 64) {                                                                                                \
            if (fma) CM_TRY((launch_rescore_t<M, true, 64>(rows, ld, queries, nq, rs, rs_cnt, threshold, out_keys, out_cnt, st, pdl)));   \
            else CM_TRY((launch_rescore_t<M, false, 64>(rows, ld, queries, nq, rs, rs_cnt, threshold, out_keys, out_cnt, st, pdl)));      \
        } else {                                                                                                       \
            if (fma) CM_TRY((launch_rescore_t<M, true, 32>(rows, ld, queries, nq, rs, rs_cnt, threshold, out_keys, out_cnt, st, pdl)));   \
            else CM_TRY((launch_rescore_t<M, false, 32>(rows, ld, queries, nq, rs, rs_cnt, threshold, out_keys, out_cnt, st, pdl)));      \
        }                                                                                                              \
        break;
    switch (metric) {
        CM_RS_CASE(CM_L2)
        CM_RS_CASE(CM_L2SQ)
        CM_RS_CASE(CM_COSINE)
    default: return fail(CM_ERR_INVALID_ARG, "unknown metric %d", metric);
    }
#undef CM_RS_CASE
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

__global__ void init_bounds_kernel(float *__restrict__ g, int nq, int nq_pad, float first) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq_pad) g[q] = q < nq ? first : INFINITY;
}

// Everything a query chunk needs before the first candidate pass, in ONE launch (it used to be three kernels and
// four memsets): Distance.Preprocess in the reference's order (distance.go:269-290; cosine: sequential sum of
// squares by one lane, then x * (1 / norm)), the zero-padded fp32 copy the re-score reads, the bf16 copy + |q|,
// |q - bf16(q)| for the error bound, the first bound, and the zeroed counters of the chunk.
// One warp per query row; rows [nq, nq_pad) are padding (zero bf16 row, bound +inf = never a candidate).
template <bool FMA>
__global__ void __launch_bounds__(128) prep_queries_kernel(
    int metric, const float *__restrict__ src, int nq, int nq_pad, int dim, float *__restrict__ qp, int ld,
    int *__restrict__ zero_flags, __nv_bfloat16 *__restrict__ q16, int ldb, float2 *__restrict__ qn,
    float *__restrict__ g, float g_first, uint4 *__restrict__ zero_a, long long zero_a_n, uint4 *__restrict__ zero_b,
    long long zero_b_n) {
    extern __shared__ float prow[];
    pdl_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {   // counters of this chunk (16-byte words)
        const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = gridDim.x * (long long)blockDim.x;
        for (long long i = t; i < zero_a_n; i += nt) zero_a[i] = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = t; i < zero_b_n; i += nt) zero_b[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int i = blockIdx.x * (blockDim.x >> 5) + w;
    if (i >= nq_pad) return;
    uint2 *d16 = reinterpret_cast<uint2 *>(q16 + (size_t)i * ldb);
    if (i >= nq) {
        for (int j = lane * 4; j < ldb; j += 128) d16[j >> 2] = make_uint2(0u, 0u);
        if (lane == 0) { g[i] = INFINITY; qn[i] = make_float2(0.0f, 0.0f); }
        return;
    }
    float *r = prow + (size_t)w * ldb;
    const float *sp = src + (size_t)i * dim;
    for (int j = lane; j < ldb; j += 32) r[j] = j < dim ? sp[j] : 0.0f;
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
    const bool do_scale = metric == CM_COSINE && !zero;
    float *dp = qp + (size_t)i * ld;
    for (int j = lane; j < ld; j += 32) {
        float v = 0.0f;
        if (j < dim) { v = do_scale ? __fmul_rn(r[j], scale) : r[j]; r[j] = v; }
        dp[j] = v;
    }
    __syncwarp();
    float sq = 0.0f, rsq = 0.0f;
    for (int j = lane * 4; j < ldb; j += 128) {
        const float4 v = *reinterpret_cast<const float4 *>(r + j);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        float2 r0 = __bfloat1622float2(p0), r1 = __bfloat1622float2(p1);
        float ex = v.x - r0.x, ey = v.y - r0.y, ez = v.z - r1.x, ew = v.w - r1.y;
        sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        rsq += ex * ex + ey * ey + ez * ez + ew * ew;
        d16[j >> 2] = make_uint2(*reinterpret_cast<uint32_t *>(&p0), *reinterpret_cast<uint32_t *>(&p1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        rsq += __shfl_xor_sync(0xffffffffu, rsq, o);
    }
    if (lane == 0) {
        qn[i] = make_float2(sqrtf(sq) * 1.0001f, sqrtf(rsq) * 1.0001f);
        g[i] = g_first;
    }
}

// key offsets as the query-resident pass reads them: +inf for rows that must not become candidates
// (soft-deleted / filtered out / beyond the last row)
__global__ void fill_inf_kernel(float *__restrict__ h, long long from, long long to) {
    long long i = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < to) h[i] = INFINITY;
}
__global__ void masked_h_kernel(const float *__restrict__ row_h, const uint8_t *__restrict__ skip, long long n,
                                long long n_pad, float *__restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = (i < n && !skip[i]) ? row_h[i] : INFINITY;
}

// ------------------------------------------------------------------------------------------------
// FlatIndex glue
// ------------------------------------------------------------------------------------------------
void FlatIndex::free_shadow() {
    cudaFree(rows_bf16);
    cudaFree(row_h);
    cudaFree(max_bits);
    rows_bf16 = nullptr; row_h = nullptr; max_bits = nullptr;
    shadow_rows = shadow_cap = 0;
}

int FlatIndex::ensure_shadow(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(shadow_mu);
    if (!staged_dev) {          // statistics words of the tensor path (concurrent searches share them)
        // one 48-byte block: staged_dev[8] then the re-scored counter, so that one kernel can zero both
        CM_CUDA(cudaMalloc(&staged_dev, 48));
        rescored_dev = reinterpret_cast<unsigned long long *>(staged_dev + 8);
        CM_CUDA(cudaHostAlloc(&staged_host, 8 * sizeof(int), cudaHostAllocDefault));
        for (int p = 0; p < 8; p++) staged_host[p] = -1;
    }
    if (rows_bf16 && shadow_rows == n && shadow_cap == cap) return CM_OK;
    ldb = (dim + GT_BK - 1) / GT_BK * GT_BK;
    if (!rows_bf16 || shadow_cap != cap) {
        free_shadow();
        CM_CUDA(cudaMalloc(&rows_bf16, (size_t)cap * ldb * sizeof(__nv_bfloat16)));
        CM_CUDA(cudaMalloc(&row_h, (size_t)(cap + 2 * TS_N) * sizeof(float)));
        CM_CUDA(cudaMalloc(&max_bits, 2 * sizeof(unsigned int)));
        CM_CUDA(cudaMemsetAsync(max_bits, 0, 2 * sizeof(unsigned int), st));
        shadow_cap = cap;
        shadow_rows = 0;
        CM_TRY(make_tmap_2d(&tmap_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows_bf16, (uint64_t)ldb, (uint64_t)cap,
                            (uint64_t)ldb * 2, GT_BK, GT_ROWS, CU_TENSOR_MAP_SWIZZLE_128B));
        CM_TRY(make_tmap_2d(&tmap_bf16_ts, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows_bf16, (uint64_t)ldb, (uint64_t)cap,
                            (uint64_t)ldb * 2, GT_BK, TS_N / 2, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    if (shadow_rows == 0) CM_CUDA(cudaMemsetAsync(max_bits, 0, 2 * sizeof(unsigned int), st));
    float h_scale = metric == CM_COSINE ? 0.0f : 0.5f;
    CM_TRY(launch_to_bf16(rows + (size_t)shadow_rows * ld, n - shadow_rows, dim, ld, rows_bf16 + (size_t)shadow_rows * ldb,
                          ldb, h_scale, row_h + shadow_rows, nullptr, max_bits, st));
    {   // rows behind the last one read as +inf offsets: the query-resident pass walks whole 64-row tiles
        const long long to = cap + 2 * TS_N;
        fill_inf_kernel<<<(unsigned)((to - n + 255) / 256), 256, 0, st>>>(row_h, (long long)n, to);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    CM_CUDA(cudaStreamSynchronize(st));
    shadow_rows = n;
    return CM_OK;
}

// CM_PATH_AUTO.  Measured on 1M x 768 (tools/nq_sweep.py, profiles/): the candidate pass beats the exact scan at EVERY batch
// size -- one query 0.44 against 0.88 ms, 32 queries 0.42 against 3.7 ms per call -- because it streams the bf16 shadow
// (half the bytes) once for up to 256 queries; it needs its shadow copy (+50 % memory, built on first use) and enough
// rows for its sampled bounds.  `sparse` = a filter / deleted set that leaves under half of the rows.
bool FlatIndex::tensor_path_eligible(int64_t nq, int64_t k_eff, bool sparse, float) const {
    return nq >= 1 && k_eff <= 256 && n >= 65536 && !sparse;
}

// q_raw: the caller's queries [nq][dim]; qp: [>= nq][ld] receives their preprocessed, zero-padded copy
// (Distance.Preprocess, flat_index_search.go:236), qflags[i] = 1 for a zero vector under cosine.
int FlatIndex::search_tensor(const float *q_raw, float *qp, int *qflags, int64_t nq, int64_t k_eff, const uint8_t *skip,
                             float threshold, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                             int64_t *out_counts, cudaStream_t st, cm_flat_stats *stats) {
    WsScope ws(st);
    if (k_eff > RS_CAP / 2) return fail(CM_ERR_UNSUPPORTED, "tensor path supports k <= %d", RS_CAP / 2);
    CM_TRY(ensure_shadow(st));
    const int cg = tensor_cta_group;
    // Query-resident candidate pass (flat_gemm_ts.cu) whenever a query row fits in tensor memory beside the
    // accumulators; wider rows (and the single-CTA debugging shape) stage both operands in shared memory.
    bool use_ts = cg == 2 && ldb <= TS_MAX_LDB;
    if (const char *e = getenv("COMET_B200_NO_TS")) if (atoi(e)) use_ts = false;
    const int tile_rows = use_ts ? TS_N : GT_ROWS * cg;
    const int T = (int)((n + tile_rows - 1) / tile_rows);
    const int K = (int)k_eff;
    bool fma = rounding_mode() == CM_ROUND_FMA;

    if (n < 16384) return fail(CM_ERR_UNSUPPORTED, "the tensor path needs at least 16384 rows (have %lld)", (long long)n);
    const int n_cta = sm_count();
    const int n_clusters = n_cta / cg;
    // ---- phase plan (see the header comment) ----
    // Geometric phases: A ~ 50 K rows (all of them candidates, so they have to fit the candidate regions), every
    // later phase ~ 10x the rows seen so far, the last one takes the rest once it is within 25x of the rows already
    // seen.  A phase then emits ~ 2 K x (its rows / rows seen before) keys per query whatever n is (sweep on 1M and
    // 12.5M x 768: profiles/r01_tensor_path.md).  Tile t belongs to the first level whose stride divides t
    // (strides nest: S_0 | S_1 | ... ).
    // Region capacity bounds phase A: shared-memory staging gives every cluster at most one 128/256-row tile
    // (128 slots per lane quadrant); the query-resident pass puts 32 keys per tile into a 256-slot region and a
    // query block is served by at least n_clusters / 4 clusters.
    const int max_tiles_a = use_ts ? (TS_SLOTS / (TS_N / 2) - 2) * std::max(1, n_clusters / (GT_MAX_NQ / TS_QBLK))
                                   : std::max(1, n_clusters / 2);
    GemmPhase ph[MAX_PH];
    int n_ph = 0;
    {
        int64_t rows_a = std::min<int64_t>(std::max<int64_t>(50 * (int64_t)K, 2048), 8192);
        if (const char *e = getenv("COMET_B200_ROWS_A")) rows_a = atoll(e);
        int64_t growth = 10;
        if (const char *e = getenv("COMET_B200_PHASE_GROWTH")) growth = std::max<int64_t>(2, atoll(e));
        int tA = std::min((int)((rows_a + tile_rows - 1) / tile_rows), max_tiles_a);
        for (;;) {
            n_ph = 0;
            // cumulative tile targets of the sampled levels
            std::vector<int64_t> cum;
            cum.push_back(tA);
            while ((int)cum.size() < MAX_PH - 1 && (T - cum.back()) > 25 * cum.back()) cum.push_back(cum.back() * (growth + 1));
            // strides from the finest level up, each a multiple of the next
            int L = (int)cum.size();
            std::vector<int> S((size_t)L);
            S[(size_t)L - 1] = (int)std::max<int64_t>(2, T / cum[(size_t)L - 1]);
            for (int i = L - 2; i >= 0; i--) {
                int R = (int)std::max<int64_t>(2, (cum[(size_t)i + 1] + cum[(size_t)i] / 2) / cum[(size_t)i]);
                S[(size_t)i] = S[(size_t)i + 1] * R;
            }
            auto mult = [&](int stride) { return (int)((T + stride - 1) / stride); };   // multiples of stride in [0, T)
            ph[n_ph++] = GemmPhase{0, S[0], S[0], mult(S[0]), 0, 1};
            for (int i = 1; i < L; i++) ph[n_ph++] = GemmPhase{1, S[(size_t)i - 1], S[(size_t)i], mult(S[(size_t)i]) - mult(S[(size_t)i - 1]), 0, 0};
            ph[n_ph++] = GemmPhase{2, S[(size_t)L - 1], S[(size_t)L - 1], T - mult(S[(size_t)L - 1]), 0, 0};
            const int cap_a = use_ts ? (TS_SLOTS / (TS_N / 2)) * std::max(1, n_clusters / (GT_MAX_NQ / TS_QBLK)) : n_clusters;
            if (ph[0].n_tiles <= cap_a) break;
            if (tA <= 1) return fail(CM_ERR_UNSUPPORTED, "phase plan: %d first-phase tiles for %d clusters", ph[0].n_tiles, n_clusters);
            tA = tA * 3 / 4;     // stride rounding made phase A larger than asked for: ask for less
        }
    }

    // debugging aid: widen the candidate band (>= 1 keeps the result exact)
    float e_scale = 1.0f;
    if (const char *es = getenv("COMET_B200_E_SCALE")) e_scale = std::max(1.0f, (float)atof(es));
    const size_t sel_smem = (size_t)SEL_STAGE_CAP * 8, sel_smem_small = (size_t)SEL_STAGE_CAP_SMALL * 8;
    CM_TRY(set_dyn_smem((const void *)cand_select_kernel<SEL_THREADS, 2>, sel_smem));
    CM_TRY(set_dyn_smem((const void *)cand_select_kernel<SEL_THREADS_SMALL, 4>, sel_smem_small));
    int passes = 0;
    if (!use_ts) CM_TRY(launch_preprocess_rows(metric, fma, q_raw, nq, dim, dim, qp, ld, qflags, st));
    // keys staged per phase: counted on the device, copied to pinned memory after the last phase; the NEXT search
    // reads them (no synchronisation: a stale or missing value only means the roomier launch shape is used)
    if (!use_ts) CM_CUDA(cudaMemsetAsync(staged_dev, 0, 8 * sizeof(int), st));      // allocated by ensure_shadow (under its lock)
    const int small_ok = SEL_STAGE_CAP_SMALL - SEL_STAGE_CAP_SMALL / 12;      // 8 % headroom
    bool sel_small[MAX_PH];
    for (int p = 0; p < n_ph; p++) {
        const int prev = staged_host[p];
        sel_small[p] = prev >= 0 && prev <= small_ok;
        if (p == 0) sel_small[p] = (int64_t)ph[0].n_tiles * tile_rows <= small_ok;   // phase A stages exactly its rows
    }
    // Measured on 512 x 1M x 768 (ncu, profiles/): 33.8 / 19.4 / 25.9 us in the one-round shape against 35.8 / 21.7 /
    // 25.4 us -- the kernel is issue-bound, not round-bound -- so the roomier shape stays the default.
    {
        const char *e = getenv("COMET_B200_SEL_SMALL");
        const bool want_small = e && atoi(e) != 0;
        for (int p = 0; p < n_ph; p++) sel_small[p] = sel_small[p] && want_small;
    }
    const bool dbg_staged = getenv("COMET_B200_DBG_STAGED") != nullptr;
    bool sample_ok = n_ph >= 2;        // first phase only samples a bound (see TsBound) when the batch allows it
    if (const char *e = getenv("COMET_B200_NO_SAMPLE")) if (atoi(e)) sample_ok = false;
    // Tail of the search: 0 = selection, re-score, merge and overflow-marking kernels; 1 = ONE kernel per query for
    // the last selection + re-score + sort (flat_finish.cu); 2 = selection + re-score whose last CTA per query sorts
    // and writes.  Measured on 512 x 1M x 768, K = 100 (step, ms): 0.673 / 0.692 / 0.680 -- a CTA per query
    // serialises its selection, gather rounds and sort (512 such CTAs are 1.7 waves), and the last-CTA sort runs on
    // 128 threads behind the gathers; 1024 small gather CTAs followed by a 13 us merge win.  All selectable
    // (COMET_B200_TAIL).
    int tail_mode = 0;
    if (const char *e = getenv("COMET_B200_TAIL")) tail_mode = std::max(0, std::min(2, atoi(e)));
    const bool fused_finish = tail_mode == 1;
    int *stat_words = staged_dev;      // staged_dev[8] + the re-scored counter: 48 bytes
    if (!use_ts) CM_CUDA(cudaMemsetAsync(rescored_dev, 0, 8, st));
    if (const char *dbg = getenv("COMET_B200_DBG_EPI")) for (int p = 0; p < n_ph; p++) ph[p].dbg = atoi(dbg);
    if (const char *e = getenv("COMET_B200_NO_DENSE")) if (atoi(e)) for (int p = 0; p < n_ph; p++) ph[p].dense = 0;
    // key offsets with the skip mask folded in (query-resident pass: +inf = never a candidate)
    const float *h_eff = row_h;
    float *h_masked = nullptr;
    if (use_ts && skip != nullptr) {
        const long long n_pad = (long long)T * TS_N;
        CM_TRY(ws.get(&h_masked, (size_t)n_pad * 4));
        masked_h_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(row_h, skip, (long long)n, n_pad, h_masked);
        count_launch();
        CM_CUDA(cudaGetLastError());
        h_eff = h_masked;
    }
    for (int64_t q0 = 0; q0 < nq; q0 += GT_MAX_NQ) {
        int nqc = (int)std::min<int64_t>(GT_MAX_NQ, nq - q0);
        int nq_pad = (nqc + GT_QBLK - 1) / GT_QBLK * GT_QBLK;
        int n_qblk = nq_pad / GT_QBLK;
        // candidate regions per query: (CTA, lane quadrant) / (cluster of the query's block, column half)
        const int n_reg = use_ts ? ts_regions(n_clusters, n_qblk) : n_cta * 4;
        const int slots = use_ts ? TS_SLOTS : CAND_SLOTS;
        if (n_reg > SEL_THREADS || n_reg > 2 * SEL_THREADS_SMALL)
            return fail(CM_ERR_UNSUPPORTED, "%d SMs: more candidate regions than the select kernel scans", n_cta);
        // ---- workspace ----
        __nv_bfloat16 *q16 = nullptr;
        float2 *qn = nullptr;
        float *g = nullptr;
        uint64_t *cand = nullptr, *keys2 = nullptr;
        int *ccnt = nullptr, *ovf = nullptr, *rcnt = nullptr, *kcnt = nullptr;
        uint64_t *rs = nullptr;   // survivor lists, ping-pong: [2][nq_pad][RS_CAP]
        CM_TRY(ws.get(&q16, (size_t)nq_pad * ldb * 2));
        CM_TRY(ws.get(&qn, (size_t)nq_pad * 8));
        CM_TRY(ws.get(&g, (size_t)nq_pad * 4));
        CM_TRY(ws.get(&cand, (size_t)nq_pad * n_reg * slots * 8));
        CM_TRY(ws.get(&ccnt, (size_t)nq_pad * (n_reg + 6) * 4));
        ovf = ccnt + (size_t)nq_pad * n_reg; rcnt = ovf + nq_pad; kcnt = rcnt + 2 * nq_pad;   // rcnt: [2][nq_pad]
        CM_TRY(ws.get(&rs, (size_t)2 * nq_pad * RS_CAP * 8));
        CM_TRY(ws.get(&keys2, (size_t)nq_pad * RS_CAP * 8));
        CUtensorMap tq;
        if (use_ts) {
            // one launch: Preprocess + padded fp32 copy + bf16 copy and norms + first bound + zeroed counters.
            // g = -(bound): phase A lets everything through for real queries -- a large FINITE value, because rows
            // that must not become candidates carry +inf offsets and inf - inf would be NaN -- nothing for padding.
            const size_t cnt_bytes = (size_t)nq_pad * (n_reg + 6) * 4;       // multiple of 16: nq_pad % 256 == 0
            const size_t smem = 4 * (size_t)ldb * 4;
            auto kern = fma ? prep_queries_kernel<true> : prep_queries_kernel<false>;
            CM_TRY(set_dyn_smem((const void *)kern, smem));
            PdlLaunch L(dim3((unsigned)((nq_pad + 3) / 4)), dim3(128), smem, st);
            CM_CUDA(cudaLaunchKernelEx(&L.cfg, kern, metric, q_raw + (size_t)q0 * dim, nqc, nq_pad, dim, qp + (size_t)q0 * ld, ld,
                                       qflags + q0, q16, ldb, qn, g, -3.0e38f, (uint4 *)ccnt, (long long)(cnt_bytes / 16),
                                       (uint4 *)stat_words, (long long)(q0 == 0 ? 3 : 0)));
            count_launch();
            CM_TRY(make_tmap_2d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q16, (uint64_t)ldb, (uint64_t)nq_pad,
                                (uint64_t)ldb * 2, GT_BK, TS_QBLK / 2, CU_TENSOR_MAP_SWIZZLE_128B));
        } else {
            CM_CUDA(cudaMemsetAsync(ccnt, 0, (size_t)nq_pad * (n_reg + 6) * 4, st));
            CM_CUDA(cudaMemsetAsync(q16, 0, (size_t)nq_pad * ldb * 2, st));
            init_bounds_kernel<<<(nq_pad + 255) / 256, 256, 0, st>>>(g, nqc, nq_pad, -INFINITY);
            count_launch();
            CM_CUDA(cudaGetLastError());
            CM_TRY(launch_to_bf16(qp + (size_t)q0 * ld, nqc, dim, ld, q16, ldb, 0.0f, nullptr, qn, nullptr, st));
            CM_TRY(make_tmap_2d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q16, (uint64_t)ldb, (uint64_t)nq_pad,
                                (uint64_t)ldb * 2, GT_BK, GT_QBLK / cg, CU_TENSOR_MAP_SWIZZLE_128B));
        }
        // Sampling first phase (query-resident pass): instead of emitting every row of phase A as a candidate and
        // selecting among them, the pass only derives a bound (TsBound) and the next phase covers A's tiles again
        // (+0.5 % of the rows): one launch boundary, one selection and the densest emission less.
        GemmPhase phc[MAX_PH];
        for (int p = 0; p < n_ph; p++) phc[p] = ph[p];
        TsBound tsb{};
        unsigned int *gmax = reinterpret_cast<unsigned int *>(kcnt + nq_pad);      // [nq_pad] words zeroed with the counters
        bool sampled = false;
        if (use_ts && sample_ok) {
            const int step_min = n_clusters / n_qblk, step_max = (n_clusters + n_qblk - 1) / n_qblk;
            const int j = (K + 2 * step_min - 1) / (2 * step_min);
            if (j <= TS_BOUND_J && ph[0].n_tiles >= step_max) {     // every epilogue thread sees at least 32 sample rows
                sampled = true;
                tsb.j = j; tsb.q_norms = qn; tsb.max_bits = max_bits; tsb.dim = dim; tsb.e_scale = e_scale; tsb.gmax_bits = gmax;
                phc[0].dense = 2;
                if (ph[1].cls == 1) {
                    // The sampled bound is the max over ~74 threads of a low order statistic of what each thread saw:
                    // it needs a few hundred rows per thread to get near the quantile an exact selection over phase A
                    // gave (64 rows per thread left the worst queries at the 28 % quantile: 16 K keys in the next
                    // phase).  Every third tile of the next phase: ~290 rows per thread at 1M rows, bound ~2.5 %.
                    const int s_a = 3 * ph[1].SB;
                    phc[0] = GemmPhase{0, s_a, s_a, (T + s_a - 1) / s_a, ph[0].dbg, 2};
                    phc[1] = GemmPhase{0, ph[1].SB, ph[1].SB, ph[1].n_tiles + ph[0].n_tiles, ph[1].dbg, 0};
                } else {
                    phc[1] = GemmPhase{3, 1, 1, T, ph[1].dbg, 0};
                }
            }
        }
        int n_sel = 0;              // selections done: survivor lists ping-pong between rs[0] and rs[1]
        for (int p = 0; p < n_ph; p++) {
            const bool has_h = metric != CM_COSINE;   // cosine keys are -dot: no per-row offset
            if (use_ts) {
                TsBound b = tsb;
                b.read_bits = sampled && p == 1;
                if (!sampled) b.j = 0;
                CM_TRY(launch_gemm_ts(tmap_bf16_ts, tq, phc[p], n_qblk, ldb, h_eff, has_h || h_masked != nullptr, n, g, b, cand, ccnt, st));
            } else if (cg == 2 && has_h)
                CM_TRY((launch_gemm_t<2, true>(tmap_bf16, tq, ph[p], n_qblk, ldb / GT_BK, n, row_h, skip, g, nq_pad, cand, ccnt, st)));
            else if (cg == 2)
                CM_TRY((launch_gemm_t<2, false>(tmap_bf16, tq, ph[p], n_qblk, ldb / GT_BK, n, row_h, skip, g, nq_pad, cand, ccnt, st)));
            else if (has_h)
                CM_TRY((launch_gemm_t<1, true>(tmap_bf16, tq, ph[p], n_qblk, ldb / GT_BK, n, row_h, skip, g, nq_pad, cand, ccnt, st)));
            else
                CM_TRY((launch_gemm_t<1, false>(tmap_bf16, tq, ph[p], n_qblk, ldb / GT_BK, n, row_h, skip, g, nq_pad, cand, ccnt, st)));
            passes++;
            if (sampled && p == 0) continue;          // the sampling phase leaves no candidates to select from
            const int in = (n_sel + 1) & 1, out = n_sel & 1;
            const uint64_t *s_in = n_sel == 0 ? nullptr : rs + (size_t)in * nq_pad * RS_CAP;
            const int *s_in_cnt = n_sel == 0 ? nullptr : rcnt + (size_t)in * nq_pad;
            uint64_t *s_out = rs + (size_t)out * nq_pad * RS_CAP;
            int *s_out_cnt = rcnt + (size_t)out * nq_pad;
            n_sel++;
            if (use_ts) {
                // selection per query; after the last phase the same kernel re-scores, sorts and writes the result
                const bool finish = p == n_ph - 1 && fused_finish;
                ProfScope prof(finish ? CM_PROF_RESCORE : CM_PROF_SELECT, st);
                TsSelectArgs a{};
                a.nq = nqc; a.cand = cand; a.cand_cnt = ccnt; a.n_reg = n_reg; a.slots = slots; a.K = K; a.dim = dim;
                a.q_norms = qn; a.max_bits = max_bits; a.g = g; a.overflow = ovf;
                a.surv_in = s_in; a.surv_in_cnt = s_in_cnt; a.surv_out = s_out; a.surv_out_cnt = s_out_cnt;
                a.e_scale = e_scale; a.staged_max = staged_dev + p;
                a.rows = rows; a.ld = ld; a.ch = (ld % 64 == 0) ? 64 : 32; a.queries = qp + (size_t)q0 * ld;
                a.threshold = threshold; a.row_ids = ids; a.out_stride = out_stride;
                a.out_ids = out_ids + (size_t)q0 * out_stride; a.out_scores = out_scores + (size_t)q0 * out_stride;
                a.out_pos = out_pos ? out_pos + (size_t)q0 * out_stride : nullptr; a.out_counts = out_counts + q0;
                a.rescored = rescored_dev;
                CM_TRY(launch_ts_select(a, finish, metric, fma, st));
            } else {
                ProfScope prof(CM_PROF_SELECT, st);
                if (sel_small[p])
                    cand_select_kernel<SEL_THREADS_SMALL, 4><<<nqc, SEL_THREADS_SMALL, sel_smem_small, st>>>(
                        cand, ccnt, nq_pad, n_reg, slots, K, dim, qn, max_bits, g, ovf, s_in, s_in_cnt, s_out, s_out_cnt,
                        RS_CAP, SEL_STAGE_CAP_SMALL, e_scale, staged_dev + p);
                else
                    cand_select_kernel<SEL_THREADS, 2><<<nqc, SEL_THREADS, sel_smem, st>>>(
                        cand, ccnt, nq_pad, n_reg, slots, K, dim, qn, max_bits, g, ovf, s_in, s_in_cnt, s_out, s_out_cnt,
                        RS_CAP, SEL_STAGE_CAP, e_scale, staged_dev + p);
                count_launch();
                CM_CUDA(cudaGetLastError());
            }
        }
        if (use_ts && tail_mode == 2) {
            const int last = (n_sel - 1) & 1;
            ProfScope prof(CM_PROF_RESCORE, st);
            RescoreFinishArgs a{};
            a.nq = nqc; a.K = K; a.rows = rows; a.ld = ld; a.queries = qp + (size_t)q0 * ld;
            a.rs = rs + (size_t)last * nq_pad * RS_CAP; a.rs_cnt = rcnt + (size_t)last * nq_pad; a.threshold = threshold;
            a.keys2 = keys2; a.keys2_cnt = kcnt; a.done = kcnt + 2 * nq_pad; a.overflow = ovf;
            a.row_ids = ids; a.out_stride = out_stride; a.out_ids = out_ids + (size_t)q0 * out_stride;
            a.out_scores = out_scores + (size_t)q0 * out_stride;
            a.out_pos = out_pos ? out_pos + (size_t)q0 * out_stride : nullptr; a.out_counts = out_counts + q0;
            a.rescored = rescored_dev;
            CM_TRY(launch_rescore_finish(a, metric, fma, st));
        } else if (!use_ts || tail_mode == 0) {
            const int last = (n_sel - 1) & 1;
            CM_TRY(launch_rescore(metric, fma, rows, ld, qp + (size_t)q0 * ld, nqc, rs + (size_t)last * nq_pad * RS_CAP,
                                  rcnt + (size_t)last * nq_pad, threshold, keys2, kcnt, st, use_ts));
            CM_TRY(launch_merge_topk(keys2, kcnt, nqc, 1, RS_CAP, K, ids, out_stride, out_ids + (size_t)q0 * out_stride,
                                     out_scores + (size_t)q0 * out_stride, out_pos ? out_pos + (size_t)q0 * out_stride : nullptr,
                                     out_counts + q0, st, use_ts, ovf, rcnt + (size_t)last * nq_pad, rescored_dev));
        }
    }
    CM_CUDA(cudaMemcpyAsync(staged_host, staged_dev, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (dbg_staged) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[comet_b200] staged keys (max over queries) per phase:");
        for (int p = 0; p < n_ph; p++) fprintf(stderr, " %d%s", staged_host[p], sel_small[p] ? "s" : "");
        fprintf(stderr, "  (caps %d / %d)\n", SEL_STAGE_CAP_SMALL, SEL_STAGE_CAP);
    }
    stats->path_used = CM_PATH_TENSOR;
    stats->passes = passes;
    stats->candidates = -1;      // on the device: cm_flat_last_stats fetches it
    return CM_OK;
}

}  // namespace cm
