// flat_gemm_ts.cu -- candidate pass of the batched flat search with the QUERIES RESIDENT IN TENSOR MEMORY.
//
// Same job as flat_gemm_kernel (flat_tensor.cu): keys  h_x - dot_bf16(q, x)  under a per-query bound become
// candidates for the reference-order re-score (flatIndexSearch.searchSingleQuery, flat_index_search.go:221-294).
// What differs is where the operands live.  A CTA pair (cta_group::2, UMMA M = 256) owns ONE block of 256
// queries for the whole launch and keeps it in tensor memory as the A operand of `tcgen05.mma` (TS form):
// 128 queries per CTA, one per TMEM lane, 768 bf16 = 384 columns, beside two 64-column fp32 accumulators.
// Only the corpus streams: 64 rows per work item (32 per CTA, TMA boxes of 32 x 64 bf16, SWIZZLE_128B, four
// k-atoms per 16 KB stage).  Per flop that is half the L2 -> SM bytes and half the shared-memory operand
// reads of the version that stages both operands, and the epilogue changes sides with the operands:
// thread = query (TMEM lane), registers = corpus rows.  The query's bound is ONE register, its candidate
// region and fill count belong to the thread alone -- no shared-memory bounds, no atomics, no ballots.
//
// Work split: cluster c serves query block c % n_qblk and takes every (clusters of that block)-th tile of the
// phase, so the clusters of different query blocks walk the same tiles at the same time and the second
// reader of a tile finds it in L2.
#include <cuda_bf16.h>

#include "flat_tensor.cuh"
#include "tcgen05.cuh"

namespace cm {

static constexpr int TS_THREADS = 384;     // warps 0-7 epilogue, 8 TMEM alloc, 9 idle, 10 TMA, 11 MMA (driver warps on top:
                                           // the issue arbiter favours the highest warp id of a scheduler)
static constexpr int TS_WARP_ALLOC = 8, TS_WARP_TMA = 10, TS_WARP_MMA = 11;
static constexpr int TS_KA = 4;                            // 64-element k-atoms per stage
static constexpr int TS_ATOM_BYTES = (TS_N / 2) * 128;     // 32 rows x 128 B
static constexpr int TS_STAGE_BYTES = TS_KA * TS_ATOM_BYTES;
static constexpr int TS_STAGES = 12;                       // 192 KB of corpus in flight per CTA
static constexpr int TS_ACC_COL0 = TS_MAX_LDB / 2;         // accumulators behind the query columns
static constexpr uint32_t TS_IDESC = tc::make_idesc_bf16(TS_QBLK, TS_N);

// Tiles of a phase class in the order phase_tile() gives them, WITHOUT its divisions: an N = 64 work item lasts
// ~1500 cycles, and every role walks the item list (an integer division is ~100 dependent cycles on this path).
struct TileWalk {
    int t;            // current tile
    int rm, D;        // position inside the current group of D kept tiles (classes 1 and 2)
    int adv_q, adv_r; // step / D, step % D
    int mul, step;    // tile = index * mul (class 1: SB, class 0: SA); items per advance
    int cls;
    __device__ __forceinline__ TileWalk(const GemmPhase &p, int i0, int step_) : step(step_), cls(p.cls) {
        t = phase_tile(p, i0);
        D = p.cls == 1 ? p.SA / p.SB - 1 : (p.cls == 2 ? p.SB - 1 : 1);
        mul = p.cls == 0 ? p.SA : (p.cls == 1 ? p.SB : 1);
        rm = (p.cls == 1 || p.cls == 2) ? i0 % D : 0;
        adv_q = step_ / D;
        adv_r = step_ % D;
    }
    __device__ __forceinline__ void next() {
        if (cls == 1 || cls == 2) {
            // index j = i + i / D + 1 grows by step + (groups crossed)
            int dj = step + adv_q;
            rm += adv_r;
            if (rm >= D) { rm -= D; dj++; }
            t += dj * mul;
        } else {
            t += step * mul;
        }
    }
};

template <bool HAS_H>
__global__ void __launch_bounds__(TS_THREADS, 1) flat_gemm_ts_kernel(
    const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_q, GemmPhase phase, int n_qblk,
    int k_atoms, const float *__restrict__ row_h, long long n_rows,
    const float *g_bound, TsBound tsb, int n_regions, uint64_t *__restrict__ cand, int *__restrict__ cand_cnt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *stage_base = smem;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)TS_STAGES * TS_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + TS_STAGES;
    uint64_t *tfull_bar = empty_bar + TS_STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint64_t *qready_bar = tempty_bar + 2;          // leader CTA: both CTAs' query blocks are in tensor memory
    uint64_t *qload_bar = qready_bar + 1;           // this CTA's query block has landed in the (still idle) stage ring
    uint64_t *qdone_bar = qload_bar + 1;            // this CTA's epilogue warps have read it: the ring is free
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qdone_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // cluster = CTA pair along x: rank and cluster index from blockIdx (ptxas knows they are warp-uniform; the
    // %cluster_ctarank / %clusterid special registers read through `asm volatile` are opaque to it)
    const uint32_t cta_rank = blockIdx.x & 1u;
    const int cluster = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const int nb = cluster % n_qblk;                       // this pair's query block
    const int j0 = cluster / n_qblk;                       // its index among the clusters of that block
    const int step = (n_clusters - nb + n_qblk - 1) / n_qblk;
    const int n_sb = (k_atoms + TS_KA - 1) / TS_KA;        // stages per work item
    const int n_items = j0 < phase.n_tiles ? (phase.n_tiles - j0 + step - 1) / step : 0;

    if (tid == 0) {
        for (int s = 0; s < TS_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }
        mbar_init(qready_bar, 16);
        mbar_init(qload_bar, 1);
        mbar_init(qdone_bar, 8);
        fence_mbar_init();
    }
    if (warp == TS_WARP_ALLOC) tc::tmem_alloc<2>(smem_u32(tmem_slot), 512);
    tc::fence_before_thread_sync();
    tc::cluster_sync_all();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: this grid may have started while the previous kernel of the search (query
    // preparation or a selection) was still running; the next one may be scheduled as soon as SMs free up.  Only
    // the epilogue warps touch what the predecessor produces (bounds) or reads (candidate regions): they wait
    // below, after the query block -- written two or more kernels ago -- is in tensor memory.
    pdl_trigger();

    if (warp == TS_WARP_TMA) {
        // ===== corpus producer: the only stream of the kernel (whole warp walks the loop, one elected lane issues) =====
        prefetch_tmap(&tmap_x);
        // The CTA's 128 query rows first: one 128-row box per k-atom into the idle stage ring (an atom of 128 rows
        // x 128 bytes is exactly one 16 KB stage), from where the epilogue warps move them into tensor memory.
        // First phase: the predecessor IS the kernel that writes the bf16 queries -- wait for it.  Later phases follow
        // a selection kernel that triggers only after ITS wait, i.e. after the previous candidate pass (and,
        // transitively, the query preparation) has completed: the block can be fetched while that selection runs.
        if (phase.cls == 0 || phase.cls == 3) pdl_wait();
        if (lane == 0) {
            prefetch_tmap(&tmap_q);
            mbar_arrive_expect_tx(qload_bar, (uint32_t)k_atoms * TS_STAGE_BYTES);
            const int qrow0 = nb * TS_QBLK + (int)cta_rank * (TS_QBLK / 2);
            for (int a = 0; a < k_atoms; a++) tma_load_2d(stage_base + (size_t)a * TS_STAGE_BYTES, &tmap_q, a * 64, qrow0, qload_bar);
        }
        __syncwarp();
        mbar_wait_parked(qdone_bar, 0);             // the ring belongs to the corpus stream from here on
        int s = 0;
        uint32_t ph = 0;
        const uint32_t stage0 = smem_u32(stage_base), full0 = smem_u32(&full_bar[0]);
        const uint32_t full0_leader = tc::mapa(full0, 0);
        int fills = 0;          // timing probe (dbg bit 64, results are garbage): no loads once the ring was filled
        TileWalk tw(phase, j0, step);
        for (int it = 0; it < n_items; it++, tw.next()) {
            const int row0 = tw.t * TS_N + (int)cta_rank * (TS_N / 2);
            for (int sb = 0; sb < n_sb; sb++) {
                mbar_wait_parked(&empty_bar[s], ph ^ 1);
                const int atoms = ((phase.dbg & 64) && fills >= TS_STAGES) ? 0 : min(TS_KA, k_atoms - sb * TS_KA);
                if (fills < TS_STAGES) fills++;
                const uint32_t bar = full0_leader + (uint32_t)s * 8u;
                if (cta_rank == 0) tc::mbar_arrive_expect_tx_warp(full0 + (uint32_t)s * 8u, (uint32_t)atoms * TS_ATOM_BYTES * 2u);
                const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)TS_STAGE_BYTES;
                for (int a = 0; a < atoms; a++)
                    tc::tma_load_2d_cg2_warp(sa + (uint32_t)a * TS_ATOM_BYTES, &tmap_x, (sb * TS_KA + a) * 64, row0, bar);
                if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == TS_WARP_MMA) {
        // ===== MMA issuer (leader CTA; the whole warp walks the loop, one elected lane issues) =====
        // D[256 q x 64 rows] += Q[tmem] . X[smem]^T; an N = 64 MMA lasts 32 cycles, so the instructions around
        // each one are counted: one elected block per k-atom (4 MMAs), all operands in uniform registers.
        if (cta_rank == 0) {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform
            mbar_wait_parked(qready_bar, 0);               // both CTAs' query blocks are in tensor memory
            tc::fence_after_thread_sync();
            int s = 0;
            uint32_t ph = 0;
            const uint32_t stage0 = smem_u32(stage_base), empty0 = smem_u32(&empty_bar[0]);
            const uint32_t desc_hi = (uint32_t)(tc::make_smem_desc_sw128(0) >> 32);
            for (int it = 0; it < n_items; it++) {
                const uint32_t acc = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
                mbar_wait_parked(&tempty_bar[acc], aph ^ 1);
                tc::fence_after_thread_sync();
                const uint32_t d_tmem = tb + TS_ACC_COL0 + acc * TS_N;
                for (int sb = 0; sb < n_sb; sb++) {
                    mbar_wait_parked(&full_bar[s], ph);
                    tc::fence_after_thread_sync();
                    const int atoms = min(TS_KA, k_atoms - sb * TS_KA);
                    const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)TS_STAGE_BYTES;
                    const uint32_t a_tmem = tb + (uint32_t)(sb * TS_KA) * 32u;     // 64 bf16 = 32 columns per atom
                    if (atoms == TS_KA) {
#pragma unroll
                        for (int a = 0; a < TS_KA; a++)
                            tc::mma_bf16_ts_atom_cg2_warp(d_tmem, a_tmem + (uint32_t)a * 32u,
                                                          ((sa + (uint32_t)a * TS_ATOM_BYTES) & 0x3FFFFu) >> 4, desc_hi, TS_IDESC,
                                                          (uint32_t)((sb | a) != 0));
                    } else {
                        for (int a = 0; a < atoms; a++)
                            tc::mma_bf16_ts_atom_cg2_warp(d_tmem, a_tmem + (uint32_t)a * 32u,
                                                          ((sa + (uint32_t)a * TS_ATOM_BYTES) & 0x3FFFFu) >> 4, desc_hi, TS_IDESC,
                                                          (uint32_t)((sb | a) != 0));
                    }
                    tc::mma_commit_cg2_warp(empty0 + (uint32_t)s * 8u);      // frees the stage in both CTAs
                    if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
                }
                tc::mma_commit_cg2_warp(smem_u32(&tfull_bar[acc]));          // accumulator ready in both CTAs
            }
        }
        __syncwarp();
    } else if (warp < 8) {
        // ===== epilogue warps: thread = one query (TMEM lane) x 32 of the 64 corpus rows of an item =====
        const int ew = warp & 3, half = warp >> 2;
        const int q = nb * TS_QBLK + (int)cta_rank * (TS_QBLK / 2) + ew * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(ew * 32) << 16);
        // ---- the query block goes to tensor memory: column j of lane m = bf16 pair (2j, 2j+1) of query m ----
        // (row m of a k-atom is 128 bytes in the SWIZZLE_128B box: its 16-byte chunk j sits at chunk j ^ (m & 7))
        {
            mbar_wait_parked(qload_bar, 0);
            const int m = ew * 32 + lane;
            for (int a = half; a < k_atoms; a += 2) {      // an atom = 32 columns (64 bf16)
                const uint8_t *rowp = stage_base + (size_t)a * TS_STAGE_BYTES + (size_t)m * 128;
                uint32_t v[32];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(rowp + ((u ^ (m & 7)) << 4));
                    v[4 * u + 0] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
                }
                tc::tmem_st_32x32(lane_addr + (uint32_t)a * 32u, v);
            }
            tc::tmem_st_wait();
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(qdone_bar);
                tc::mbar_arrive_cluster(tc::mapa(smem_u32(qready_bar), 0));
            }
        }
        pdl_wait();
        float gq = ld_pdl_f32(g_bound + q);                // -(bound): a key is a candidate iff (dot - gq) - h >= 0
        const bool padding = gq == INFINITY;               // query slot behind the batch: never a candidate
        if (tsb.read_bits && !padding) {                   // bound from the sampling phase (see TsBound)
            const unsigned int b = (unsigned int)ld_pdl_s32(reinterpret_cast<const int *>(tsb.gmax_bits + q));
            gq = b == 0u ? INFINITY : -ordered_to_float(b);
        }
        float sm[TS_BOUND_J];                              // sampling phase: this thread's smallest keys, ascending
#pragma unroll
        for (int u = 0; u < TS_BOUND_J; u++) sm[u] = INFINITY;
        const int region = 2 * j0 + half;
        uint64_t *my_cand = cand + ((size_t)q * n_regions + region) * TS_SLOTS;
        int cnt = 0;
        const uint32_t tempty0 = tc::mapa(smem_u32(&tempty_bar[0]), 0), tempty1 = tc::mapa(smem_u32(&tempty_bar[1]), 0);
        // key offsets of the item's 32 rows (the same for every lane: broadcast loads), fetched one item ahead.
        // !HAS_H (cosine, nothing masked): offsets are all zero; rows behind the last one are cut from the hit mask.
        float hh[32];
        TileWalk tw(phase, j0, step);
        auto fetch_h = [&](int tile) {
            const float4 *hp = reinterpret_cast<const float4 *>(row_h + (size_t)tile * TS_N + half * 32);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float4 t = __ldg(hp + u);
                hh[4 * u + 0] = t.x; hh[4 * u + 1] = t.y; hh[4 * u + 2] = t.z; hh[4 * u + 3] = t.w;
            }
        };
        if (HAS_H && n_items > 0) fetch_h(tw.t);
        for (int it = 0; it < n_items; it++) {
            const uint32_t acc = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
            const uint32_t row0 = (uint32_t)tw.t * TS_N + (uint32_t)half * 32u;
            tw.next();
            mbar_wait_parked(&tfull_bar[acc], aph);
            tc::fence_after_thread_sync();
            uint32_t vv[32];
            tc::tmem_ld_32x32(lane_addr + TS_ACC_COL0 + acc * TS_N + (uint32_t)half * 32u, vv);
            tc::tmem_ld_wait();
            // the accumulator is in registers: hand it back before looking at the values
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(acc ? tempty1 : tempty0);
            if ((phase.dbg & 15) == 1) continue;
            // rows of this half-tile that exist (column 0 is bit 31 of the masks below)
            uint32_t live = 0xFFFFFFFFu;
            if (!HAS_H) {
                const long long left = n_rows - (long long)row0;
                live = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ~(0xFFFFFFFFu >> (int)left));
            }
            if (phase.dense == 2) {
                // sampling phase: no candidates, only the thread's TS_BOUND_J smallest keys (insertions are rare
                // after the first few dozen values: one compare per value on the common path)
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float key = (HAS_H ? hh[c] : 0.0f) - __uint_as_float(vv[c]);
                    if (key < sm[TS_BOUND_J - 1] && ((live >> (31 - c)) & 1u)) {
                        sm[TS_BOUND_J - 1] = key;
#pragma unroll
                        for (int u = TS_BOUND_J - 1; u > 0; u--)
                            if (sm[u] < sm[u - 1]) { const float t = sm[u]; sm[u] = sm[u - 1]; sm[u - 1] = t; }
                    }
                }
            } else if (phase.dense) {
                // phase A: (nearly) every value is a candidate -- straight-line predicated appends
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float h = HAS_H ? hh[c] : 0.0f;
                    const float t = (__uint_as_float(vv[c]) - gq) - h;
                    if (t >= 0.0f && ((live >> (31 - c)) & 1u)) {
                        if (cnt < TS_SLOTS) my_cand[cnt] = make_key(h - __uint_as_float(vv[c]), row0 + c);
                        cnt++;
                    }
                }
            } else {
                // sign bits of t through four independent funnel-shift chains (one chain of 32 is ~150 dependent cycles)
                // (the subtractions two values at a time: FADD2 -- every instruction of this loop is paid ~400 times
                // per CTA and microsecond)
                uint32_t m[4] = {0u, 0u, 0u, 0u};
                const uint64_t gq2 = pk2(gq, gq);
#pragma unroll
                for (int c8 = 0; c8 < 8; c8 += 2) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = k * 8 + c8;
                        uint64_t t2 = sub2_rn(pk2(__uint_as_float(vv[c]), __uint_as_float(vv[c + 1])), gq2);
                        if (HAS_H) t2 = sub2_rn(t2, pk2(hh[c], hh[c + 1]));
                        float t0, t1;
                        unpk2(t2, t0, t1);
                        m[k] = __funnelshift_l(__float_as_uint(t0), m[k], 1);
                        m[k] = __funnelshift_l(__float_as_uint(t1), m[k], 1);
                    }
                }
                const uint32_t mask = (m[0] << 24) | ((m[1] & 0xFFu) << 16) | ((m[2] & 0xFFu) << 8) | (m[3] & 0xFFu);
                uint32_t hits = ~mask & live;
                if (__any_sync(0xffffffffu, hits != 0u)) {
                    if (__popc(__ballot_sync(0xffffffffu, hits != 0u)) > 3) {
                        while (hits != 0u) {               // one pass serves one hit of every lane that has one
                            const int c = __clz(hits);
                            hits &= ~(0x80000000u >> c);
                            const float dot = __uint_as_float(pick32_sel(vv, c));
                            const float h = HAS_H ? __ldg(row_h + row0 + c) : 0.0f;
                            if (cnt < TS_SLOTS) my_cand[cnt] = make_key(h - dot, row0 + c);
                            cnt++;
                        }
                    } else {
                        while (hits != 0u) {               // usually one lane, one hit
                            const int c = __clz(hits);
                            hits &= ~(0x80000000u >> c);
                            const float dot = __uint_as_float(pick32(vv, c));
                            const float h = HAS_H ? __ldg(row_h + row0 + c) : 0.0f;
                            if (cnt < TS_SLOTS) my_cand[cnt] = make_key(h - dot, row0 + c);
                            cnt++;
                        }
                    }
                    __syncwarp();
                }
            }
            if (HAS_H && it + 1 < n_items) fetch_h(tw.t);
        }
        if (phase.dense == 2) {
            if (!padding) {
                const float kth = tsb.j <= 1 ? sm[0] : (tsb.j == 2 ? sm[1] : (tsb.j == 3 ? sm[2] : sm[3]));
                const float2 qn = tsb.q_norms[q];
                const float nq = qn.x, dq = qn.y, X = __uint_as_float(tsb.max_bits[0]), Dx = __uint_as_float(tsb.max_bits[1]);
                const float E = 1.001f * (nq * Dx + dq * X + dq * Dx) +
                                1.01f * (float)(tsb.dim + 8) * 1.1920929e-07f * (nq * X + 0.5f * (nq + X) * (nq + X));
                float bound = kth + 2.0f * E * tsb.e_scale;
                bound = bound + fabsf(bound) * 1e-6f;
                atomicMax(tsb.gmax_bits + q, float_to_ordered(bound));
            }
        } else {
            cand_cnt[(size_t)q * n_regions + region] = cnt;    // > TS_SLOTS: the select kernel flags the overflow
        }
    }

    // ---- teardown: nobody exits (or frees TMEM) while the peer may still touch this CTA ----
    tc::fence_before_thread_sync();
    tc::cluster_sync_all();
    if (warp == TS_WARP_ALLOC) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc<2>(tmem_base, 512);
    }
}

int launch_gemm_ts(const CUtensorMap &tmap_x32, const CUtensorMap &tmap_q128, const GemmPhase &ph, int n_qblk, int ldb,
                   const float *row_h, bool has_h, int64_t n_rows, const float *g_bound, const TsBound &tsb, uint64_t *cand,
                   int *cand_cnt, cudaStream_t st) {
    if (ldb % 64 != 0 || ldb > TS_MAX_LDB) return fail(CM_ERR_UNSUPPORTED, "query-resident pass: ldb %d", ldb);
    const size_t smem = (size_t)TS_STAGES * TS_STAGE_BYTES + (size_t)(2 * TS_STAGES + 7) * 8 + 16;
    auto kern = has_h ? flat_gemm_ts_kernel<true> : flat_gemm_ts_kernel<false>;
    CM_TRY(set_dyn_smem((const void *)kern, smem));
    const int n_clusters = sm_count() / 2;
    PdlLaunch L(dim3((unsigned)(n_clusters * 2)), dim3(TS_THREADS), smem, st, 2);
    ProfScope prof(CM_PROF_FLAT_GEMM, st);
    CM_CUDA(cudaLaunchKernelEx(&L.cfg, kern, tmap_x32, tmap_q128, ph, n_qblk, ldb / 64, row_h,
                               (long long)n_rows, g_bound, tsb, ts_regions(n_clusters, n_qblk), cand, cand_cnt));
    count_launch();
    return CM_OK;
}

}  // namespace cm
