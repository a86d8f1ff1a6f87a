// tcgen05.cuh -- thin PTX wrappers for the sm_100a tensor-core path: TMEM allocation, UMMA shared
// memory / instruction descriptors, tcgen05.mma / commit / ld, cluster helpers.  Everything here is
// inline PTX; there is no library dependency.
#pragma once

#include "common.cuh"

namespace cm {
namespace tc {

// ---- cluster helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// arrive (count 1) on an mbarrier given by a shared::cluster address (possibly in the peer CTA).
// Relaxed: the only data it orders are TMEM reads, which tcgen05.fence::before_thread_sync covers;
// a .release here costs a MEMBAR that also waits for every outstanding global store of the warp.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_addr(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}

// ---- TMA loads whose completion lands on a barrier given by shared::cluster address -----------
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(uint32_t smem_dst, const CUtensorMap *tmap, int x, int y,
                                               uint32_t mbar_cluster_addr) {
    if (CG == 2) {
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
            "l"(tmap), "r"(x), "r"(y), "r"(mbar_cluster_addr)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
            "l"(tmap), "r"(x), "r"(y), "r"(mbar_cluster_addr)
            : "memory");
    }
}

// ---- TMEM -------------------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {   // one full warp
    if (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {        // same warp as alloc
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- descriptors ------------------------------------------------------------------------------
// K-major operand tile in shared memory, rows of 64 bf16 (128 B) written by TMA with SWIZZLE_128B:
// 8-row groups are 1024 B apart (SBO); LBO is unused for a single swizzle atom along K.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                                // leading byte offset (ignored)
    d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // layout: SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA / commit ------------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// TS form: the A operand (M x 16 bf16, row m on TMEM lane m, two consecutive k per 32-bit column) comes from
// tensor memory -- an operand that stays resident there costs no shared-memory or L2 traffic at all
template <int CG>
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// Warp-collective forms: EVERY lane of the (converged) warp executes the call with warp-uniform operands and one
// elected lane issues.  With the whole warp on the same path ptxas keeps addresses and descriptors in uniform
// registers; under `if (lane == 0)` it moves every operand of every MMA there through an ELECT / R2UR loop
// (about a dozen dependent instructions per MMA -- more than a 32-cycle N = 64 MMA lasts).
__device__ __forceinline__ void mma_bf16_ts_cg2_warp(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Four consecutive K = 16 steps of one 64-element k-atom in ONE elected block: the A operand advances by 8 TMEM
// columns, the K-major SWIZZLE_128B B descriptor by 32 bytes (2 in its 16-byte units) per step.
__device__ __forceinline__ void mma_bf16_ts_atom_cg2_warp(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                                          uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, a2, a3, l1, l2, l3;\n\t.reg .b64 b0, b1, b2, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, %4, %4;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u32 l1, %2, 2;\n\tadd.u32 l2, %2, 4;\n\tadd.u32 l3, %2, 6;\n\t"
        "mov.b64 b0, {%2, %3};\n\tmov.b64 b1, {l1, %3};\n\tmov.b64 b2, {l2, %3};\n\tmov.b64 b3, {l3, %3};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], b0, %4, p;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], b1, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], b2, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], b3, %4, t;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
}
__device__ __forceinline__ void mma_commit_cg2_warp(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar_addr),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2_warp(uint32_t smem_dst, const CUtensorMap *tmap, int x, int y,
                                                     uint32_t mbar_cluster_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];\n\t}" ::"r"(smem_dst),
        "l"(tmap), "r"(x), "r"(y), "r"(mbar_cluster_addr)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_warp(uint32_t addr, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(addr), "r"(bytes)
        : "memory");
}
// all previously issued MMAs of this thread arrive (count 1) on the barrier at this smem offset --
// in CTA-pair mode on the barrier at the same offset in BOTH CTAs of the pair
template <int CG>
__device__ __forceinline__ void mma_commit(uint32_t bar_addr) {
    if (CG == 2) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
    }
}

// ---- TMEM -> registers: 32 lanes x 32 columns of 32 bit; thread t of the warp gets lane base+t ----
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// registers -> TMEM: thread t of the warp writes lane base+t, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc
}  // namespace cm
