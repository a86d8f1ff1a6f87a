// common.cuh -- shared host/device helpers of libcomet_b200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/comet_b200.h"

namespace cm {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
extern std::atomic<int64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

#define CM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::cm::fail(CM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                             \
    } while (0)

#define CM_TRY(call)                 \
    do {                             \
        int _rc = (call);            \
        if (_rc != CM_OK) return _rc; \
    } while (0)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
// leg).  Disabled by default: a scope costs nothing but a relaxed load.
struct ProfScope {
    int cls; cudaStream_t st; void *rec;
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
};

int ensure_device();            // CM_OK when a CUDA device is usable
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when this kernel needs more than it was last given
int set_dyn_smem(const void *func, size_t bytes);
int sm_count();
int rounding_mode();            // CM_ROUND_*
size_t max_smem_optin();

// ---------------------------------------------------------------------------------------------
// reference arithmetic on device (distance.go).  __fmul_rn/__fadd_rn/__fsub_rn are never
// contracted into FMA by nvcc, so these reproduce gc/amd64's separately rounded `sum += d*d`.
// ---------------------------------------------------------------------------------------------
template <bool FMA>
__device__ __forceinline__ float l2_step(float acc, float a, float b) {
    float diff = __fsub_rn(a, b);
    if (FMA) return __fmaf_rn(diff, diff, acc);
    return __fadd_rn(acc, __fmul_rn(diff, diff));
}
template <bool FMA>
__device__ __forceinline__ float dot_step(float acc, float a, float b) {
    if (FMA) return __fmaf_rn(a, b, acc);
    return __fadd_rn(acc, __fmul_rn(a, b));
}
// metric-specific accumulate step: L2 / L2SQ accumulate squared differences, cosine the dot product
template <int METRIC, bool FMA>
__device__ __forceinline__ float metric_step(float acc, float q, float x) {
    if (METRIC == CM_COSINE) return dot_step<FMA>(acc, q, x);
    return l2_step<FMA>(acc, q, x);
}
// finish: distance.go:120 float32(math.Sqrt(float64(sum))) == correctly rounded fp32 sqrt;
// distance.go:208-215 clamp to [-1,1] then 1 - dot.
template <int METRIC>
__device__ __forceinline__ float metric_finish(float acc) {
    if (METRIC == CM_L2) return __fsqrt_rn(acc);
    if (METRIC == CM_COSINE) {
        if (acc > 1.0f) acc = 1.0f; else if (acc < -1.0f) acc = -1.0f;
        return __fsub_rn(1.0f, acc);
    }
    return acc;
}

#ifdef __CUDACC__
// packed fp32x2 helpers (sm_100a FADD2 / FMUL2): lane by lane identical to the scalar round-to-nearest operations
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpk2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2_rn(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// (A packed add behind a packed multiply is NOT usable on this path: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even with explicit .rn and --fmad=false, which would fuse the two roundings the reference performs separately.)
__device__ __forceinline__ uint64_t mul2_rn(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// four metric_step<METRIC, false> on elements j..j+3 of one query: the element-wise part packed (two IEEE round-to-nearest
// operations per instruction), the accumulation four sequential scalar adds -- the reference's order
// (distance.go:114-121, 158-165, 201-216)
template <int METRIC>
__device__ __forceinline__ float metric_step4_unfused(float acc, const float4 &q, const float4 &x) {
    uint64_t a01 = pk2(q.x, q.y), a23 = pk2(q.z, q.w), b01 = pk2(x.x, x.y), b23 = pk2(x.z, x.w);
    if (METRIC != CM_COSINE) {
        a01 = sub2_rn(a01, b01); a23 = sub2_rn(a23, b23);
        b01 = a01; b23 = a23;
    }
    float p0, p1, p2, p3;
    unpk2(mul2_rn(a01, b01), p0, p1);
    unpk2(mul2_rn(a23, b23), p2, p3);
    acc = __fadd_rn(acc, p0);
    acc = __fadd_rn(acc, p1);
    acc = __fadd_rn(acc, p2);
    return __fadd_rn(acc, p3);
}
#endif

// ---------------------------------------------------------------------------------------------
// ordering keys: (score, scan position) ascending == ascending u64.  Scores on this path are
// never -0.0 (sums of squares, 1 - clamp(dot), sqrt), so the sign-flip transform keeps ties equal.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t pos) {
    return ((uint64_t)float_to_ordered(score) << 32) | (uint64_t)pos;
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_float((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_pos(uint64_t k) { return (uint32_t)k; }
static constexpr uint64_t KEY_INF = ~0ull;

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA / proxy fences (PTX, sm_90+ forms valid on sm_100a)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Wait that parks the warp in hardware until the phase completes (suspend-time hint in ns): it wakes
// within ~60 cycles of the arrive and, unlike a short-timeout try_wait loop, issues nothing meanwhile.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(10000000u)
            : "memory");
    } while (!ok);
}
// 2-D tiled TMA load: box lands in smem, completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int x, int y,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy global -> smem (bytes multiple of 16, both sides 16-B aligned)
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Programmatic dependent launch: a kernel launched with PdlLaunch may be scheduled while its predecessor in the
// stream is still running (once every CTA of the predecessor has called pdl_trigger() or exited).  It must call
// pdl_wait() before it touches anything the predecessor reads or writes; pdl_wait() returns when the predecessor
// grid has completed and its writes are visible.  EVERY kernel of a chain launched this way calls pdl_wait()
// (completion is then transitive along the stream).
// What the predecessor wrote must be read with ordinary (coherent) loads AFTER pdl_wait(): a `const __restrict__`
// pointer or __ldg() turns a load into ld.global.nc, which the compiler may hoist above the wait -- it assumes the
// data does not change while the kernel runs, and under PDL it does.  ld_pdl_*() are volatile loads for the first
// reads behind a wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float ld_pdl_f32(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_pdl_s32(const int *p) {
    int v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
#endif

// host: launch configuration with the programmatic-dependent-launch attribute (and optionally a cluster shape)
struct PdlLaunch {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x = 0, bool pdl = true) {
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        int n = 0;
        if (cluster_x > 0) {
            attr[n].id = cudaLaunchAttributeClusterDimension;
            attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
            n++;
        }
        if (pdl) {
            attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[n].val.programmaticStreamSerializationAllowed = 1;
            n++;
        }
        cfg.attrs = attr; cfg.numAttrs = (unsigned)n;
    }
};

// host: encode a 2-D row-major tensor map (driver entry point fetched at run time, no -lcuda)
int make_tmap_2d(CUtensorMap *out, CUtensorMapDataType dtype, uint32_t elem_bytes, const void *base,
                 uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner,
                 uint32_t box_outer, CUtensorMapSwizzle swizzle);

}  // namespace cm
