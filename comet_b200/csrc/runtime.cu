// runtime.cu -- process-wide state of libcomet_b200: error strings, device checks, rounding mode,
// tensor-map encoding, pinned host memory.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cm {

static thread_local char tl_error[512] = "";
std::atomic<int64_t> g_kernel_launches{0};
static std::atomic<int> g_rounding{CM_ROUND_SEPARATE};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}
int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
    return code;
}

// dynamic shared memory opt-in per (device, kernel): the runtime call is made only when a launch needs more than
// the kernel was last given on that device (a search step launches the same dozen kernels over and over)
static std::mutex g_smem_mu;
static std::map<std::pair<int, const void *>, size_t> g_smem_set;
int set_dyn_smem(const void *func, size_t bytes) {
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_smem_mu);
    size_t &have = g_smem_set[std::make_pair(dev, func)];
    if (bytes > have) {
        CM_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return CM_OK;
}

static std::once_flag g_dev_once;
static int g_dev_status = CM_ERR_CUDA;
static int g_sm_count = 0;
static size_t g_smem_optin = 0;
static char g_dev_msg[256] = "";

static void probe_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(g_dev_msg, sizeof(g_dev_msg),
                 "no CUDA device available (%s); libcomet_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        snprintf(g_dev_msg, sizeof(g_dev_msg), "cudaGetDeviceProperties failed");
        return;
    }
    if (prop.major != 10) {
        snprintf(g_dev_msg, sizeof(g_dev_msg),
                 "device %d is sm_%d%d; libcomet_b200 is built for sm_100a (B200) only", dev, prop.major,
                 prop.minor);
        return;
    }
    g_sm_count = prop.multiProcessorCount;
    g_smem_optin = prop.sharedMemPerBlockOptin;
    g_dev_status = CM_OK;
}

int ensure_device() {
    std::call_once(g_dev_once, probe_device);
    if (g_dev_status != CM_OK) return fail(CM_ERR_CUDA, "%s", g_dev_msg);
    return CM_OK;
}
int sm_count() { return g_sm_count; }
size_t max_smem_optin() { return g_smem_optin; }
int rounding_mode() { return g_rounding.load(); }

// ---- per-kernel-class event timing -----------------------------------------------------------
struct ProfRec { int cls; cudaEvent_t e0, e1; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec *> g_prof_recs;
static double g_prof_ms[CM_PROF_CLASSES];
static int64_t g_prof_n[CM_PROF_CLASSES];

ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), rec(nullptr) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec *r = new ProfRec();
    r->cls = c;
    cudaEventCreate(&r->e0);
    cudaEventCreate(&r->e1);
    cudaEventRecord(r->e0, s);
    rec = r;
}
ProfScope::~ProfScope() {
    if (!rec) return;
    ProfRec *r = (ProfRec *)rec;
    cudaEventRecord(r->e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
}
static void prof_drain() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec *r : g_prof_recs) {
        float ms = 0.0f;
        if (cudaEventSynchronize(r->e1) == cudaSuccess && cudaEventElapsedTime(&ms, r->e0, r->e1) == cudaSuccess &&
            r->cls >= 0 && r->cls < CM_PROF_CLASSES) {
            g_prof_ms[r->cls] += ms;
            g_prof_n[r->cls] += 1;
        }
        cudaEventDestroy(r->e0);
        cudaEventDestroy(r->e1);
        delete r;
    }
    g_prof_recs.clear();
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;
static std::once_flag g_encode_once;

int make_tmap_2d(CUtensorMap *out, CUtensorMapDataType dtype, uint32_t elem_bytes, const void *base,
                 uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner,
                 uint32_t box_outer, CUtensorMapSwizzle swizzle) {
    std::call_once(g_encode_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = (encode_tiled_fn)fn;
    });
    if (!g_encode) return fail(CM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)elem_bytes;
    CUresult r = g_encode(out, dtype, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CM_OK;
}

}  // namespace cm

extern "C" {

int cm_init(const int *device_ids, int n_devices) {
    if (device_ids && n_devices > 0) {
        cudaError_t e = cudaSetDevice(device_ids[0]);
        if (e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "cudaSetDevice(%d): %s", device_ids[0], cudaGetErrorString(e));
    }
    return cm::ensure_device();
}
void cm_shutdown(void) {}
const char *cm_last_error(void) { return cm::tl_error; }
int cm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
int cm_set_rounding(int mode) {
    if (mode != CM_ROUND_SEPARATE && mode != CM_ROUND_FMA) return cm::fail(CM_ERR_INVALID_ARG, "unknown rounding mode %d", mode);
    cm::g_rounding.store(mode);
    return CM_OK;
}
int cm_get_rounding(void) { return cm::g_rounding.load(); }
const char *cm_version(void) { return "comet_b200 0.1 (sm_100a)"; }
int64_t cm_kernel_launches(void) { return cm::g_kernel_launches.load(); }
int cm_profile_enable(int on) {
    cm::prof_drain();
    cm::g_prof_on.store(on ? 1 : 0);
    return CM_OK;
}
int cm_profile_reset(void) {
    cm::prof_drain();
    std::lock_guard<std::mutex> lk(cm::g_prof_mu);
    for (int i = 0; i < CM_PROF_CLASSES; i++) { cm::g_prof_ms[i] = 0.0; cm::g_prof_n[i] = 0; }
    return CM_OK;
}
int cm_profile_get(int kernel_class, double *total_ms, int64_t *launches) {
    if (kernel_class < 0 || kernel_class >= CM_PROF_CLASSES) return cm::fail(CM_ERR_INVALID_ARG, "bad kernel class");
    cm::prof_drain();
    std::lock_guard<std::mutex> lk(cm::g_prof_mu);
    if (total_ms) *total_ms = cm::g_prof_ms[kernel_class];
    if (launches) *launches = cm::g_prof_n[kernel_class];
    return CM_OK;
}
int cm_host_alloc(void **ptr, size_t bytes) {
    CM_TRY(cm::ensure_device());
    CM_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return CM_OK;
}
int cm_host_free(void *ptr) {
    CM_CUDA(cudaFreeHost(ptr));
    return CM_OK;
}

}  // extern "C"
