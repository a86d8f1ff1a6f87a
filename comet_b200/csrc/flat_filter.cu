// flat_filter.cu -- document filters on the device (WithDocumentIDs, flat_index_search.go:255-263) without a host sort
// and without a mid-pipeline synchronise.
//
//   * the caller's IDs go through a per-thread pinned staging buffer (an event guards its reuse) and are sorted on the
//     device (cub radix sort);
//   * NON-selective filters become the per-row skip mask the scans already take (build_skip_kernel: binary search of
//     every row's ID in the sorted filter);
//   * SELECTIVE filters (few IDs against the index size) never stream the corpus: the index keeps its node IDs sorted
//     with their scan positions (id_sorted / pos_sorted, rebuilt lazily after Add / Flush), every filter ID is looked up
//     there, the live positions come out in ascending order -- candidate number == scan order, so ties keep the
//     reference's (score, position) order -- and gather_scan_topk (ivf.cu) scores exactly those rows with the IVF list
//     scan (rows gathered by position, reference-order distance).
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "flat_index.cuh"
#include "flat_kernels.cuh"

namespace cm {

// ---- pinned staging of the caller's filter IDs -------------------------------------------------------------------
struct FilterStage {
    uint32_t *p = nullptr;
    int64_t cap = 0;
    cudaEvent_t ev = nullptr;
    bool pending = false;
    int device = -1;
    ~FilterStage() {
        if (p) cudaFreeHost(p);
        if (ev) cudaEventDestroy(ev);
    }
};
static thread_local FilterStage tl_filter;

// copies ids[nf] into pinned memory (waiting only for this thread's PREVIOUS filter upload) and enqueues the upload
int upload_filter_ids(const uint32_t *ids, int64_t nf, uint32_t *dst_dev, cudaStream_t st) {
    FilterStage &f = tl_filter;
    int dev = 0;
    CM_CUDA(cudaGetDevice(&dev));
    if (f.pending) { cudaEventSynchronize(f.ev); f.pending = false; }
    if (f.device != dev && f.ev) { cudaEventDestroy(f.ev); f.ev = nullptr; }
    if (!f.ev) { CM_CUDA(cudaEventCreateWithFlags(&f.ev, cudaEventDisableTiming)); f.device = dev; }
    if (nf > f.cap) {
        if (f.p) cudaFreeHost(f.p);
        f.p = nullptr;
        f.cap = std::max<int64_t>(nf + nf / 2, 4096);
        if (cudaHostAlloc((void **)&f.p, (size_t)f.cap * 4, cudaHostAllocDefault) != cudaSuccess) {
            f.cap = 0;
            return fail(CM_ERR_CUDA, "pinned staging for %lld filter IDs failed", (long long)nf);
        }
    }
    memcpy(f.p, ids, (size_t)nf * 4);
    CM_CUDA(cudaMemcpyAsync(dst_dev, f.p, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    CM_CUDA(cudaEventRecord(f.ev, st));
    f.pending = true;
    return CM_OK;
}

// in-place ascending sort of n uint32 keys (workspace from the call's scope)
int sort_u32_device(uint32_t *keys, int64_t n, WsScope &ws, cudaStream_t st) {
    if (n <= 1) return CM_OK;
    uint32_t *alt = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint32_t> db(keys, alt);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, (int64_t)n, 0, 32, st);
    CM_TRY(ws.get(&alt, (size_t)n * 4));
    CM_TRY(ws.get(&tmp, tmp_bytes));
    db = cub::DoubleBuffer<uint32_t>(keys, alt);
    CM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, (int64_t)n, 0, 32, st));
    count_launch(3);
    if (db.Current() != keys) CM_CUDA(cudaMemcpyAsync(keys, db.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return CM_OK;
}

// ---- node IDs sorted with their scan positions ---------------------------------------------------------------------
__global__ void iota_u32_kernel(uint32_t *p, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void max_run_kernel(const uint32_t *__restrict__ sorted, long long n, int *__restrict__ out_max) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || sorted[i - 1] != sorted[i]) {            // start of a run: walk it (runs are short: IDs are nearly unique)
        int len = 1;
        while (i + len < n && sorted[i + len] == sorted[i]) len++;
        if (len > 1) atomicMax(out_max, len);
    }
}

int FlatIndex::ensure_id_sort(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(shadow_mu);
    if (sorted_n == n && id_sorted) return CM_OK;
    if (n > sorted_cap) {
        cudaFree(id_sorted); cudaFree(pos_sorted);
        id_sorted = pos_sorted = nullptr;
        sorted_cap = n + n / 2 + 1024;
        CM_CUDA(cudaMalloc(&id_sorted, (size_t)sorted_cap * 4));
        CM_CUDA(cudaMalloc(&pos_sorted, (size_t)sorted_cap * 4));
    }
    WsScope ws(st);
    uint32_t *iota = nullptr;
    int *max_run = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ids, id_sorted, iota, pos_sorted, (int64_t)n, 0, 32, st);
    CM_TRY(ws.get(&iota, (size_t)n * 4));
    CM_TRY(ws.get(&tmp, tmp_bytes));
    CM_TRY(ws.get(&max_run, sizeof(int)));
    iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(iota, (long long)n);
    // stable radix sort: equal IDs keep ascending positions
    CM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ids, id_sorted, iota, pos_sorted, (int64_t)n, 0, 32, st));
    CM_CUDA(cudaMemsetAsync(max_run, 0, sizeof(int), st));
    max_run_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(id_sorted, (long long)n, max_run);
    count_launch(5);
    int mr = 0;
    CM_CUDA(cudaMemcpyAsync(&mr, max_run, sizeof(int), cudaMemcpyDeviceToHost, st));
    CM_CUDA(cudaStreamSynchronize(st));        // once per Add / Flush generation, not per search
    max_id_run = std::max(1, mr);
    sorted_n = n;
    return CM_OK;
}

// sorted filter IDs -> positions of the live rows that carry one of them (each ID once, every row of a repeated node ID)
__global__ void filter_positions_kernel(const uint32_t *__restrict__ filt, long long nf, const uint32_t *__restrict__ id_sorted,
                                        const uint32_t *__restrict__ pos_sorted, long long n, const uint8_t *__restrict__ deleted,
                                        uint32_t *__restrict__ cand_pos, long long cap, int *__restrict__ cand_cnt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const uint32_t id = filt[i];
    if (i > 0 && filt[i - 1] == id) return;                // the caller listed it twice
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (id_sorted[mid] < id) lo = mid + 1; else hi = mid;
    }
    for (; lo < n && id_sorted[lo] == id; lo++) {
        const uint32_t pos = pos_sorted[lo];
        if (deleted[pos]) continue;
        const int slot = atomicAdd(cand_cnt, 1);
        if (slot < cap) cand_pos[slot] = pos;
    }
}
__global__ void fill_u32_kernel(uint32_t *p, long long n, uint32_t v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// cand_pos[cap] ascending live positions of the filtered rows (padded with 0xFFFFFFFF), *cand_cnt their number
int FlatIndex::filter_candidates(const uint32_t *filt_sorted, int64_t nf, uint32_t *cand_pos, int64_t cap, int *cand_cnt,
                                 WsScope &ws, cudaStream_t st) {
    CM_CUDA(cudaMemsetAsync(cand_cnt, 0, sizeof(int), st));
    fill_u32_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(cand_pos, (long long)cap, 0xFFFFFFFFu);
    filter_positions_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(filt_sorted, (long long)nf, id_sorted, pos_sorted, (long long)n,
                                                                         deleted, cand_pos, (long long)cap, cand_cnt);
    count_launch(2);
    CM_CUDA(cudaGetLastError());
    return sort_u32_device(cand_pos, cap, ws, st);
}

}  // namespace cm
