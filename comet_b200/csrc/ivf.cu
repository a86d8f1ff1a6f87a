// ivf.cu -- IVFIndex device state, search and the cm_ivf_* entry points of the C ABI.
//
// Replaces ivfIndexSearch.searchSingleQuery (ivf_index_search.go:217-322): coarse scan of the nlist
// centroids + full sort (K5), exact scan of the vectors of the nprobes nearest lists (K6), full sort
// of the candidates, top k -- with scores, ranks and ids bit-identical to the reference.
//
// Device layout (replaces lists [][]VectorNode, ivf_index.go:82-119):
//   coarse   FlatIndex of the nlist centroids, rows stored RAW (k-means output is never normalised,
//            clustering.go:213-239), scan position == list index;
//   store    FlatIndex of every added vector in ARRIVAL order (preprocessed like IVFIndex.Add,
//            ivf_index.go:269-277), with ids and the soft-delete mask;
//   members  u32 [n]        store positions grouped by list, insertion order inside a list (CSR);
//   list_off i64 [nlist+1]
// A search never moves rows: list scans gather 3 KB rows by position with 16-byte cp.async pieces.
//
// Pipeline per batch of queries: preprocess -> exact coarse scan (flat_scan_kernel, k = nprobes) ->
// ivf_offsets_kernel (per query prefix sums of the probed list lengths = the candidate numbering of
// the reference's append loop) -> ivf_scan_kernel (128 candidates per CTA, reference-order distance,
// delete / document filter / threshold) -> merge_topk_kernel ordered by (score, candidate number) ->
// ivf_emit_kernel (candidate number -> store position -> id).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "kmeans.cuh"
#include "select.cuh"
#include "list_shards.cuh"
#include "wire.cuh"

namespace cm {

struct IVFIndex {
    int dim = 0, nlist = 0, metric = 0, device = 0;
    bool trained = false;
    FlatIndex coarse, store;
    std::vector<std::vector<uint32_t>> lists;   // store positions per list, insertion order
    std::vector<int32_t> list_of;               // list of every store position
    uint32_t *members = nullptr;
    long long *list_off = nullptr;
    int64_t members_cap = 0;
    bool csr_dirty = true;
    std::vector<int64_t> sizes_desc;            // list lengths, descending (bound on candidates per query)
    std::mutex csr_mu;
    unsigned long long *scanned_total = nullptr;   // device: candidates scanned by the last search (all queries)

    ~IVFIndex() { cudaFree(members); cudaFree(list_off); cudaFree(scanned_total); }
    int sync_csr(cudaStream_t st);
    int64_t candidate_bound(int nprobes) const;
};

int IVFIndex::sync_csr(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(csr_mu);
    if (!csr_dirty) return CM_OK;
    int64_t n = store.n;
    if (n > members_cap || !members) {
        cudaFree(members);
        members_cap = std::max<int64_t>(n + n / 2, 1024);
        CM_CUDA(cudaMalloc(&members, (size_t)members_cap * sizeof(uint32_t)));
    }
    if (!list_off) CM_CUDA(cudaMalloc(&list_off, (size_t)(nlist + 1) * sizeof(long long)));
    std::vector<uint32_t> flat;
    flat.reserve((size_t)n);
    std::vector<long long> off((size_t)nlist + 1, 0);
    sizes_desc.assign((size_t)nlist, 0);
    for (int l = 0; l < nlist; l++) {
        off[(size_t)l] = (long long)flat.size();
        flat.insert(flat.end(), lists[(size_t)l].begin(), lists[(size_t)l].end());
        sizes_desc[(size_t)l] = (int64_t)lists[(size_t)l].size();
    }
    off[(size_t)nlist] = (long long)flat.size();
    std::sort(sizes_desc.begin(), sizes_desc.end(), std::greater<int64_t>());
    if (!flat.empty()) CM_CUDA(cudaMemcpyAsync(members, flat.data(), flat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CM_CUDA(cudaMemcpyAsync(list_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    CM_CUDA(cudaStreamSynchronize(st));
    csr_dirty = false;
    return CM_OK;
}

int64_t IVFIndex::candidate_bound(int nprobes) const {
    int64_t b = 0;
    for (int i = 0; i < nprobes && i < (int)sizes_desc.size(); i++) b += sizes_desc[(size_t)i];
    return b;
}

// per query: q_off[q][p] = number of candidates contributed by probes 0..p-1 (p = 0..nprobes)
__global__ void ivf_offsets_kernel(const long long *__restrict__ probe_list, const long long *__restrict__ probe_cnt,
                                   const long long *__restrict__ list_off, int nprobes, long long *__restrict__ q_off,
                                   unsigned long long *__restrict__ total) {
    const int q = blockIdx.x;
    __shared__ long long carry;
    __shared__ long long wsum[8];
    if (threadIdx.x == 0) { carry = 0; q_off[(size_t)q * (nprobes + 1)] = 0; }
    __syncthreads();
    const int np = (int)min((long long)nprobes, probe_cnt[q]);
    for (int p0 = 0; p0 < nprobes; p0 += 256) {
        int p = p0 + threadIdx.x;
        long long len = 0;
        if (p < np) {
            long long l = probe_list[(size_t)q * nprobes + p];
            len = list_off[l + 1] - list_off[l];
        }
        long long inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        long long base = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += wsum[w];
        if (p < nprobes) q_off[(size_t)q * (nprobes + 1) + p + 1] = base + inc;
        __syncthreads();
        if (threadIdx.x == 255) carry = base + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) atomicAdd(total, (unsigned long long)carry);
}

int launch_ivf_offsets(const long long *probe_list, const long long *probe_cnt, const long long *list_off, int nprobes,
                       int64_t nq, long long *q_off, unsigned long long *total, cudaStream_t st) {
    ivf_offsets_kernel<<<(unsigned)nq, 256, 0, st>>>(probe_list, probe_cnt, list_off, nprobes, q_off, total);
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// candidate number c of query q -> store position (largest p with q_off[p] <= c)
__device__ __forceinline__ uint32_t ivf_locate(const long long *__restrict__ qo, int nprobes,
                                               const long long *__restrict__ probe_list,
                                               const long long *__restrict__ list_off,
                                               const uint32_t *__restrict__ members, long long c) {
    int lo = 0, hi = nprobes;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (qo[mid] <= c) lo = mid; else hi = mid;
    }
    long long l = probe_list[lo];
    return members[list_off[l] + (c - qo[lo])];
}

// 128 candidates of one query per CTA: gather rows, reference-order distance, filters, keys out
// CH = floats of a row fetched per step: 32, or 64 (two consecutive 128-byte lines) when ld % 64 == 0
template <int METRIC, bool FMA, int CH>
__global__ void __launch_bounds__(128) ivf_scan_kernel(const float *__restrict__ rows, int ld, const float *__restrict__ queries,
                                                       const long long *__restrict__ probe_list, const long long *__restrict__ q_off,
                                                       const long long *__restrict__ list_off, const uint32_t *__restrict__ members,
                                                       int nprobes, const uint8_t *__restrict__ skip, float threshold,
                                                       long long cap_c, int n_chunks, uint64_t *__restrict__ out_keys,
                                                       int *__restrict__ out_cnt) {
    extern __shared__ __align__(16) uint8_t smem[];
    float *q_s = reinterpret_cast<float *>(smem);
    uint8_t *stage = smem + (size_t)ld * 4;
    __shared__ uint32_t pos_s[128];
    __shared__ int s_cnt;
    const int q = blockIdx.y, tid = threadIdx.x;
    const long long *qo = q_off + (size_t)q * (nprobes + 1);
    const long long cq = qo[nprobes];
    const long long base = (long long)blockIdx.x * 128;
    if (base >= cq) return;
    const long long mine = base + tid;
    const bool live = mine < cq;
    const uint32_t pos = ivf_locate(qo, nprobes, probe_list + (size_t)q * nprobes, list_off, members, live ? mine : base);
    pos_s[tid] = pos;
    if (tid == 0) s_cnt = 0;
    for (int j = tid; j < ld; j += 128) q_s[j] = queries[(size_t)q * ld + j];
    __syncthreads();
    constexpr int PCS = CH / 4, ROW_B = CH * 4, STAGE_B = 128 * ROW_B;
    const int n_ch = ld / CH;
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
    for (int c = 0; c < n_ch; c++) {
        if (c + 1 < n_ch) {
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
    bool pass = live && !(skip != nullptr && skip[pos]) && !(threshold > 0.0f && dist > threshold);
    if (pass) {
        int slot = atomicAdd(&s_cnt, 1);
        out_keys[(size_t)q * cap_c + base + slot] = make_key(dist, (uint32_t)mine);
    }
    __syncthreads();
    if (tid == 0) out_cnt[(size_t)q * n_chunks + blockIdx.x] = s_cnt;
}

// ------------------------------------------------------------------------------------------------
// List-major scan (batches).  The query-major kernel above streams every probed list once per query that
// probes it (16 times at 512 queries x 32 probes over 1024 lists) and sits on the DRAM roofline doing so.
// Here the (query, probe) pairs are grouped by list and a CTA takes 128 rows of a list for up to IVF_QB
// pairs at once: the rows are gathered once per block of pairs (the blocks of a list follow each other
// in the grid, so the second one finds them in L2), each thread walks its row for all the block's
// queries in the reference's order.  A pair's candidates keep their numbers q_off[query][probe] + j
// (ivf_index_search.go:252-268, the tie-break key); its keys go to the parts
// tile_off[query][probe] + tile of the query, 128 slots each, which merge_topk_kernel reads as before.
// ------------------------------------------------------------------------------------------------
static constexpr int IVF_QB = 8;

// per query: tile_off[q][p] = number of 128-row tiles of probes 0..p-1; per list: pairs probing it
__global__ void ivf_group_count_kernel(const long long *__restrict__ probe_list, const long long *__restrict__ probe_cnt,
                                       const long long *__restrict__ list_off, int nprobes, int nq,
                                       long long *__restrict__ tile_off, int *__restrict__ list_cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int np = (int)min((long long)nprobes, probe_cnt[q]);       // probes the coarse step returned (ivf_offsets_kernel)
    long long acc = 0;
    for (int p = 0; p < nprobes; p++) {
        tile_off[(size_t)q * (nprobes + 1) + p] = acc;
        if (p >= np) continue;
        const long long l = probe_list[(size_t)q * nprobes + p];
        acc += (list_off[l + 1] - list_off[l] + 127) >> 7;
        atomicAdd(&list_cnt[l], 1);
    }
    tile_off[(size_t)q * (nprobes + 1) + nprobes] = acc;
}

// one CTA: pair_off[l] = pairs of lists < l, blk_off[l] = blocks of IVF_QB pairs of lists < l; cursors start at pair_off
__global__ void ivf_group_scan_kernel(const int *__restrict__ list_cnt, int nlist, int *__restrict__ pair_off,
                                      int *__restrict__ blk_off, int *__restrict__ cursor) {
    __shared__ int carry_p, carry_b;
    __shared__ int wp[32], wb[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_p = 0; carry_b = 0; }
    __syncthreads();
    for (int l0 = 0; l0 <= nlist; l0 += 1024) {
        const int l = l0 + tid;
        const int c = l < nlist ? list_cnt[l] : 0;
        int ip = c, ib = (c + IVF_QB - 1) / IVF_QB;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tp = __shfl_up_sync(0xffffffffu, ip, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ip += tp; ib += tb; }
        }
        if (lane == 31) { wp[warp] = ip; wb[warp] = ib; }
        __syncthreads();
        int bp = carry_p, bb = carry_b;
        for (int w = 0; w < warp; w++) { bp += wp[w]; bb += wb[w]; }
        if (l <= nlist) {
            const int ep = bp + ip - c, eb = bb + ib - (c + IVF_QB - 1) / IVF_QB;     // exclusive
            pair_off[l] = ep; blk_off[l] = eb;
            if (l < nlist) cursor[l] = ep;
        }
        __syncthreads();
        if (tid == 1023) { carry_p = bp + ip; carry_b = bb + ib; }
        __syncthreads();
    }
}

__global__ void ivf_group_scatter_kernel(const long long *__restrict__ probe_list, const long long *__restrict__ probe_cnt,
                                         int nprobes, long long n_pairs, int *__restrict__ cursor, int *__restrict__ pairs_sorted) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const long long q = i / nprobes;
    if (i - q * nprobes >= probe_cnt[q]) return;
    pairs_sorted[atomicAdd(&cursor[probe_list[i]], 1)] = (int)i;
}

template <int METRIC, bool FMA>
__global__ void __launch_bounds__(128) ivf_scan_lists_kernel(
    const float *__restrict__ rows, int ld, const float *__restrict__ queries, const long long *__restrict__ q_off,
    const long long *__restrict__ tile_off, const long long *__restrict__ list_off, const uint32_t *__restrict__ members,
    int nprobes, int nlist, const int *__restrict__ pair_off, const int *__restrict__ blk_off,
    const int *__restrict__ pairs_sorted, const uint8_t *__restrict__ skip, float threshold, long long parts_per_q,
    uint64_t *__restrict__ out_keys, int *__restrict__ out_cnt) {
    constexpr int CH = 32, PCS = CH / 4, ROW_B = CH * 4, STAGE_B = 128 * ROW_B;
    extern __shared__ __align__(16) uint8_t smem[];
    float *q_s = reinterpret_cast<float *>(smem);                    // [IVF_QB][ld]
    uint8_t *stage = smem + (size_t)IVF_QB * ld * 4;                 // [2][128 rows][128 B]
    __shared__ uint32_t pos_s[128];
    __shared__ int s_cnt[IVF_QB];
    __shared__ int s_pair[IVF_QB];
    const int tid = threadIdx.x;
    // block of pairs -> list: largest l with blk_off[l] <= blockIdx.y
    const int y = blockIdx.x;                                         // x: blocks of pairs (the blocks of a list are neighbours)
    if (y >= blk_off[nlist]) return;
    int lo = 0, hi = nlist;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blk_off[mid] <= y) lo = mid; else hi = mid;
    }
    const int l = lo;
    const long long len = list_off[l + 1] - list_off[l];
    const int first = pair_off[l] + (y - blk_off[l]) * IVF_QB;
    const int nb = min(IVF_QB, pair_off[l + 1] - first);             // pairs of this block
    if ((long long)blockIdx.y * 128 >= len) return;
    if (tid < IVF_QB) s_pair[tid] = tid < nb ? pairs_sorted[first + tid] : -1;
    __syncthreads();
    for (int qi = 0; qi < IVF_QB; qi++) {
        const int pair = s_pair[qi];
        const float *src = pair >= 0 ? queries + (size_t)(pair / nprobes) * ld : nullptr;
        for (int e = tid; e < ld; e += 128) q_s[(size_t)qi * ld + e] = src ? src[e] : 0.0f;
    }
    const int n_ch = ld / CH;
    // the block's tiles of the list: gridDim.y CTAs share them (long lists: several tiles per CTA, the queries stay)
    for (int tile = blockIdx.y; (long long)tile * 128 < len; tile += gridDim.y) {
        const long long t0 = (long long)tile * 128;
        const long long j = t0 + tid;
        const bool live = j < len;
        const uint32_t pos = members[list_off[l] + (live ? j : t0)];
        __syncthreads();                                              // the previous tile is done with pos_s / s_cnt / stage
        pos_s[tid] = pos;
        if (tid < IVF_QB) s_cnt[tid] = 0;
        __syncthreads();
        auto issue = [&](int c) {
            uint8_t *dst = stage + (size_t)(c & 1) * STAGE_B;
    #pragma unroll
            for (int p = 0; p < PCS; p++) {
                const int idx = p * 128 + tid;
                const int r = idx / PCS, piece = idx % PCS;
                const float *src = rows + (size_t)pos_s[r] * ld + c * CH + piece * 4;
                const uint32_t d = smem_u32(dst + r * ROW_B + ((piece ^ (r & 7)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
        float acc[IVF_QB];
    #pragma unroll
        for (int qi = 0; qi < IVF_QB; qi++) acc[qi] = 0.0f;
        for (int c = 0; c < n_ch; c++) {
            if (c + 1 < n_ch) {
                issue(c + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                                              // also orders the q_s fill before its first use
            const uint8_t *sp = stage + (size_t)(c & 1) * STAGE_B + tid * ROW_B;
            const float *qc = q_s + c * CH;
    #pragma unroll
            for (int jj = 0; jj < PCS; jj++) {
                const float4 xv = *reinterpret_cast<const float4 *>(sp + ((jj ^ (tid & 7)) << 4));
    #pragma unroll
                for (int qi = 0; qi < IVF_QB; qi++) {
                    const float4 qv = *reinterpret_cast<const float4 *>(qc + (size_t)qi * ld + jj * 4);
                    float a = acc[qi];
                    if (FMA) {
                        a = metric_step<METRIC, FMA>(a, qv.x, xv.x);
                        a = metric_step<METRIC, FMA>(a, qv.y, xv.y);
                        a = metric_step<METRIC, FMA>(a, qv.z, xv.z);
                        a = metric_step<METRIC, FMA>(a, qv.w, xv.w);
                    } else {
                        a = metric_step4_unfused<METRIC>(a, qv, xv);
                    }
                    acc[qi] = a;
                }
            }
            __syncthreads();
        }
        const bool keep = live && !(skip != nullptr && skip[pos]);
    #pragma unroll
        for (int qi = 0; qi < IVF_QB; qi++) {
            const int pair = s_pair[qi];
            if (pair < 0) continue;                                       // uniform
            const int q = pair / nprobes, pr = pair - q * nprobes;
            const float dist = metric_finish<METRIC>(acc[qi]);
            const size_t part = (size_t)q * parts_per_q + (size_t)tile_off[(size_t)q * (nprobes + 1) + pr] + tile;
            if (keep && !(threshold > 0.0f && dist > threshold)) {
                const int slot = atomicAdd(&s_cnt[qi], 1);
                out_keys[part * 128 + slot] = make_key(dist, (uint32_t)(q_off[(size_t)q * (nprobes + 1) + pr] + j));
            }
        }
        __syncthreads();
        if (tid < nb) {
            const int pair = s_pair[tid];
            const int q = pair / nprobes, pr = pair - q * nprobes;
            out_cnt[(size_t)q * parts_per_q + tile_off[(size_t)q * (nprobes + 1) + pr] + tile] = s_cnt[tid];
        }
    }
}

// candidate numbers of the final lists -> store positions and ids.  List shards (cm_ivf_sharded_*): this index holds
// only SOME lists, so its candidate numbers count its own vectors only; with the GLOBAL list lengths the same walk also
// yields the number the candidate has in the reference's append loop over ALL probed lists (ivf_index_search.go:252-268)
// -- the tie-break key the cross-shard merge needs.
__global__ void ivf_emit_kernel(const long long *__restrict__ probe_list, const long long *__restrict__ q_off,
                                const long long *__restrict__ list_off, const uint32_t *__restrict__ members, int nprobes,
                                const uint32_t *__restrict__ row_ids, long long out_stride, uint32_t *__restrict__ out_ids,
                                long long *__restrict__ out_pos, const long long *__restrict__ out_counts,
                                const long long *__restrict__ glob_len, uint32_t *__restrict__ out_gno) {
    const int q = blockIdx.x;
    const long long m = out_counts[q];
    const long long *qo = q_off + (size_t)q * (nprobes + 1);
    const long long *pl = probe_list + (size_t)q * nprobes;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        size_t o = (size_t)q * out_stride + i;
        const long long c = (long long)out_ids[o];
        int lo = 0, hi = nprobes;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (qo[mid] <= c) lo = mid; else hi = mid;
        }
        const uint32_t pos = members[list_off[pl[lo]] + (c - qo[lo])];
        out_ids[o] = row_ids[pos];
        if (out_pos) out_pos[o] = pos;
        if (out_gno) {
            long long before = 0;
            for (int j = 0; j < lo; j++) before += glob_len[pl[j]];
            out_gno[o] = (uint32_t)(before + (c - qo[lo]));
        }
    }
}

template <int METRIC>
static int launch_ivf_scan_m(bool fma, dim3 grid, size_t smem, cudaStream_t st, const float *rows, int ld, const float *queries,
                             const long long *probe_list, const long long *q_off, const long long *list_off,
                             const uint32_t *members, int nprobes, const uint8_t *skip, float threshold, long long cap_c,
                             int n_chunks, uint64_t *out_keys, int *out_cnt) {
    const bool wide = ld % 64 == 0;
    smem = (size_t)ld * 4 + 2 * 128 * (size_t)(wide ? 256 : 128);
#define CM_IVF_GO(F, C)                                                                                                   \
    do {                                                                                                                 \
        CM_TRY(set_dyn_smem((const void *)ivf_scan_kernel<METRIC, F, C>, smem));                                               \
        ivf_scan_kernel<METRIC, F, C><<<grid, 128, smem, st>>>(rows, ld, queries, probe_list, q_off, list_off, members, nprobes, \
                                                               skip, threshold, cap_c, n_chunks, out_keys, out_cnt);     \
    } while (0)
    if (fma) { if (wide) CM_IVF_GO(true, 64); else CM_IVF_GO(true, 32); }
    else { if (wide) CM_IVF_GO(false, 64); else CM_IVF_GO(false, 32); }
#undef CM_IVF_GO
    count_launch();
    CM_CUDA(cudaGetLastError());
    return CM_OK;
}

// nq independent searchSingleQuery calls; q_dev raw queries [nq][dim] on the device
static int ivf_search_device(IVFIndex &ix, const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                             uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, cudaStream_t st,
                             bool check_zero_queries, const long long *glob_len = nullptr, uint32_t *out_gno = nullptr) {
    WsScope ws(st);
    if (nq <= 0) return CM_OK;
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");   // ivf_index_search.go:223
    int nprobes = p->nprobes;
    if (nprobes <= 0 || nprobes > ix.nlist) nprobes = ix.nlist;                                    // :233-236
    CM_TRY(ix.sync_csr(st));
    FlatIndex &S = ix.store;
    const int ld = S.ld;
    bool fma = rounding_mode() == CM_ROUND_FMA;
    const int64_t bound_c = ix.candidate_bound(nprobes);
    int64_t k_eff = p->k;
    if (k_eff <= 0 || k_eff > bound_c) k_eff = bound_c;    // sanitizeK against the most candidates any query can have
    if (out_stride < k_eff)
        return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)k_eff);
    // k beyond the shared-memory merge (WithK(0) on long lists): launch_merge_topk falls back to a sort per query

    // 1. Distance.Preprocess on the queries (ivf_index_search.go:239), zero-padded block for the scans
    int64_t nq_pad = (nq + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
    float *qp = nullptr;
    int *qflags = nullptr;
    CM_TRY(ws.get(&qp, (size_t)nq_pad * ld * 4));
    CM_TRY(ws.get(&qflags, (size_t)nq * sizeof(int)));
    if (nq_pad > nq) CM_CUDA(cudaMemsetAsync(qp + (size_t)nq * ld, 0, (size_t)(nq_pad - nq) * ld * 4, st));
    CM_TRY(launch_preprocess_rows(ix.metric, fma, q_dev, nq, ix.dim, ix.dim, qp, ld, qflags, st));
    if (check_zero_queries && ix.metric == CM_COSINE) {
        std::vector<int> hf((size_t)nq);
        CM_CUDA(cudaMemcpyAsync(hf.data(), qflags, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < nq; i++)
            if (hf[(size_t)i]) {
                return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)i);
            }
    }
    if (bound_c == 0 || k_eff == 0) {
        CM_TRY(launch_fill_counts(out_counts, nq, 0, st));       // a kernel: the counts may live on a peer device (list shards)
        return CM_OK;
    }

    // 2. coarse quantiser: exact scan of the centroids, the nprobes nearest by (distance, list index)
    uint32_t *c_ids = nullptr;
    float *c_sc = nullptr;
    long long *probe_list = nullptr, *probe_cnt = nullptr, *q_off = nullptr;
    CM_TRY(ws.get(&c_ids, (size_t)nq * nprobes * 4));
    CM_TRY(ws.get(&c_sc, (size_t)nq * nprobes * 4));
    CM_TRY(ws.get(&probe_list, (size_t)nq * nprobes * 8));
    CM_TRY(ws.get(&probe_cnt, (size_t)nq * 8));
    CM_TRY(ws.get(&q_off, (size_t)nq * (nprobes + 1) * 8));
    cm_flat_stats cst{};
    CM_TRY(ix.coarse.search_exact(qp, nq, nq_pad, nprobes, nullptr, 0.0f, nprobes, c_ids, c_sc, (int64_t *)probe_list,
                                  (int64_t *)probe_cnt, st, &cst));

    // 3. candidate numbering
    if (!ix.scanned_total) CM_CUDA(cudaMalloc(&ix.scanned_total, 8));
    CM_CUDA(cudaMemsetAsync(ix.scanned_total, 0, 8, st));
    CM_TRY(launch_ivf_offsets(probe_list, probe_cnt, ix.list_off, nprobes, nq, q_off, ix.scanned_total, st));

    // 4. soft deletes + document filter -> skip mask over store positions
    const uint8_t *skip = nullptr;
    uint8_t *skip_buf = nullptr;
    uint32_t *filt_dev = nullptr;
    if (p->filter_ids && p->nfilter > 0) {
        std::vector<uint32_t> f(p->filter_ids, p->filter_ids + p->nfilter);
        std::sort(f.begin(), f.end());
        f.erase(std::unique(f.begin(), f.end()), f.end());
        CM_TRY(ws.get(&filt_dev, f.size() * 4));
        CM_TRY(ws.get(&skip_buf, (size_t)S.n));
        CM_CUDA(cudaMemcpyAsync(filt_dev, f.data(), f.size() * 4, cudaMemcpyHostToDevice, st));
        CM_TRY(launch_build_skip(S.ids, S.deleted, S.n, filt_dev, (int64_t)f.size(), skip_buf, st));
        CM_CUDA(cudaStreamSynchronize(st));
        skip = skip_buf;
    } else if (S.n_deleted_rows > 0) {
        skip = S.deleted;
    }

    // 5. list scans, in query groups so that the key workspace stays bounded (<= 1 GiB)
    const int n_chunks = (int)((bound_c + 127) / 128);
    // batches take the list-major scan: a pair's keys go to whole 128-slot parts per tile of its list, so a query has up
    // to nprobes more parts than candidates / 128
    const int64_t max_len = ix.sizes_desc.empty() ? 0 : ix.sizes_desc[0];
    const int64_t max_tiles = (max_len + 127) / 128;
    bool list_major = nq >= 32 && max_tiles > 0;
    if (const char *e = getenv("COMET_B200_IVF_LIST_MAJOR")) list_major = atoi(e) != 0 && max_tiles > 0;
    const int n_parts = list_major ? n_chunks + nprobes : n_chunks;
    const long long cap_c = (long long)n_parts * 128;
    int64_t qgroup = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 30) / (cap_c * 8)));
    uint64_t *keys = nullptr;
    int *kcnt = nullptr;
    CM_TRY(ws.get(&keys, (size_t)qgroup * cap_c * 8));
    CM_TRY(ws.get(&kcnt, (size_t)qgroup * n_parts * 4));
    long long *tile_off = nullptr;
    int *grp = nullptr, *pairs_sorted = nullptr;                  // grp: list_cnt | pair_off | blk_off | cursor, nlist + 1 ints each
    if (list_major) {
        CM_TRY(ws.get(&tile_off, (size_t)qgroup * (nprobes + 1) * 8));
        CM_TRY(ws.get(&grp, (size_t)4 * (ix.nlist + 1) * sizeof(int)));
        CM_TRY(ws.get(&pairs_sorted, (size_t)qgroup * nprobes * sizeof(int)));
    }
    size_t smem = (size_t)ld * 4 + 2 * 128 * 128;
    for (int64_t q0 = 0; q0 < nq; q0 += qgroup) {
        int64_t m = std::min(qgroup, nq - q0);
        CM_CUDA(cudaMemsetAsync(kcnt, 0, (size_t)m * n_parts * 4, st));
        dim3 grid((unsigned)n_chunks, (unsigned)m);
        const float *qq = qp + (size_t)q0 * ld;
        const long long *pl = probe_list + (size_t)q0 * nprobes, *qo = q_off + (size_t)q0 * (nprobes + 1);
        ProfScope prof(CM_PROF_IVF_SCAN, st);
        if (list_major) {
            const int nl1 = ix.nlist + 1;
            int *list_cnt = grp, *pair_off = grp + nl1, *blk_off = grp + 2 * nl1, *cursor = grp + 3 * nl1;
            CM_CUDA(cudaMemsetAsync(list_cnt, 0, (size_t)nl1 * sizeof(int), st));
            ivf_group_count_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(pl, probe_cnt + q0, ix.list_off, nprobes, (int)m, tile_off, list_cnt);
            ivf_group_scan_kernel<<<1, 1024, 0, st>>>(list_cnt, ix.nlist, pair_off, blk_off, cursor);
            const long long n_pairs = (long long)m * nprobes;
            ivf_group_scatter_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(pl, probe_cnt + q0, nprobes, n_pairs, cursor, pairs_sorted);
            const long long blocks_max = n_pairs / IVF_QB + std::min<long long>(ix.nlist, n_pairs);
            const size_t smem_l = (size_t)IVF_QB * ld * 4 + 2 * 128 * 128;
            dim3 grid_l((unsigned)blocks_max, (unsigned)std::min<int64_t>(max_tiles, 16));    // a hub list's tiles are looped over
#define CM_IVF_LISTS(MM, F)                                                                                              \
    do {                                                                                                                 \
        CM_TRY(set_dyn_smem((const void *)ivf_scan_lists_kernel<MM, F>, smem_l));                                        \
        ivf_scan_lists_kernel<MM, F><<<grid_l, 128, smem_l, st>>>(S.rows, ld, qq, qo, tile_off, ix.list_off, ix.members, nprobes, \
                                                                 ix.nlist, pair_off, blk_off, pairs_sorted, skip, p->threshold, \
                                                                 (long long)n_parts, keys, kcnt);                       \
    } while (0)
            switch (ix.metric) {
            case CM_L2: if (fma) CM_IVF_LISTS(CM_L2, true); else CM_IVF_LISTS(CM_L2, false); break;
            case CM_L2SQ: if (fma) CM_IVF_LISTS(CM_L2SQ, true); else CM_IVF_LISTS(CM_L2SQ, false); break;
            default: if (fma) CM_IVF_LISTS(CM_COSINE, true); else CM_IVF_LISTS(CM_COSINE, false); break;
            }
#undef CM_IVF_LISTS
            for (int i = 0; i < 4; i++) count_launch();
            CM_CUDA(cudaGetLastError());
        } else
        switch (ix.metric) {
        case CM_L2: CM_TRY(launch_ivf_scan_m<CM_L2>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, ix.list_off, ix.members, nprobes, skip, p->threshold, cap_c, n_chunks, keys, kcnt)); break;
        case CM_L2SQ: CM_TRY(launch_ivf_scan_m<CM_L2SQ>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, ix.list_off, ix.members, nprobes, skip, p->threshold, cap_c, n_chunks, keys, kcnt)); break;
        default: CM_TRY(launch_ivf_scan_m<CM_COSINE>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, ix.list_off, ix.members, nprobes, skip, p->threshold, cap_c, n_chunks, keys, kcnt)); break;
        }
        CM_TRY(launch_merge_topk(keys, kcnt, (int)m, n_parts, 128, (int)k_eff, nullptr, out_stride, out_ids + (size_t)q0 * out_stride,
                                 out_scores + (size_t)q0 * out_stride, nullptr, out_counts + q0, st));
        ivf_emit_kernel<<<(unsigned)m, 128, 0, st>>>(pl, qo, ix.list_off, ix.members, nprobes, S.ids, (long long)out_stride,
                                                     out_ids + (size_t)q0 * out_stride,
                                                     out_pos ? (long long *)out_pos + (size_t)q0 * out_stride : nullptr,
                                                     (const long long *)(out_counts + q0), glob_len,
                                                     out_gno ? out_gno + (size_t)q0 * out_stride : nullptr);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    return CM_OK;
}

// ---- top-k over an explicit candidate list (selective document filters of the flat index, flat_filter.cu) ---------------
// The list scan with ONE "list" shared by every query: candidate c is scan position cand_pos[c] (ascending), so the key
// order (score, candidate number) is the flat search's (score, position).
__global__ void gather_view_kernel(const int *__restrict__ cand_cnt, long long cap, long long nq, long long *__restrict__ probe_list,
                                   long long *__restrict__ q_off, long long *__restrict__ list_off) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long m = min((long long)*cand_cnt, cap);
    if (q == 0) { list_off[0] = 0; list_off[1] = m; }
    if (q < nq) { probe_list[q] = 0; q_off[2 * q] = 0; q_off[2 * q + 1] = m; }
}

int gather_scan_topk(FlatIndex &S, const float *qp, int64_t nq, const uint32_t *cand_pos, const int *cand_cnt_dev, int64_t cap,
                     float threshold, int64_t k_eff, int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                     int64_t *out_counts, cudaStream_t st) {
    WsScope ws(st);
    if (nq <= 0) return CM_OK;
    const int ld = S.ld;
    const bool fma = rounding_mode() == CM_ROUND_FMA;
    if (k_eff > cap) k_eff = cap;
    if (k_eff <= 0) return launch_fill_counts(out_counts, nq, 0, st);
    long long *probe_list = nullptr, *q_off = nullptr, *list_off = nullptr;
    CM_TRY(ws.get(&probe_list, (size_t)nq * 8));
    CM_TRY(ws.get(&q_off, (size_t)nq * 16));
    CM_TRY(ws.get(&list_off, 16));
    gather_view_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(cand_cnt_dev, (long long)cap, (long long)nq, probe_list, q_off, list_off);
    count_launch();
    CM_CUDA(cudaGetLastError());
    const int n_chunks = (int)((cap + 127) / 128);
    const long long cap_c = (long long)n_chunks * 128;
    int64_t qgroup = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 30) / (cap_c * 8)));
    qgroup = std::min<int64_t>(qgroup, 65535);
    uint64_t *keys = nullptr;
    int *kcnt = nullptr;
    CM_TRY(ws.get(&keys, (size_t)qgroup * cap_c * 8));
    CM_TRY(ws.get(&kcnt, (size_t)qgroup * n_chunks * 4));
    size_t smem = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += qgroup) {
        const int64_t m = std::min(qgroup, nq - q0);
        CM_CUDA(cudaMemsetAsync(kcnt, 0, (size_t)m * n_chunks * 4, st));
        dim3 grid((unsigned)n_chunks, (unsigned)m);
        const float *qq = qp + (size_t)q0 * ld;
        const long long *pl = probe_list + q0, *qo = q_off + 2 * q0;
        ProfScope prof(CM_PROF_IVF_SCAN, st);
        switch (S.metric) {
        case CM_L2: CM_TRY(launch_ivf_scan_m<CM_L2>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, list_off, cand_pos, 1, nullptr, threshold, cap_c, n_chunks, keys, kcnt)); break;
        case CM_L2SQ: CM_TRY(launch_ivf_scan_m<CM_L2SQ>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, list_off, cand_pos, 1, nullptr, threshold, cap_c, n_chunks, keys, kcnt)); break;
        default: CM_TRY(launch_ivf_scan_m<CM_COSINE>(fma, grid, smem, st, S.rows, ld, qq, pl, qo, list_off, cand_pos, 1, nullptr, threshold, cap_c, n_chunks, keys, kcnt)); break;
        }
        CM_TRY(launch_merge_topk(keys, kcnt, (int)m, n_chunks, 128, (int)k_eff, nullptr, out_stride, out_ids + (size_t)q0 * out_stride,
                                 out_scores + (size_t)q0 * out_stride, nullptr, out_counts + q0, st));
        ivf_emit_kernel<<<(unsigned)m, 128, 0, st>>>(pl, qo, list_off, cand_pos, 1, S.ids, (long long)out_stride,
                                                     out_ids + (size_t)q0 * out_stride,
                                                     out_pos ? (long long *)out_pos + (size_t)q0 * out_stride : nullptr,
                                                     (const long long *)(out_counts + q0), nullptr, nullptr);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    return CM_OK;
}

}  // namespace cm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct cm_ivf {
    cm::IVFIndex ix;
};

extern "C" {

int cm_ivf_create(int dim, int nlist, int metric, cm_ivf **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");             // ivf_index.go:149
    if (nlist <= 0) return cm::fail(CM_ERR_INVALID_ARG, "nlist must be positive");               // ivf_index.go:152
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    CM_TRY(cm::ensure_device());
    cm_ivf *h = new cm_ivf();
    h->ix.dim = dim; h->ix.nlist = nlist; h->ix.metric = metric;
    cudaGetDevice(&h->ix.device);
    for (cm::FlatIndex *f : {&h->ix.coarse, &h->ix.store}) {
        f->dim = dim;
        f->ld = (dim + cm::SCAN_CHUNK - 1) / cm::SCAN_CHUNK * cm::SCAN_CHUNK;
        f->metric = metric;
        f->device = h->ix.device;
    }
    h->ix.coarse.raw_rows = true;
    h->ix.lists.resize((size_t)nlist);
    *out = h;
    return CM_OK;
}
int cm_ivf_destroy(cm_ivf *h) {
    delete h;
    return CM_OK;
}

// Load trained centroids (nlist x dim, row-major) -- the result of IVFIndex.Train (ivf_index.go:205-246).
int cm_ivf_set_centroids(cm_ivf *h, const float *centroids) {
    if (!h || !centroids) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    if (h->ix.store.n > 0) return cm::fail(CM_ERR_INVALID_ARG, "cannot replace centroids of a non-empty index");
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    cm::FlatIndex &c = h->ix.coarse;
    c.n = 0; c.ids_host_mirror.clear();
    float *stage = nullptr;
    size_t bytes = (size_t)h->ix.nlist * h->ix.dim * 4;
    int rc = cm::ws_alloc((void **)&stage, bytes, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(stage, centroids, bytes, cudaMemcpyHostToDevice, st);
        std::vector<uint32_t> ids((size_t)h->ix.nlist);
        std::iota(ids.begin(), ids.end(), 0u);
        rc = c.add_from_device(ids.data(), stage, h->ix.nlist, nullptr, st);
    }
    cm::ws_free(stage, st);
    cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK) h->ix.trained = true;
    return rc;
}
// IVFIndex.Train (ivf_index.go:205-246): KMeans(raw vectors, nlist, index distance, 20) on the device
int cm_ivf_train(cm_ivf *h, const float *rows, int64_t n) {
    if (!h || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n < h->ix.nlist)
        return cm::fail(CM_ERR_TOO_FEW, "need at least %d training vectors for %d clusters (got %lld)", h->ix.nlist, h->ix.nlist, (long long)n);
    if (h->ix.store.n > 0) return cm::fail(CM_ERR_UNSUPPORTED, "retraining a non-empty index is not supported");
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *d = nullptr;
    int rc = cm::upload_training_rows(rows, n, h->ix.dim, h->ix.coarse.ld, &d, st);
    if (rc == CM_OK) rc = cm::kmeans_full(h->ix.coarse, d, n, h->ix.nlist, 20, nullptr, st);
    cm::ws_free(d, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) rc = cm::fail(CM_ERR_CUDA, "ivf_train: %s", cudaGetErrorString(e));
    if (rc == CM_OK) h->ix.trained = true;
    return rc;
}
// trained centroids back to the host (nlist x dim), e.g. for IVFIndex.WriteTo
int cm_ivf_get_centroids(const cm_ivf *h, float *out) {
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained");
    CM_CUDA(cudaSetDevice(h->ix.device));
    CM_CUDA(cudaMemcpy2D(out, (size_t)h->ix.dim * 4, h->ix.coarse.rows, (size_t)h->ix.coarse.ld * 4, (size_t)h->ix.dim * 4,
                         (size_t)h->ix.nlist, cudaMemcpyDeviceToHost));
    return CM_OK;
}
int cm_ivf_trained(const cm_ivf *h) { return h && h->ix.trained ? 1 : 0; }
int64_t cm_ivf_size(const cm_ivf *h) { return h ? h->ix.store.n : 0; }
int cm_ivf_default_nprobes(const cm_ivf *h) {   // ivf_index.go:410 int(sqrt(nlist))
    if (!h) return 0;
    int r = (int)std::sqrt((double)h->ix.nlist);
    while ((long long)(r + 1) * (r + 1) <= h->ix.nlist) r++;
    while ((long long)r * r > h->ix.nlist) r--;
    return r;
}

// n successive IVFIndex.Add calls (ivf_index.go:258-283): PreprocessInPlace, nearest centroid
// (FindNearestCentroidIndex, clustering.go:252-272: first minimum wins), append to that list.
int cm_ivf_add(cm_ivf *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");   // ivf_index.go:264
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::IVFIndex &ix = h->ix;
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    int rc = CM_OK;
    const int64_t slab = std::max<int64_t>(1, (int64_t)(128u << 20) / ((int64_t)ix.dim * 4));
    float *stage = nullptr;
    uint32_t *a_ids = nullptr;
    float *a_sc = nullptr;
    long long *a_pos = nullptr, *a_cnt = nullptr;
    int64_t sl = std::min(slab, n);
    rc = cm::ws_alloc((void **)&stage, (size_t)sl * ix.dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&a_ids, (size_t)sl * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&a_sc, (size_t)sl * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&a_pos, (size_t)sl * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&a_cnt, (size_t)sl * 8, st);
    std::vector<long long> hpos((size_t)sl);
    for (int64_t i0 = 0; rc == CM_OK && i0 < n; i0 += slab) {
        int64_t m = std::min(slab, n - i0);
        int64_t n_before = ix.store.n;
        cudaMemcpyAsync(stage, rows + (size_t)i0 * ix.dim, (size_t)m * ix.dim * 4, cudaMemcpyHostToDevice, st);
        int rc_add = ix.store.add_from_device(ids + i0, stage, m, writeback ? rows + (size_t)i0 * ix.dim : nullptr, st);
        int64_t good = ix.store.n - n_before;          // rows before a zero vector were added, like n successive Add()s
        if (good > 0) {
            // assignment: the stored (preprocessed) rows are the queries of a k = 1 exact scan of the centroids
            int64_t gpad = (good + cm::SCAN_MAX_QB - 1) / cm::SCAN_MAX_QB * cm::SCAN_MAX_QB;
            rc = ix.store.reserve(n_before + gpad);     // the scan reads query blocks of 8 rows: keep the tail in bounds
            cm_flat_stats cst{};
            if (rc == CM_OK)
                rc = ix.coarse.search_exact(ix.store.rows + (size_t)n_before * ix.store.ld, good, gpad, 1, nullptr, 0.0f, 1, a_ids,
                                            a_sc, (int64_t *)a_pos, (int64_t *)a_cnt, st, &cst);
            if (rc == CM_OK) {
                cudaMemcpyAsync(hpos.data(), a_pos, (size_t)good * 8, cudaMemcpyDeviceToHost, st);
                cudaError_t e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) rc = cm::fail(CM_ERR_CUDA, "ivf_add: %s", cudaGetErrorString(e));
            }
            if (rc == CM_OK) {
                for (int64_t i = 0; i < good; i++) {
                    int32_t l = (int32_t)hpos[(size_t)i];
                    ix.lists[(size_t)l].push_back((uint32_t)(n_before + i));
                    ix.list_of.push_back(l);
                    if (out_lists) out_lists[i0 + i] = l;
                }
                ix.csr_dirty = true;
            }
        }
        if (rc == CM_OK) rc = rc_add;
    }
    cm::ws_free(stage, st); cm::ws_free(a_ids, st); cm::ws_free(a_sc, st); cm::ws_free(a_pos, st); cm::ws_free(a_cnt, st);
    cudaStreamSynchronize(st);
    cm::release_stream(st);
    return rc;
}

// candidates (vectors of probed lists) scanned by the last search on this handle, all queries together:
// the measured factor of the scan's algorithmic bytes
int64_t cm_ivf_last_scanned(const cm_ivf *h) {
    if (!h || !h->ix.scanned_total) return 0;
    unsigned long long v = 0;
    cudaSetDevice(h->ix.device);
    if (cudaMemcpy(&v, h->ix.scanned_total, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}

// IVFIndex.ReadFrom (ivf_index.go:611-785): restore stored vectors with the list each one belongs to,
// in the order given (the order inside a list is the order of appearance here).  Nothing is re-derived.
int cm_ivf_load_lists(cm_ivf *h, const uint32_t *ids, const float *rows, const int32_t *list_of, int64_t n) {
    if (!h || (n > 0 && (!ids || !rows || !list_of))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained (centroids loaded) before vectors are restored");
    if (n <= 0) return CM_OK;
    for (int64_t i = 0; i < n; i++)
        if (list_of[i] < 0 || list_of[i] >= h->ix.nlist) return cm::fail(CM_ERR_INVALID_ARG, "row %lld: list %d out of range", (long long)i, list_of[i]);
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::IVFIndex &ix = h->ix;
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    const int64_t slab = std::max<int64_t>(1, (int64_t)(128u << 20) / ((int64_t)ix.dim * 4));
    float *stage = nullptr;
    int rc = cm::ws_alloc((void **)&stage, (size_t)std::min(slab, n) * ix.dim * 4, st);
    const bool was_raw = ix.store.raw_rows;
    ix.store.raw_rows = true;
    for (int64_t i0 = 0; rc == CM_OK && i0 < n; i0 += slab) {
        int64_t m = std::min(slab, n - i0);
        int64_t n_before = ix.store.n;
        cudaMemcpyAsync(stage, rows + (size_t)i0 * ix.dim, (size_t)m * ix.dim * 4, cudaMemcpyHostToDevice, st);
        rc = ix.store.add_from_device(ids + i0, stage, m, nullptr, st);
        if (rc != CM_OK) break;
        for (int64_t i = 0; i < m; i++) {
            int32_t l = list_of[i0 + i];
            ix.lists[(size_t)l].push_back((uint32_t)(n_before + i));
            ix.list_of.push_back(l);
        }
        ix.csr_dirty = true;
    }
    ix.store.raw_rows = was_raw;
    cm::ws_free(stage, st);
    cudaStreamSynchronize(st);
    cm::release_stream(st);
    return rc;
}

// node IDs by store position (with cm_ivf_get_rows: what a host mirror needs after cm_ivf_load)
int cm_ivf_get_ids(const cm_ivf *h, int64_t first, int64_t n, uint32_t *out) {
    if (!h || first < 0 || n < 0 || first + n > h->ix.store.n || (n > 0 && !out)) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n > 0) memcpy(out, h->ix.store.ids_host_mirror.data() + first, (size_t)n * 4);
    return CM_OK;
}

int cm_ivf_remove(cm_ivf *h, uint32_t id) {     // ivf_index.go:296-330 soft delete
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.store.remove(id);
}

int cm_ivf_flush(cm_ivf *h) {                   // ivf_index.go:342-390: drop deleted vectors from every list
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::IVFIndex &ix = h->ix;
    if (ix.store.deleted_ids.empty()) return CM_OK;
    int64_t n_old = ix.store.n;
    std::vector<int64_t> new_pos((size_t)n_old, -1);
    int64_t m = 0;
    for (int64_t i = 0; i < n_old; i++)
        if (!ix.store.deleted_ids.count(ix.store.ids_host_mirror[(size_t)i])) new_pos[(size_t)i] = m++;
    CM_TRY(ix.store.flush());
    std::vector<int32_t> nlo((size_t)m);
    for (auto &l : ix.lists) {
        size_t w = 0;
        for (uint32_t pos : l)
            if (new_pos[pos] >= 0) l[w++] = (uint32_t)new_pos[pos];
        l.resize(w);
    }
    for (int64_t i = 0; i < n_old; i++)
        if (new_pos[(size_t)i] >= 0) nlo[(size_t)new_pos[(size_t)i]] = ix.list_of[(size_t)i];
    ix.list_of.swap(nlo);
    ix.csr_dirty = true;
    return CM_OK;
}

int cm_ivf_get_rows(const cm_ivf *h, const int64_t *positions, int64_t n, float *out) {
    if (!h || (n > 0 && (!positions || !out))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n <= 0) return CM_OK;
    const cm::FlatIndex &S = h->ix.store;
    for (int64_t i = 0; i < n; i++)
        if (positions[i] < 0 || positions[i] >= S.n) return cm::fail(CM_ERR_NOT_FOUND, "position %lld out of range", (long long)positions[i]);
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    int64_t *dpos = nullptr;
    float *dout = nullptr;
    int rc = cm::ws_alloc((void **)&dpos, (size_t)n * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dout, (size_t)n * S.dim * 4, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dpos, positions, (size_t)n * 8, cudaMemcpyHostToDevice, st);
        rc = cm::launch_gather_rows(S.rows, S.ld, S.dim, dpos, n, dout, st);
        cudaMemcpyAsync(out, dout, (size_t)n * S.dim * 4, cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(dpos, st); cm::ws_free(dout, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "ivf_get_rows: %s", cudaGetErrorString(e));
    return rc;
}

int cm_ivf_search_device(cm_ivf *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                         int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                         int64_t *out_counts_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ix.dim)
        return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::ivf_search_device(h->ix, queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                                 out_counts_dev, (cudaStream_t)stream, false);
}

int cm_ivf_search(cm_ivf *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                  uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ix.dim)
        return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    if (nq <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *dq = nullptr, *dsc = nullptr;
    uint32_t *dids = nullptr;
    int64_t *dpos = nullptr, *dcnt = nullptr;
    size_t no = (size_t)nq * (size_t)(out_stride > 0 ? out_stride : 1);
    int rc = cm::ws_alloc((void **)&dq, (size_t)nq * dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dids, no * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dsc, no * 4, st);
    if (rc == CM_OK && out_pos) rc = cm::ws_alloc((void **)&dpos, no * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dcnt, (size_t)nq * 8, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dq, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st);
        rc = cm::ivf_search_device(h->ix, dq, nq, p, out_stride, dids, dsc, dpos, dcnt, st, true);
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(out_ids, dids, no * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_scores, dsc, no * 4, cudaMemcpyDeviceToHost, st);
        if (out_pos) cudaMemcpyAsync(out_pos, dpos, no * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, dcnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, st);
    }
    cm::ws_free(dq, st); cm::ws_free(dids, st); cm::ws_free(dsc, st); cm::ws_free(dpos, st); cm::ws_free(dcnt, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "ivf_search: %s", cudaGetErrorString(e));
    return rc;
}

// ---- wire format "IVFX" (ivf_index.go:468-785) ------------------------------------------------------------------
// back to the state NewIVFIndex leaves: untrained, no vectors
static int ivf_reset(cm_ivf *h) {
    cm::IVFIndex &ix = h->ix;
    CM_TRY(ix.store.reset());
    CM_TRY(ix.coarse.reset());
    for (auto &l : ix.lists) l.clear();
    ix.list_of.clear();
    ix.trained = false;
    ix.csr_dirty = true;
    return CM_OK;
}

// IVFIndex.WriteTo: Flush, header, nlist, trained flag, centroids (size + data each), list count, per list (size, per
// vector ID + data), roaring blob.
static int ivf_save(cm_ivf *h, cm::wire::Sink &s) {
    cm::IVFIndex &ix = h->ix;
    CM_TRY(cm_ivf_flush(h));
    CM_TRY(cm::wire::write_header(s, "IVFX", ix.dim, ix.metric));
    CM_WIRE_PUT(s.u32((uint32_t)ix.nlist), "nlist");
    CM_WIRE_PUT(s.u8(ix.trained ? 1 : 0), "trained flag");
    if (ix.trained) {
        std::vector<float> c((size_t)ix.nlist * ix.dim);
        CM_TRY(cm_ivf_get_centroids(h, c.data()));
        for (int l = 0; l < ix.nlist; l++) {
            CM_WIRE_PUT(s.u32((uint32_t)ix.dim), "centroid size");
            CM_WIRE_PUT(s.put(&c[(size_t)l * ix.dim], (size_t)ix.dim * 4), "centroid data");
        }
    }
    CM_WIRE_PUT(s.u32((uint32_t)ix.lists.size()), "list count");
    const size_t per = 4 + (size_t)ix.dim * 4;
    const int64_t slab = std::max<int64_t>(1, (int64_t)(64u << 20) / (int64_t)per);
    std::vector<float> rows;
    std::vector<int64_t> pos;
    std::vector<uint8_t> rec;
    for (size_t l = 0; l < ix.lists.size(); l++) {
        const std::vector<uint32_t> &L = ix.lists[l];
        CM_WIRE_PUT(s.u32((uint32_t)L.size()), "list size");
        for (int64_t i0 = 0; i0 < (int64_t)L.size(); i0 += slab) {
            const int64_t m = std::min<int64_t>(slab, (int64_t)L.size() - i0);
            pos.assign(L.begin() + i0, L.begin() + i0 + m);
            rows.resize((size_t)m * ix.dim);
            CM_TRY(cm_ivf_get_rows(h, pos.data(), m, rows.data()));
            rec.resize((size_t)m * per);
            for (int64_t i = 0; i < m; i++) {
                const uint32_t id = ix.store.ids_host_mirror[(size_t)pos[(size_t)i]];
                memcpy(&rec[(size_t)i * per], &id, 4);
                memcpy(&rec[(size_t)i * per + 4], &rows[(size_t)i * ix.dim], (size_t)ix.dim * 4);
            }
            CM_WIRE_PUT(s.put(rec.data(), rec.size()), "list data");
        }
    }
    CM_WIRE_PUT(cm::wire::write_empty_bitmap(s), "bitmap");
    return CM_OK;
}

// IVFIndex.ReadFrom: everything is decoded and validated before the index state is replaced.  The store keeps the
// stream's order (list by list); the search only depends on the order inside each list.
static int ivf_load(cm_ivf *h, cm::wire::Source &s) {
    cm::IVFIndex &ix = h->ix;
    CM_TRY(cm::wire::read_header(s, "IVFX", ix.dim, ix.metric));
    uint32_t nlist = 0, list_count = 0;
    uint8_t trained = 0;
    CM_WIRE_GET(s.u32(&nlist), "nlist");
    if ((int64_t)nlist != ix.nlist) return cm::fail(CM_ERR_INVALID_ARG, "nlist mismatch: index has nlist=%d, serialized data has nlist=%u", ix.nlist, nlist);
    CM_WIRE_GET(s.u8(&trained), "trained flag");
    std::vector<float> cent;
    if (trained == 1) {
        cent.resize((size_t)ix.nlist * ix.dim);
        for (int l = 0; l < ix.nlist; l++) {
            uint32_t sz = 0;
            CM_WIRE_GET(s.u32(&sz), "centroid size");
            if ((int64_t)sz != ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "centroid %d has dimension %u, expected %d", l, sz, ix.dim);
            CM_WIRE_GET(s.get(&cent[(size_t)l * ix.dim], (size_t)ix.dim * 4), "centroid data");
        }
    }
    CM_WIRE_GET(s.u32(&list_count), "list count");
    if (list_count > (uint32_t)ix.nlist) return cm::fail(CM_ERR_INVALID_ARG, "serialized data has %u lists, index has nlist=%d", list_count, ix.nlist);
    std::vector<uint32_t> ids;
    std::vector<int32_t> list_of;
    std::vector<float> rows;
    for (uint32_t l = 0; l < list_count; l++) {
        uint32_t sz = 0;
        CM_WIRE_GET(s.u32(&sz), "list size");
        const size_t at = ids.size();
        ids.resize(at + sz);
        list_of.resize(at + sz, (int32_t)l);
        rows.resize((at + sz) * (size_t)ix.dim);
        for (uint32_t i = 0; i < sz; i++) {
            CM_WIRE_GET(s.u32(&ids[at + i]), "list vector ID");
            CM_WIRE_GET(s.get(&rows[(at + i) * (size_t)ix.dim], (size_t)ix.dim * 4), "list vector data");
        }
    }
    std::vector<uint32_t> dead;
    CM_TRY(cm::wire::read_bitmap(s, &dead));
    if (!ids.empty() && trained != 1) return cm::fail(CM_ERR_NOT_TRAINED, "serialized data holds vectors but no centroids");
    CM_TRY(ivf_reset(h));
    if (trained == 1) CM_TRY(cm_ivf_set_centroids(h, cent.data()));
    if (!ids.empty()) CM_TRY(cm_ivf_load_lists(h, ids.data(), rows.data(), list_of.data(), (int64_t)ids.size()));
    for (uint32_t id : dead)
        if (ix.store.remove(id) != CM_OK) ix.store.deleted_ids.insert(id);
    return CM_OK;
}

int cm_ivf_save(cm_ivf *h, uint8_t *buf, int64_t cap, int64_t *bytes) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_buffer([&](cm::wire::Sink &s) { return ivf_save(h, s); }, buf, cap, bytes);
}
int cm_ivf_load(cm_ivf *h, const uint8_t *buf, int64_t len, int64_t *consumed) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_buffer([&](cm::wire::Source &s) { return ivf_load(h, s); }, buf, len, consumed);
}
int cm_ivf_save_file(cm_ivf *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_file([&](cm::wire::Sink &s) { return ivf_save(h, s); }, path);
}
int cm_ivf_load_file(cm_ivf *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_file([&](cm::wire::Source &s) { return ivf_load(h, s); }, path);
}

// ================================================================================================
// IVF list shards over the GPUs of one box, ONE host process: cm_ivf_sharded_*.  The search driver, the
// global candidate numbering and the cross-shard merge are shared with IVFPQ: list_shards.cuh.
// ================================================================================================
struct cm_ivf_sharded {
    cm::ListShards ls;
    std::vector<cm_ivf *> shard;
    cm_ivf *assigner = nullptr;             // devices[0]: centroids only; Add runs PreprocessInPlace + nearest centroid here
};

static int ivfs_clear_vectors(cm_ivf *h) {          // drop the vectors, keep the centroids
    cm::IVFIndex &ix = h->ix;
    CM_TRY(ix.store.reset());
    for (auto &l : ix.lists) l.clear();
    ix.list_of.clear();
    ix.csr_dirty = true;
    return CM_OK;
}

int cm_ivf_sharded_destroy(cm_ivf_sharded *h) {
    if (!h) return CM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    h->ls.workers.clear();
    for (size_t r = 0; r < h->shard.size(); r++)
        if (h->shard[r]) { cudaSetDevice(h->ls.dev[r]); cm_ivf_destroy(h->shard[r]); }
    if (h->assigner) { cudaSetDevice(h->ls.dev[0]); cm_ivf_destroy(h->assigner); }
    h->ls.destroy();
    cudaSetDevice(prev);
    delete h;
    return CM_OK;
}

int cm_ivf_sharded_create(int dim, int nlist, int metric, const int *devices, int n_devices, cm_ivf_sharded **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");
    if (nlist <= 0) return cm::fail(CM_ERR_INVALID_ARG, "nlist must be positive");
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    cm_ivf_sharded *h = new cm_ivf_sharded();
    int rc = h->ls.init(dim, nlist, metric, devices, n_devices);
    if (rc != CM_OK) { delete h; return rc; }
    int prev = 0;
    cudaGetDevice(&prev);
    h->shard.assign((size_t)n_devices, nullptr);
    for (int r = 0; r < n_devices && rc == CM_OK; r++) {
        cudaSetDevice(devices[r]);
        rc = cm_ivf_create(dim, nlist, metric, &h->shard[(size_t)r]);
    }
    if (rc == CM_OK) {
        cudaSetDevice(devices[0]);
        rc = cm_ivf_create(dim, nlist, metric, &h->assigner);
    }
    cudaSetDevice(prev);
    if (rc != CM_OK) { cm_ivf_sharded_destroy(h); return rc; }
    *out = h;
    return CM_OK;
}

int cm_ivf_sharded_shards(const cm_ivf_sharded *h) { return h ? h->ls.W() : 0; }
int cm_ivf_sharded_trained(const cm_ivf_sharded *h) { return h && h->assigner && h->assigner->ix.trained ? 1 : 0; }
int64_t cm_ivf_sharded_size(const cm_ivf_sharded *h) {
    int64_t n = 0;
    if (h) for (cm_ivf *s : h->shard) n += cm_ivf_size(s);
    return n;
}
int cm_ivf_sharded_default_nprobes(const cm_ivf_sharded *h) { return h ? cm_ivf_default_nprobes(h->assigner) : 0; }
int cm_ivf_sharded_owner(const cm_ivf_sharded *h, int list) { return (h && list >= 0 && list < h->ls.nlist) ? h->ls.owner[(size_t)list] : -1; }
int cm_ivf_sharded_shard_size(const cm_ivf_sharded *h, int shard, int64_t *rows) {
    if (!h || shard < 0 || shard >= h->ls.W() || !rows) return cm::fail(CM_ERR_INVALID_ARG, "bad shard");
    *rows = cm_ivf_size(h->shard[(size_t)shard]);
    return CM_OK;
}

// the result of IVFIndex.Train goes to every shard (and the assigner): the coarse step is replicated
int cm_ivf_sharded_set_centroids(cm_ivf_sharded *h, const float *centroids) {
    if (!h || !centroids) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (cm_ivf_sharded_size(h) > 0) return cm::fail(CM_ERR_INVALID_ARG, "cannot replace centroids of a non-empty index");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = cm_ivf_set_centroids(h->assigner, centroids);
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) rc = cm_ivf_set_centroids(h->shard[r], centroids);
    cudaSetDevice(prev);
    return rc;
}
int cm_ivf_sharded_train(cm_ivf_sharded *h, const float *rows, int64_t n) {     // IVFIndex.Train (ivf_index.go:205-246) on devices[0]
    if (!h || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (cm_ivf_sharded_size(h) > 0) return cm::fail(CM_ERR_UNSUPPORTED, "retraining a non-empty index is not supported");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = cm_ivf_train(h->assigner, rows, n);
    std::vector<float> c((size_t)h->ls.nlist * h->ls.dim);
    if (rc == CM_OK) rc = cm_ivf_get_centroids(h->assigner, c.data());
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) rc = cm_ivf_set_centroids(h->shard[r], c.data());
    cudaSetDevice(prev);
    return rc;
}
int cm_ivf_sharded_get_centroids(const cm_ivf_sharded *h, float *out) {
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    return cm_ivf_get_centroids(h->assigner, out);
}

// n successive IVFIndex.Add calls (ivf_index.go:258-283): PreprocessInPlace + FindNearestCentroidIndex run once, on
// devices[0]; every stored vector then goes to the shard that owns its list, lists keep arrival order.
int cm_ivf_sharded_add(cm_ivf_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivf_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
    if (n <= 0) return CM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cm::ListShards &ls = h->ls;
    const int W = ls.W(), dim = ls.dim;
    const int64_t slab = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)dim * 4));
    std::vector<float> stored;
    std::vector<int32_t> lists((size_t)std::min(slab, n));
    std::vector<std::vector<uint32_t>> sub_ids((size_t)W);
    std::vector<std::vector<int32_t>> sub_lists((size_t)W);
    std::vector<std::vector<float>> sub_rows((size_t)W);
    int rc = CM_OK, rc_zero = CM_OK;
    std::string zero_msg;
    for (int64_t i0 = 0; i0 < n && rc == CM_OK && rc_zero == CM_OK; i0 += slab) {
        const int64_t m = std::min(slab, n - i0);
        float *batch = rows + (size_t)i0 * dim;
        if (!writeback) { stored.assign(batch, batch + (size_t)m * dim); batch = stored.data(); }
        int rc_a = cm_ivf_add(h->assigner, ids + i0, batch, m, 1, lists.data());      // rows come back preprocessed
        const int64_t good = cm_ivf_size(h->assigner);
        if (rc_a == CM_ERR_ZERO_VECTOR) { rc_zero = rc_a; zero_msg = cm_last_error(); } else rc = rc_a;
        if (rc == CM_OK) { cudaSetDevice(ls.dev[0]); rc = ivfs_clear_vectors(h->assigner); }
        for (int r = 0; r < W; r++) { sub_ids[(size_t)r].clear(); sub_lists[(size_t)r].clear(); sub_rows[(size_t)r].clear(); }
        for (int64_t i = 0; i < good && rc == CM_OK; i++) {
            const int32_t l = lists[(size_t)i];
            const int r = ls.owner[(size_t)l];
            sub_ids[(size_t)r].push_back(ids[i0 + i]);
            sub_lists[(size_t)r].push_back(l);
            sub_rows[(size_t)r].insert(sub_rows[(size_t)r].end(), batch + (size_t)i * dim, batch + (size_t)(i + 1) * dim);
            ls.glob_len[(size_t)l]++;
            if (out_lists) out_lists[i0 + i] = l;
        }
        for (int r = 0; r < W && rc == CM_OK; r++)
            if (!sub_ids[(size_t)r].empty())
                rc = cm_ivf_load_lists(h->shard[(size_t)r], sub_ids[(size_t)r].data(), sub_rows[(size_t)r].data(), sub_lists[(size_t)r].data(),
                                       (int64_t)sub_ids[(size_t)r].size());
        ls.len_dirty = true;
    }
    cudaSetDevice(prev);
    if (rc != CM_OK) return rc;
    return rc_zero == CM_OK ? CM_OK : cm::fail(CM_ERR_ZERO_VECTOR, "%s", zero_msg.c_str());
}

int cm_ivf_sharded_remove(cm_ivf_sharded *h, uint32_t id) {      // ivf_index.go:296-330: soft delete wherever the ID lives
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    bool found = false, already = false;
    for (cm_ivf *s : h->shard) {
        int rc = cm_ivf_remove(s, id);
        if (rc == CM_OK) found = true;
        else if (rc == CM_ERR_NOT_FOUND && strstr(cm_last_error(), "already")) already = true;
        else if (rc != CM_ERR_NOT_FOUND) { cudaSetDevice(prev); return rc; }
    }
    cudaSetDevice(prev);
    if (found) return CM_OK;
    return cm::fail(CM_ERR_NOT_FOUND, already ? "vector with ID %u already deleted" : "vector with ID %u not found", id);
}

int cm_ivf_sharded_flush(cm_ivf_sharded *h) {                    // ivf_index.go:342-390, shard by shard
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    std::fill(h->ls.glob_len.begin(), h->ls.glob_len.end(), 0);
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) {
        rc = cm_ivf_flush(h->shard[r]);
        for (int l = 0; l < h->ls.nlist; l++) h->ls.glob_len[(size_t)l] += (long long)h->shard[r]->ix.lists[(size_t)l].size();
    }
    h->ls.len_dirty = true;
    cudaSetDevice(prev);
    return rc;
}

// Greedy-by-length assignment of lists to shards (longest list first onto the lightest shard), then every list that
// changed owner moves.  Ties in the reference's order are unaffected: a list moves whole, in its order.
int cm_ivf_sharded_rebalance(cm_ivf_sharded *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    cm::ListShards &ls = h->ls;
    const int W = ls.W(), dim = ls.dim;
    const std::vector<int> want = ls.greedy_plan();
    for (cm_ivf *s : h->shard)
        if (!s->ix.store.deleted_ids.empty()) return cm::fail(CM_ERR_UNSUPPORTED, "flush before rebalancing");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    // movers per source shard are appended to their targets, then the source keeps what stays
    for (int r = 0; r < W && rc == CM_OK; r++) {
        cm::IVFIndex &ix = h->shard[(size_t)r]->ix;
        std::vector<std::vector<int64_t>> pos_to((size_t)W);
        for (int l = 0; l < ls.nlist; l++)
            if (ls.owner[(size_t)l] == r && want[(size_t)l] != r)
                for (uint32_t p : ix.lists[(size_t)l]) pos_to[(size_t)want[(size_t)l]].push_back((int64_t)p);
        bool any = false;
        for (int t = 0; t < W && rc == CM_OK; t++) {
            std::vector<int64_t> &pos = pos_to[(size_t)t];
            if (pos.empty()) continue;
            any = true;
            std::vector<float> rows(pos.size() * (size_t)dim);
            std::vector<uint32_t> ids(pos.size());
            std::vector<int32_t> lo(pos.size());
            rc = cm_ivf_get_rows(h->shard[(size_t)r], pos.data(), (int64_t)pos.size(), rows.data());
            for (size_t i = 0; i < pos.size(); i++) { ids[i] = ix.store.ids_host_mirror[(size_t)pos[i]]; lo[i] = ix.list_of[(size_t)pos[i]]; }
            if (rc == CM_OK) rc = cm_ivf_load_lists(h->shard[(size_t)t], ids.data(), rows.data(), lo.data(), (int64_t)pos.size());
        }
        if (any && rc == CM_OK) {
            std::vector<int64_t> keep;
            for (int64_t i = 0; i < ix.store.n; i++) {
                const int l = ix.list_of[(size_t)i];
                if (!(ls.owner[(size_t)l] == r && want[(size_t)l] != r)) keep.push_back(i);      // everything that is not leaving
            }
            std::vector<float> rows(keep.size() * (size_t)dim);
            std::vector<uint32_t> ids(keep.size());
            std::vector<int32_t> lo(keep.size());
            if (!keep.empty()) rc = cm_ivf_get_rows(h->shard[(size_t)r], keep.data(), (int64_t)keep.size(), rows.data());
            for (size_t i = 0; i < keep.size(); i++) { ids[i] = ix.store.ids_host_mirror[(size_t)keep[i]]; lo[i] = ix.list_of[(size_t)keep[i]]; }
            if (rc == CM_OK) { cudaSetDevice(ls.dev[(size_t)r]); rc = ivfs_clear_vectors(h->shard[(size_t)r]); }
            if (rc == CM_OK && !keep.empty()) rc = cm_ivf_load_lists(h->shard[(size_t)r], ids.data(), rows.data(), lo.data(), (int64_t)keep.size());
        }
    }
    if (rc == CM_OK) ls.owner = want;
    cudaSetDevice(prev);
    return rc;
}

static cm::ListShards::ShardSearch ivfs_search_fn(cm_ivf_sharded *h) {
    return [h](int r, const float *q, int64_t nq, const cm_search_params *p, int64_t K, uint32_t *o_ids, float *o_sc, int64_t *o_cnt,
               cudaStream_t s, const long long *glob_len, uint32_t *o_gno) {
        return cm::ivf_search_device(h->shard[(size_t)r]->ix, q, nq, p, K, o_ids, o_sc, nullptr, o_cnt, s, false, glob_len, o_gno);
    };
}

// queries / outputs on devices[0], enqueued on `stream`; never synchronises.  A zero query under cosine is NOT detected
// here (its row of results is undefined); the host entry point checks before it enqueues.
int cm_ivf_sharded_search_device(cm_ivf_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                 int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev,
                                 void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivf_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ls.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ls.dim, dim);
    if (nq <= 0) return CM_OK;
    return h->ls.search_device(ivfs_search_fn(h), queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_counts_dev,
                               (cudaStream_t)stream);
}

// nq independent searchSingleQuery calls (ivf_index_search.go:217-322) against the whole list-sharded index
int cm_ivf_sharded_search(cm_ivf_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivf_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ls.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ls.dim, dim);
    if (nq <= 0) return CM_OK;
    int rc = h->ls.search_host(ivfs_search_fn(h), queries, nq, p, out_stride, out_ids, out_scores, out_counts);
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t r = 0; r < h->shard.size(); r++) h->ls.last_scanned[r] = cm_ivf_last_scanned(h->shard[r]);
    cudaSetDevice(prev);
    return rc;
}

// vectors each shard scanned in the last host search (all queries): the balance of the list assignment
int cm_ivf_sharded_last_scanned(const cm_ivf_sharded *h, int64_t *per_shard) {
    if (!h || !per_shard) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    for (size_t r = 0; r < h->shard.size(); r++) per_shard[r] = h->ls.last_scanned[r];
    return CM_OK;
}

}  // extern "C"
