// pq.cu -- PQIndex and IVFPQIndex device state, search and the cm_pq_* / cm_ivfpq_* entry points.
//
// Replaces pqIndexSearch.searchSingleQuery (pq_index_search.go:218-325) and
// ivfpqIndexSearch.searchSingleQuery / computeDistanceTables / asymmetricDistance
// (ivfpq_index_search.go:231-390): per-query (PQ) or per-probe residual (IVFPQ) lookup tables
//   LUT[m][c] = sum_j (r[m*dsub+j] - codebook[m][c*dsub+j])^2     sequential fp32 (K7, K9)
// and the ADC scan  dist = float32(sqrt(float64(sum_m LUT[m][code[m]])))  sequential over m (K8),
// delete / document filter / threshold, full sort of the candidates, top k -- bit-identical.
//
// Device layout (replaces codes [][]uint8 / lists [][]CompressedVector of heap slices):
//   codebooks f32 [M][Ksub][dsub]
//   codes     u8  [cap][M]      arrival order;  ids u32 [cap];  deleted u8 [cap]
//   IVFPQ adds: coarse FlatIndex of the nlist centroids (raw), members u32 [n] (store positions
//   grouped by list, CSR), list_off i64 [nlist+1] and codes_by_list u8 [n][M] (the codes in CSR order).
// One kernel, adc_scan_kernel, serves both: a CTA owns one (query, probe) pair and a slice of its
// codes; it builds the pair's LUT in shared memory straight from the codebooks (only the
// 256 entries per sub-quantiser a uint8 code can address, pq_index.go:469), scans, and keeps its
// K best keys (score, candidate number).  merge_topk_kernel + adc_emit_kernel finish the query.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <unordered_set>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "kmeans.cuh"
#include "list_shards.cuh"
#include "select.cuh"
#include "wire.cuh"

namespace cm {

static constexpr int ADC_THREADS = 256;
static constexpr int ADC_CHUNK = 4096;      // fewest codes a CTA is given when a pair is sliced
static constexpr int ADC_RING_RUNS = 2;     // independent row pipelines per warp of the ring scan
static constexpr int ADC_BIGK_PART = 3072;  // slice length of the big-k path: a multiple of every R * ADC_THREADS (256 .. 1024)

struct CodeStore {
    int M = 0;
    int64_t n = 0, cap = 0;
    uint8_t *codes = nullptr;
    uint32_t *ids = nullptr;
    uint8_t *deleted = nullptr;
    std::vector<uint32_t> ids_host;
    std::unordered_set<uint32_t> deleted_ids;
    int64_t n_deleted_rows = 0;

    ~CodeStore() { cudaFree(codes); cudaFree(ids); cudaFree(deleted); }
    int reserve(int64_t want) {
        if (want <= cap) return CM_OK;
        int64_t ncap = cap ? cap : 4096;
        while (ncap < want) ncap = ncap + ncap / 2 + 4096;
        uint8_t *nc = nullptr, *nd = nullptr;
        uint32_t *ni = nullptr;
        CM_CUDA(cudaMalloc(&nc, (size_t)ncap * M));
        CM_CUDA(cudaMalloc(&ni, (size_t)ncap * 4));
        CM_CUDA(cudaMalloc(&nd, (size_t)ncap));
        CM_CUDA(cudaMemset(nd, 0, (size_t)ncap));
        if (n > 0) {
            CM_CUDA(cudaMemcpy(nc, codes, (size_t)n * M, cudaMemcpyDeviceToDevice));
            CM_CUDA(cudaMemcpy(ni, ids, (size_t)n * 4, cudaMemcpyDeviceToDevice));
            CM_CUDA(cudaMemcpy(nd, deleted, (size_t)n, cudaMemcpyDeviceToDevice));
        }
        cudaFree(codes); cudaFree(ids); cudaFree(deleted);
        codes = nc; ids = ni; deleted = nd; cap = ncap;
        return CM_OK;
    }
    // bookkeeping after `m` codes were written at [n, n+m) by an encode kernel
    int commit(const uint32_t *ids_h, int64_t m, cudaStream_t st) {
        std::vector<uint8_t> del((size_t)m, 0);
        bool any = false;
        for (int64_t i = 0; i < m; i++)
            if (!deleted_ids.empty() && deleted_ids.count(ids_h[i])) { del[(size_t)i] = 1; any = true; n_deleted_rows++; }
        CM_CUDA(cudaMemcpyAsync(ids + n, ids_h, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        if (any) CM_CUDA(cudaMemcpyAsync(deleted + n, del.data(), (size_t)m, cudaMemcpyHostToDevice, st));
        else CM_CUDA(cudaMemsetAsync(deleted + n, 0, (size_t)m, st));
        CM_CUDA(cudaStreamSynchronize(st));
        ids_host.insert(ids_host.end(), ids_h, ids_h + m);
        n += m;
        return CM_OK;
    }
    int remove(uint32_t id) {      // pq_index.go:299-330 / ivfpq_index.go:330-360: soft delete
        std::vector<int64_t> hits;
        for (int64_t i = 0; i < n; i++)
            if (ids_host[(size_t)i] == id) hits.push_back(i);
        if (hits.empty()) return fail(CM_ERR_NOT_FOUND, "vector with ID %u not found", id);
        if (deleted_ids.count(id)) return fail(CM_ERR_NOT_FOUND, "vector with ID %u already deleted", id);
        deleted_ids.insert(id);
        uint8_t one = 1;
        for (int64_t i : hits) CM_CUDA(cudaMemcpy(deleted + i, &one, 1, cudaMemcpyHostToDevice));
        n_deleted_rows += (int64_t)hits.size();
        return CM_OK;
    }
    // drop deleted rows keeping order; new_pos[i] = new position or -1
    int flush(std::vector<int64_t> *new_pos_out) {
        std::vector<int64_t> new_pos((size_t)n, -1);
        std::vector<int64_t> keep;
        for (int64_t i = 0; i < n; i++)
            if (!deleted_ids.count(ids_host[(size_t)i])) { new_pos[(size_t)i] = (int64_t)keep.size(); keep.push_back(i); }
        int64_t m = (int64_t)keep.size();
        if (m < n) {
            std::vector<uint8_t> hc((size_t)n * M), nc((size_t)std::max<int64_t>(m, 1) * M);
            CM_CUDA(cudaMemcpy(hc.data(), codes, (size_t)n * M, cudaMemcpyDeviceToHost));
            std::vector<uint32_t> ni((size_t)m);
            for (int64_t i = 0; i < m; i++) {
                memcpy(&nc[(size_t)i * M], &hc[(size_t)keep[(size_t)i] * M], (size_t)M);
                ni[(size_t)i] = ids_host[(size_t)keep[(size_t)i]];
            }
            if (m > 0) {
                CM_CUDA(cudaMemcpy(codes, nc.data(), (size_t)m * M, cudaMemcpyHostToDevice));
                CM_CUDA(cudaMemcpy(ids, ni.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
            }
            CM_CUDA(cudaMemset(deleted, 0, (size_t)cap));
            ids_host.swap(ni);
            n = m;
        }
        n_deleted_rows = 0;
        deleted_ids.clear();
        if (new_pos_out) new_pos_out->swap(new_pos);
        return CM_OK;
    }
};

struct PQCore {
    int dim = 0, metric = 0, M = 0, nbits = 0, Ksub = 0, dsub = 0, device = 0, ld = 0;
    bool trained = false;
    float *codebooks = nullptr;     // [M][Ksub][dsub]
    CodeStore store;
    // IVFPQ only
    int nlist = 0;
    FlatIndex coarse;
    std::vector<std::vector<uint32_t>> lists;
    std::vector<int32_t> list_of;
    uint32_t *members = nullptr;
    long long *list_off = nullptr;
    int64_t members_cap = 0;
    uint8_t *codes_by_list = nullptr;   // the codes again, list by list: a probed list is one contiguous stream.  M % 16 == 0: every
                                        // list starts on a 32-row boundary (tile_off) and each 32-row tile is stored word-major,
                                        // [M/16][32 rows][16 bytes], so a warp's 16-byte loads of 32 rows are one 512-byte run
    long long *tile_off = nullptr;      // [nlist+1] padded row offset of every list in codes_by_list (multiples of 32); M % 16 == 0 only
                                        // (ring layout: the list's first pipeline step)
    int64_t tiled_rows = 0;             // rows of codes_by_list including the padding
    int64_t cbl_bytes = 0;              // bytes allocated for codes_by_list
    float *codebooks_t = nullptr;       // dsub == 8: [M][2][Ksub] float4 -- the table build reads a codeword half per lane, coalesced
    // Ring layout (adc_ring_kernel; 8-bit codes, M = 32 * PER): lane l of a warp owns sub-quantisers PER*l .. PER*l+PER-1.
    // codes_by_list holds, per list, one 128-byte line per pipeline step t: lane l's 4 bytes = the PER code bytes of row t - l
    // (16-byte groups of four steps: [(t / 4)][lane][t % 4][4 bytes]); codebooks_r [c][j][dsub / 4][lane] float4.
    bool ring_layout = false;           // what codes_by_list currently holds
    float *codebooks_r = nullptr;
    bool cbt_dirty = true;
    bool csr_dirty = true;
    std::vector<int64_t> sizes_desc;
    std::mutex csr_mu;
    unsigned long long *scanned_total = nullptr;   // device: codes scanned by the last IVFPQ search (all queries)

    ~PQCore() {
        cudaFree(codebooks); cudaFree(members); cudaFree(list_off); cudaFree(scanned_total); cudaFree(codes_by_list);
        cudaFree(tile_off); cudaFree(codebooks_t); cudaFree(codebooks_r);
    }
    bool tiled() const { return (M & 15) == 0; }
    static bool ring_combo(int per, int d4) {       // adc_ring_kernel instantiations
        return (per == 1 && (d4 == 2 || d4 == 6)) || (per == 2 && d4 >= 1 && d4 <= 3) || (per == 3 && (d4 == 1 || d4 == 2)) ||
               (per == 4 && (d4 == 1 || d4 == 2));
    }
    bool ring_wanted() const {
        if (Ksub != 256 || (M & 31) != 0 || (dsub & 3) != 0 || !ring_combo(M / 32, dsub / 4)) return false;
        if (const char *e = getenv("COMET_B200_ADC_RING")) return atoi(e) != 0;
        return true;
    }
    int sync_codebooks_t(cudaStream_t st);
    int lut_entries() const { return Ksub < 256 ? Ksub : 256; }
    int sync_csr(cudaStream_t st);
};

// codes_by_list from the store: CSR entry i (list l, j-th member) -> 16-byte word k of its code
//   M % 16 == 0: tiled -- row p = tile_off[l] + j, word k at ((p / 32) * 32 * M) + k * 512 + (p % 32) * 16
//   else       : linear, codes_by_list[i] = codes[members[i]]
__global__ void gather_codes_kernel(const uint8_t *__restrict__ codes, const uint32_t *__restrict__ members, long long n,
                                    int M, const long long *__restrict__ list_off, const long long *__restrict__ tile_off, int nlist,
                                    uint8_t *__restrict__ out) {
    if ((M & 15) == 0) {
        const int w = M / 16;
        long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        if (t >= n * w) return;
        long long i = t / w;
        int k = (int)(t - i * w);
        int lo = 0, hi = nlist;                    // list of CSR entry i: largest l with list_off[l] <= i
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (list_off[mid] <= i) lo = mid; else hi = mid;
        }
        const long long p = tile_off[lo] + (i - list_off[lo]);
        const size_t dst = (size_t)(p >> 5) * 32 * M + (size_t)k * 512 + (size_t)(p & 31) * 16;
        *reinterpret_cast<uint4 *>(out + dst) = __ldg(reinterpret_cast<const uint4 *>(codes + (size_t)members[i] * M) + k);
    } else {
        long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        if (t >= n * M) return;
        long long i = t / M;
        out[t] = codes[(size_t)members[i] * M + (t - i * M)];
    }
}

// codebooks [M][Ksub][8] -> [M][2][Ksub] float4
__global__ void transpose_codebooks_kernel(const float4 *__restrict__ cb, int M, int Ksub, float4 *__restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)M * Ksub * 2) return;
    const int h = (int)(t & 1);
    const long long mc = t >> 1;                   // m * Ksub + c
    const long long m = mc / Ksub, c = mc - m * Ksub;
    out[(m * 2 + h) * Ksub + c] = cb[t];
}

// ring layout of the codes: CSR entry i (list lo, row r), lane l -> the PER code bytes of sub-quantisers PER*l.. at step r + l
__global__ void gather_codes_ring_kernel(const uint8_t *__restrict__ codes, const uint32_t *__restrict__ members, long long n,
                                         int M, int per, const long long *__restrict__ list_off,
                                         const long long *__restrict__ step_off, int nlist, uint8_t *__restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * 32) return;
    const long long i = t >> 5;
    const int l = (int)(t & 31);
    int lo = 0, hi = nlist;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (list_off[mid] <= i) lo = mid; else hi = mid;
    }
    const long long st = step_off[lo] + (i - list_off[lo]) + l;
    const uint8_t *src = codes + (size_t)(members ? members[i] : (uint32_t)i) * M + (size_t)per * l;
    uint32_t w = 0;
    for (int j = 0; j < per; j++) w |= (uint32_t)src[j] << (8 * j);
    *reinterpret_cast<uint32_t *>(out + ((size_t)(st >> 2) * 32 + l) * 16 + (size_t)(st & 3) * 4) = w;
}

// codebooks [M][256][dsub] -> [c][j][dsub / 4][lane] float4 with m = per * lane + j
__global__ void ring_codebooks_kernel(const float4 *__restrict__ cb, int per, int d4n, float4 *__restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = 256ll * per * d4n * 32;
    if (t >= total) return;
    const int l = (int)(t & 31);
    long long u = t >> 5;
    const int d4 = (int)(u % d4n); u /= d4n;
    const int j = (int)(u % per);
    const long long c = u / per;
    const long long m = (long long)per * l + j;
    out[t] = cb[(m * 256 + c) * d4n + d4];
}

int PQCore::sync_codebooks_t(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(csr_mu);
    if (!codebooks) return CM_OK;
    const bool need_r = ring_wanted();
    if (!cbt_dirty && (!need_r || codebooks_r)) return CM_OK;
    if (need_r || codebooks_r) {
        if (!codebooks_r) CM_CUDA(cudaMalloc(&codebooks_r, (size_t)M * Ksub * dsub * 4));
        const long long work = (long long)M * Ksub * (dsub / 4);
        ring_codebooks_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(codebooks), M / 32, dsub / 4,
                                                                             reinterpret_cast<float4 *>(codebooks_r));
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    if (dsub != 8) {
        CM_CUDA(cudaStreamSynchronize(st));
        cbt_dirty = false;
        return CM_OK;
    }
    if (!codebooks_t) CM_CUDA(cudaMalloc(&codebooks_t, (size_t)M * Ksub * 8 * 4));
    const long long work = (long long)M * Ksub * 2;
    transpose_codebooks_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(codebooks), M, Ksub,
                                                                              reinterpret_cast<float4 *>(codebooks_t));
    count_launch();
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(st));
    cbt_dirty = false;
    return CM_OK;
}

int PQCore::sync_csr(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(csr_mu);
    const bool ring = ring_wanted();
    if (!csr_dirty && ring == ring_layout) return CM_OK;
    if (nlist == 0) {
        // PQ: the store is one list.  Ring form: a pre-skewed copy of all codes (positions = candidate numbers)
        if (ring && store.n > 0) {
            const long long steps = ((long long)store.n + 32 + 31) & ~31ll;
            if (steps * 128 > cbl_bytes || !codes_by_list) {
                cudaFree(codes_by_list);
                codes_by_list = nullptr;
                cbl_bytes = (steps + steps / 2) * 128;
                CM_CUDA(cudaMalloc(&codes_by_list, (size_t)cbl_bytes));
            }
            if (!list_off) CM_CUDA(cudaMalloc(&list_off, 2 * sizeof(long long)));
            if (!tile_off) CM_CUDA(cudaMalloc(&tile_off, 2 * sizeof(long long)));
            const long long off[2] = {0, (long long)store.n}, toff[2] = {0, steps};
            CM_CUDA(cudaMemcpyAsync(list_off, off, sizeof(off), cudaMemcpyHostToDevice, st));
            CM_CUDA(cudaMemcpyAsync(tile_off, toff, sizeof(toff), cudaMemcpyHostToDevice, st));
            CM_CUDA(cudaMemsetAsync(codes_by_list, 0, (size_t)steps * 128, st));
            const long long work = (long long)store.n * 32;
            gather_codes_ring_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(store.codes, nullptr, (long long)store.n, M, M / 32,
                                                                                     list_off, tile_off, 1, codes_by_list);
            count_launch();
            CM_CUDA(cudaGetLastError());
            CM_CUDA(cudaStreamSynchronize(st));      // off / toff live on this stack
        }
        csr_dirty = false;
        ring_layout = ring && store.n > 0;
        return CM_OK;
    }
    int64_t n = store.n;
    std::vector<uint32_t> flat;
    flat.reserve((size_t)n);
    std::vector<long long> off((size_t)nlist + 1, 0), toff((size_t)nlist + 1, 0);
    sizes_desc.assign((size_t)nlist, 0);
    for (int l = 0; l < nlist; l++) {
        off[(size_t)l] = (long long)flat.size();
        flat.insert(flat.end(), lists[(size_t)l].begin(), lists[(size_t)l].end());
        sizes_desc[(size_t)l] = (int64_t)lists[(size_t)l].size();
        // tiled: whole 32-row tiles; ring: len + 31 pipeline steps (row r reaches lane 31 at step r + 31), whole 32-step chunks
        const long long len = (long long)lists[(size_t)l].size();
        toff[(size_t)l + 1] = toff[(size_t)l] + (ring ? ((len + 32 + 31) & ~31ll) : ((len + 31) & ~31ll));
    }
    off[(size_t)nlist] = (long long)flat.size();
    const int64_t rows_needed = (tiled() || ring) ? std::max<int64_t>(toff[(size_t)nlist], 32) : n;
    const int64_t row_bytes = ring ? 128 : M;
    if (n > members_cap || !members || rows_needed * row_bytes > cbl_bytes || !codes_by_list) {
        cudaFree(members);
        cudaFree(codes_by_list);
        codes_by_list = nullptr;
        members_cap = std::max<int64_t>(n + n / 2, 1024);
        tiled_rows = std::max<int64_t>(rows_needed + rows_needed / 2, 1024);
        cbl_bytes = tiled_rows * row_bytes;
        CM_CUDA(cudaMalloc(&members, (size_t)members_cap * sizeof(uint32_t)));
        CM_CUDA(cudaMalloc(&codes_by_list, (size_t)cbl_bytes));
    }
    if (!list_off) CM_CUDA(cudaMalloc(&list_off, (size_t)(nlist + 1) * sizeof(long long)));
    if (!tile_off) CM_CUDA(cudaMalloc(&tile_off, (size_t)(nlist + 1) * sizeof(long long)));
    std::sort(sizes_desc.begin(), sizes_desc.end(), std::greater<int64_t>());
    if (!flat.empty()) CM_CUDA(cudaMemcpyAsync(members, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, st));
    CM_CUDA(cudaMemcpyAsync(list_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    CM_CUDA(cudaMemcpyAsync(tile_off, toff.data(), toff.size() * 8, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        // the padding rows / steps are read (never emitted)
        if (tiled() || ring) CM_CUDA(cudaMemsetAsync(codes_by_list, 0, (size_t)rows_needed * row_bytes, st));
        if (ring) {
            const long long work = (long long)n * 32;
            gather_codes_ring_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(store.codes, members, (long long)n, M, M / 32, list_off,
                                                                                     tile_off, nlist, codes_by_list);
        } else {
            const long long work = tiled() ? (long long)n * (M / 16) : (long long)n * M;
            gather_codes_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(store.codes, members, (long long)n, M, list_off, tile_off, nlist,
                                                                                codes_by_list);
        }
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    CM_CUDA(cudaStreamSynchronize(st));
    csr_dirty = false;
    ring_layout = ring;
    return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// encode: PQIndex.encode (pq_index.go:439-473) / IVFPQIndex.encodeResidual (ivfpq_index.go:467-500)
// one thread per (row, sub-quantiser): first minimum over ALL Ksub centroids, then uint8(minIdx)
// ------------------------------------------------------------------------------------------------
template <bool FMA>
__global__ void pq_encode_kernel(const float *__restrict__ rows, int ld, long long n, int M, int Ksub, int dsub,
                                 const float *__restrict__ codebooks, const float *__restrict__ centroids, int cld,
                                 const long long *__restrict__ row_list, uint8_t *__restrict__ codes) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * M) return;
    long long row = t / M;
    int m = (int)(t - row * M);
    const float *v = rows + (size_t)row * ld + (size_t)m * dsub;
    const float *cen = centroids ? centroids + (size_t)row_list[row] * cld + (size_t)m * dsub : nullptr;
    const float *cb = codebooks + (size_t)m * Ksub * dsub;
    float best = INFINITY;
    int best_i = 0;
    for (int c = 0; c < Ksub; c++) {
        float dist = 0.0f;
        for (int j = 0; j < dsub; j++) {
            float r = cen ? __fsub_rn(v[j], cen[j]) : v[j];
            dist = l2_step<FMA>(dist, r, cb[(size_t)c * dsub + j]);
        }
        if (dist < best) { best = dist; best_i = c; }
    }
    codes[(size_t)row * M + m] = (uint8_t)best_i;
}

// ------------------------------------------------------------------------------------------------
// ADC scan.  blockIdx.y = pair = query * nprobes + probe, blockIdx.x = slice of the pair's codes.
//   members == nullptr (PQ): the pair's codes are store positions [0, n), candidate number = position
//   else (IVFPQ): list l = probe_list[pair]; codes at members[list_off[l] + j]; candidate number =
//   q_off[query][probe] + j (the order of the reference's append loop)
// A CTA builds the pair's table ONCE and then walks its whole slice of the codes (slices are sized on
// the host so that a launch has a few waves of CTAs: 1M codes x 128 queries is 10 slices per query,
// not 245 rebuilds of the same table).  The walk is in rounds of R code rows per thread: all R x MW
// 16-byte code loads of a round are issued together (the rows of the NEXT round were prefetched to
// L2 a round earlier; a pair's rows are one contiguous stream -- IVFPQ scans a list-ordered copy of
// the codes, rebuilt with the CSR after Add / Flush), the R table-sum chains are interleaved
// (each chain is the reference's sequential m = 0..M-1 order), and the CTA meets at ONE barrier per
// round, which also decides whether the candidate buffer must be compacted.
//   MW = M / 16 (code bytes per row in 16-byte words) or 0 for the generic byte-wise path.
// ------------------------------------------------------------------------------------------------
template <int MW> struct AdcRows { static constexpr int R = MW == 0 ? 1 : (MW <= 4 ? 4 : (MW <= 6 ? 3 : 2)); };

__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// one thread asks for `bytes` (a multiple of 16) contiguous bytes to be brought into L2
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <bool FMA, int MW>
__global__ void __launch_bounds__(ADC_THREADS, 2) adc_scan_kernel(
    const float *__restrict__ queries, int ld, int dim, int M, int Ksub, int dsub, int lut_n,
    const float *__restrict__ codebooks, const uint8_t *__restrict__ codes, long long n_store,
    const float *__restrict__ centroids, int cld, const long long *__restrict__ probe_list,
    const long long *__restrict__ q_off, const long long *__restrict__ list_off, const uint32_t *__restrict__ members,
    int nprobes, const uint8_t *__restrict__ skip, float threshold, int K, int C, int n_slices,
    uint64_t *__restrict__ part_keys, int *__restrict__ part_counts, const long long *__restrict__ tile_off,
    const float4 *__restrict__ codebooks_t) {
    constexpr int R = AdcRows<MW>::R;
    constexpr int T = ADC_THREADS;
    extern __shared__ __align__(16) uint8_t smem[];
    float *lut = reinterpret_cast<float *>(smem);                       // [M][lut_n]
    float *res = lut + (size_t)M * lut_n;                               // [dim] query (residual)
    uint64_t *buf = reinterpret_cast<uint64_t *>(res + ((dim + 3) & ~3));   // [C]
    __shared__ int cnt;
    __shared__ uint64_t tau;
    __shared__ int sel_hist[256];
    __shared__ uint32_t sel_red[4];
    const CtaBarrier bar;
    const int tid = threadIdx.x;
    const long long pair = blockIdx.y;
    const int q = (int)(pair / nprobes), pr = (int)(pair % nprobes);
    long long len, order0;
    const uint32_t *mem = nullptr;
    long long list = -1;
    if (members) {
        list = probe_list[pair];
        len = list_off[list + 1] - list_off[list];
        mem = members + list_off[list];
        order0 = q_off[(size_t)q * (nprobes + 1) + pr];
    } else {
        len = n_store;
        order0 = 0;
    }
    // this CTA's slice [c0, c1): equal shares of the pair's codes, rounded up to whole rounds
    long long per = (len + n_slices - 1) / n_slices;
    per = (per + (long long)R * T - 1) / ((long long)R * T) * ((long long)R * T);
    const long long c0 = (long long)blockIdx.x * per;
    const size_t part = (size_t)pair * n_slices + blockIdx.x;
    if (c0 >= len) return;                                              // part_counts was zeroed by the host
    const long long c1 = min(len, c0 + per);
    if (tid == 0) { cnt = 0; tau = KEY_INF; }
    // query residual (ivfpq_index_search.go:285-296) or the query itself
    for (int j = tid; j < dim; j += T) {
        float v = queries[(size_t)q * ld + j];
        res[j] = centroids ? __fsub_rn(v, centroids[(size_t)list * cld + j]) : v;
    }
    __syncthreads();
    // lookup table in the reference's summation order
    if (dsub == 8 && lut_n == T) {
        // thread = table column c, four sub-quantisers in flight (independent chains); the residual
        // piece is a shared-memory broadcast, the codebook row a coalesced 32-byte load
        // codeword halves: [m][h][c] float4 in the transposed copy (a warp's load is one 512-byte run; from the reference
        // layout [m][c][8] the same two loads touch twice the lines)
        const float4 *cb4 = codebooks_t ? codebooks_t + tid : reinterpret_cast<const float4 *>(codebooks) + (size_t)tid * 2;
        const size_t m_stride4 = (size_t)Ksub * 2;
        const size_t h_stride4 = codebooks_t ? (size_t)Ksub : 1;
        int m = 0;
        for (; m + 4 <= M; m += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                a[u] = __ldg(cb4 + (size_t)(m + u) * m_stride4);
                b[u] = __ldg(cb4 + (size_t)(m + u) * m_stride4 + h_stride4);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 r0 = *reinterpret_cast<const float4 *>(res + (m + u) * 8);
                const float4 r1 = *reinterpret_cast<const float4 *>(res + (m + u) * 8 + 4);
                float d = 0.0f;
                d = l2_step<FMA>(d, r0.x, a[u].x); d = l2_step<FMA>(d, r0.y, a[u].y);
                d = l2_step<FMA>(d, r0.z, a[u].z); d = l2_step<FMA>(d, r0.w, a[u].w);
                d = l2_step<FMA>(d, r1.x, b[u].x); d = l2_step<FMA>(d, r1.y, b[u].y);
                d = l2_step<FMA>(d, r1.z, b[u].z); d = l2_step<FMA>(d, r1.w, b[u].w);
                lut[(m + u) * T + tid] = d;
            }
        }
        for (; m < M; m++) {
            const float4 a = __ldg(cb4 + (size_t)m * m_stride4), b = __ldg(cb4 + (size_t)m * m_stride4 + h_stride4);
            const float4 r0 = *reinterpret_cast<const float4 *>(res + m * 8);
            const float4 r1 = *reinterpret_cast<const float4 *>(res + m * 8 + 4);
            float d = 0.0f;
            d = l2_step<FMA>(d, r0.x, a.x); d = l2_step<FMA>(d, r0.y, a.y);
            d = l2_step<FMA>(d, r0.z, a.z); d = l2_step<FMA>(d, r0.w, a.w);
            d = l2_step<FMA>(d, r1.x, b.x); d = l2_step<FMA>(d, r1.y, b.y);
            d = l2_step<FMA>(d, r1.z, b.z); d = l2_step<FMA>(d, r1.w, b.w);
            lut[m * T + tid] = d;
        }
    } else {
        for (int e = tid; e < M * lut_n; e += T) {
            int m = e / lut_n, c = e - m * lut_n;
            const float *cb = codebooks + ((size_t)m * Ksub + c) * dsub;
            const float *r = res + (size_t)m * dsub;
            float dist = 0.0f;
            if ((dsub & 3) == 0) {                       // 16-byte codebook loads; same summation order
                const float4 *cb4 = reinterpret_cast<const float4 *>(cb);
                for (int j = 0; j < dsub / 4; j++) {
                    float4 v = __ldg(cb4 + j);
                    dist = l2_step<FMA>(dist, r[4 * j + 0], v.x);
                    dist = l2_step<FMA>(dist, r[4 * j + 1], v.y);
                    dist = l2_step<FMA>(dist, r[4 * j + 2], v.z);
                    dist = l2_step<FMA>(dist, r[4 * j + 3], v.w);
                }
            } else {
                for (int j = 0; j < dsub; j++) dist = l2_step<FMA>(dist, r[j], cb[j]);
            }
            lut[e] = dist;
        }
    }
    __syncthreads();

    const int limit = C - R * T;              // appends of one round always fit above this fill
    const int keep_mid = K + (limit - K) / 2; // keys a mid-stream compaction may leave (K <= keep_mid <= limit)
    int keep_end = 64;
    while (keep_end < K) keep_end <<= 1;
    if (keep_end > limit) keep_end = K;
    // the pair's code rows are one contiguous stream: the store itself (PQ) or the list's slice of the
    // list-ordered copy (IVFPQ; `codes` then points at codes_by_list)
    // tiled (tile_off != nullptr): the list starts at padded row tile_off[list]; row j, 16-byte word w sits at
    // (j / 32) * 32 * M + w * 512 + (j % 32) * 16 -- c0 and the round size are multiples of 32, so j % 32 is the lane
    const uint8_t *rows = codes + (mem ? (size_t)(tile_off ? tile_off[list] : list_off[list]) * M : 0);
    const int lane = tid & 31;
    for (long long base = c0; base < c1; base += (long long)R * T) {
        bool live[R];
        float dist[R];
        if (MW > 0) {
            uint32_t wd[R][MW > 0 ? MW * 4 : 1];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const long long j = base + (long long)r * T + tid;
                live[r] = j < c1;
                const long long js = live[r] ? j : c0 + lane;          // dead rows re-read the slice's first tile: never emitted
                const uint4 *c4 = tile_off ? reinterpret_cast<const uint4 *>(rows + (size_t)(js >> 5) * 32 * M) + lane
                                           : reinterpret_cast<const uint4 *>(rows + (size_t)js * M);
                const int wstride = tile_off ? 32 : 1;
#pragma unroll
                for (int w = 0; w < MW; w++) {
                    uint4 v = __ldg(c4 + w * wstride);
                    wd[r][w * 4 + 0] = v.x; wd[r][w * 4 + 1] = v.y; wd[r][w * 4 + 2] = v.z; wd[r][w * 4 + 3] = v.w;
                }
            }
            // next round's rows towards L2 while this round computes (one 128-byte line per 128 bytes of rows)
            {
                const long long c1p = tile_off ? ((c1 + 31) & ~31ll) : c1;
                const long long nb = (base + (long long)R * T) * M, ne = min(c1p, base + 2ll * R * T) * M;
                for (long long o = nb + (long long)tid * 128; o < ne; o += (long long)T * 128) prefetch_l2(rows + o);
            }
            float sum[R];
#pragma unroll
            for (int r = 0; r < R; r++) sum[r] = 0.0f;
            const float *lp = lut;
#pragma unroll
            for (int w = 0; w < MW * 4; w++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
#pragma unroll
                    for (int r = 0; r < R; r++) sum[r] = __fadd_rn(sum[r], lp[(wd[r][w] >> (8 * b)) & 0xffu]);
                    lp += lut_n;
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) dist[r] = __fsqrt_rn(sum[r]);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const long long j = base + (long long)r * T + tid;
                live[r] = j < c1;
                float sum = 0.0f;
                if (live[r]) {
                    if (tile_off) {      // tiled store, byte-wise walk: byte m of row j is in word m / 16 of its tile
                        const uint8_t *tile = rows + (size_t)(j >> 5) * 32 * M + (size_t)(j & 31) * 16;
                        for (int m = 0; m < M; m++) sum = __fadd_rn(sum, lut[m * lut_n + tile[(size_t)(m >> 4) * 512 + (m & 15)]]);
                    } else {
                        const uint8_t *code = rows + (size_t)j * M;
                        for (int m = 0; m < M; m++) sum = __fadd_rn(sum, lut[m * lut_n + code[m]]);
                    }
                }
                dist[r] = __fsqrt_rn(sum);
            }
        }
        // selection: append below the running bound; one barrier per round decides on compaction
        int top = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            bool ok = live[r];
            const long long j = base + (long long)r * T + tid;
            if (ok && skip != nullptr && skip[mem ? mem[j] : (uint32_t)j]) ok = false;
            if (ok && threshold > 0.0f && dist[r] > threshold) ok = false;
            if (ok) {
                uint64_t key = make_key(dist[r], (uint32_t)(order0 + j));
                if (key < tau) {
                    int slot = atomicAdd(&cnt, 1);
                    buf[slot] = key;
                    top = max(top, slot + 1);
                }
            }
        }
        // the thread that took the highest slot sees the final fill, so the OR is exact
        // (a radix-select compaction: no sort; it leaves up to keep_mid keys and a bound just above them)
        if (__syncthreads_or(top > limit)) {
            if (!compact_select(buf, C, K, keep_mid, &cnt, &tau, tid, T, sel_hist, sel_red))
                compact_topk(buf, C, K, &cnt, &tau, tid, T, bar);
        }
    }
    // the CTA's answer: exactly the K smallest -- select down to at most next_pow2(K) keys, sort those
    if (!compact_select(buf, C, K, keep_end, &cnt, &tau, tid, T, sel_hist, sel_red)) { /* ties: sort them all */ }
    compact_topk(buf, C, K, &cnt, &tau, tid, T, bar);
    int mcount = cnt;
    uint64_t *dst = part_keys + part * K;
    for (int i = tid; i < mcount; i += T) dst[i] = buf[i];
    if (tid == 0) part_counts[part] = mcount;
}

// ------------------------------------------------------------------------------------------------
// ADC scan, ring form (IVFPQ, 8-bit codes, M = 32 * PER).  The row-per-lane scan above pays ~4 shared-memory
// wavefronts per table read: 32 lanes, 32 random banks.  Here a LANE owns sub-quantisers (PER*lane ..
// PER*lane+PER-1) and their table rows live in bank `lane` only (lut[(j*256 + c)*32 + lane]), so every
// table read of a warp is one conflict-free wavefront.  A row's sum must still be the reference's
// sequential m = 0..M-1 chain (pq_index_search.go:300-311), so the rows flow through the lanes as a
// pipeline: at step t lane l takes the partial sum lane l-1 produced at step t-1 (one shuffle), adds
// its PER table values in order, and lane 31 holds the finished sum of row t - 31.  The codes are stored
// pre-skewed for this (PQCore::ring_layout): step t is one 128-byte line, lane l's word = its bytes of
// row t - l.  Each warp walks a contiguous run of the slice's rows (32 steps of fill, then one row per
// step); lane 31's sums are staged in shared memory and selected 32 at a time by the whole warp.
// The table build: lane = owner of the sub-quantisers, warp = table column, residual in registers,
// codewords from the ring copy of the codebooks (512-byte coalesced runs), stores conflict-free.
// ------------------------------------------------------------------------------------------------
// one table read of the ring scan: byte j of the packed code word selects the row; lut_lane = this lane's bank
// (table row (j, c) at lut_lane + j * 32 KB + c * 128 bytes)
__device__ __forceinline__ float ring_lut_read(const float *lut_lane, uint32_t word, int j) {
    uint32_t c;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(c) : "r"(word), "r"(0x4440u | (uint32_t)j));
    return *reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(lut_lane) + (size_t)j * 32768u + (c << 7));
}

template <bool FMA, int PER, int D4>
__global__ void __launch_bounds__(ADC_THREADS, 2) adc_ring_kernel(
    const float *__restrict__ queries, int ld, const float4 *__restrict__ cbr, const uint8_t *__restrict__ ring,
    const float *__restrict__ centroids, int cld, const long long *__restrict__ probe_list,
    const long long *__restrict__ q_off, const long long *__restrict__ list_off, const long long *__restrict__ step_off,
    const uint32_t *__restrict__ members, int nprobes, const uint8_t *__restrict__ skip, float threshold, int K, int C,
    int n_slices, uint64_t *__restrict__ part_keys, int *__restrict__ part_counts, unsigned long long *q_tau, int nq) {
    constexpr int T = ADC_THREADS, NW = T / 32, DS = D4 * 4, S = ADC_RING_RUNS;
    extern __shared__ __align__(16) uint8_t smem[];
    float *lut = reinterpret_cast<float *>(smem);                         // [PER][256][32 lanes]
    float *stage = lut + PER * 256 * 32;                                  // [NW][S][32] finished sums of a chunk
    uint64_t *buf = reinterpret_cast<uint64_t *>(stage + NW * S * 32);    // [C]
    __shared__ int cnt;
    __shared__ uint64_t tau;
    __shared__ int sel_hist[256];
    __shared__ uint32_t sel_red[4];
    const CtaBarrier bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTAs run probe-major: every query's nearest list first.  By the time the far lists of a query are scanned, q_tau
    // holds the K-th key of a near one and next to nothing of a far list gets past it.
    const int q = (int)(blockIdx.y % nq), pr = (int)(blockIdx.y / nq);
    const long long pair = (long long)q * nprobes + pr;
    // PQ (members == nullptr): one list, the store itself; candidate number = position
    const long long list = members ? probe_list[pair] : 0;
    const long long len = list_off[list + 1] - list_off[list];
    const uint32_t *mem = members ? members + list_off[list] : nullptr;
    const long long order0 = members ? q_off[(size_t)q * (nprobes + 1) + pr] : 0;
    long long per = (len + n_slices - 1) / n_slices;
    per = (per + T * S - 1) / (T * S) * (T * S);
    const long long c0 = (long long)blockIdx.x * per;
    const size_t part = (size_t)pair * n_slices + blockIdx.x;
    if (c0 >= len) return;                                                // part_counts was zeroed by the host
    const long long c1 = min(len, c0 + per);
    if (tid == 0) {                                                       // the bound the query's earlier CTAs found, if any
        cnt = 0;
        tau = q_tau != nullptr ? (uint64_t)*reinterpret_cast<volatile unsigned long long *>(q_tau + q) : KEY_INF;
    }

    // this warp's rows of the slice: S runs [a_s, b_s) of equal length (whole chunks of 32), walked as S
    // independent pipelines -- a pipeline step is a dependent shuffle -> add chain, two of them interleave
    long long pw = (c1 - c0 + NW * S - 1) / (NW * S);
    pw = (pw + 31) & ~31ll;                   // rows per run
    const int nchunks = (int)(pw >> 5) + 1;   // 31 steps of fill, then a row per step
    const uint8_t *list_bytes = ring + (size_t)step_off[list] * 128;
    const uint4 *lines = reinterpret_cast<const uint4 *>(list_bytes) + lane;
    const int gmax = (int)((len + 31) >> 2);  // last 4-step group of the list that holds codes (512 bytes a group)
    long long ra[S], rb[S];
    int G[S];
    uint4 cw[S][4];
#pragma unroll
    for (int st = 0; st < S; st++) {
        ra[st] = c0 + ((long long)warp * S + st) * pw;
        rb[st] = min(c1, ra[st] + pw);
        G[st] = (int)min(ra[st] >> 2, (long long)gmax);
        // the first code lines are requested before the table build: their latency hides under it
#pragma unroll
        for (int u = 0; u < 4; u++) cw[st][u] = __ldg(lines + (size_t)min(G[st] + u, gmax) * 32);
        // and the run's first two chunks (4 KB each) towards L2
        {
            const int g1 = min(G[st] + 16, gmax + 1);
            if (lane == 0 && g1 > G[st]) prefetch_l2_bulk(list_bytes + (size_t)G[st] * 512, (uint32_t)(g1 - G[st]) * 512u);
        }
    }

    // ---- table: this lane's residual pieces (ivfpq_index_search.go:285-296), then 256 / NW columns per warp ----
    {
        // the residual, fetched once per CTA (coalesced) and handed to the lanes through the candidate buffer's space
        float *res_s = reinterpret_cast<float *>(buf);
        for (int i = tid; i < 32 * PER * DS; i += T)
            res_s[i] = centroids ? __fsub_rn(__ldg(queries + (size_t)q * ld + i), __ldg(centroids + (size_t)list * cld + i))
                                 : __ldg(queries + (size_t)q * ld + i);
        __syncthreads();
        float res[PER][DS];
#pragma unroll
        for (int j = 0; j < PER; j++)
#pragma unroll
            for (int d4 = 0; d4 < D4; d4++) {
                const float4 v = *reinterpret_cast<const float4 *>(res_s + (lane * PER + j) * DS + 4 * d4);
                res[j][4 * d4 + 0] = v.x; res[j][4 * d4 + 1] = v.y; res[j][4 * d4 + 2] = v.z; res[j][4 * d4 + 3] = v.w;
            }
        constexpr int U = 4;                                              // columns in flight
#pragma unroll
        for (int j = 0; j < PER; j++) {
            for (int cg = 0; cg < 256 / NW; cg += U) {
                float4 a[U][D4];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int c = warp + NW * (cg + u);
#pragma unroll
                    for (int d4 = 0; d4 < D4; d4++) a[u][d4] = __ldg(cbr + ((size_t)(c * PER + j) * D4 + d4) * 32 + lane);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int c = warp + NW * (cg + u);
                    float d = 0.0f;
#pragma unroll
                    for (int d4 = 0; d4 < D4; d4++) {
                        d = l2_step<FMA>(d, res[j][4 * d4 + 0], a[u][d4].x);
                        d = l2_step<FMA>(d, res[j][4 * d4 + 1], a[u][d4].y);
                        d = l2_step<FMA>(d, res[j][4 * d4 + 2], a[u][d4].z);
                        d = l2_step<FMA>(d, res[j][4 * d4 + 3], a[u][d4].w);
                    }
                    lut[(j * 256 + c) * 32 + lane] = d;
                }
            }
        }
    }
    __syncthreads();

    const int round_keys = T * S;             // a round (one chunk of every run) appends at most this many keys
    const int limit = C - round_keys;
    // keys a mid-stream compaction may leave: barely more than K, so that the bound it sets lets the rest of a list
    // through rarely (a bin boundary of the radix select's second pass nearly always falls inside the margin)
    const int keep_mid = min(K + (limit - K) / 2, K + max(16, K / 8));
    const float *lut_s = lut + lane;          // this lane's bank
    float sum_in[S];                          // the partial sum this lane produced at the previous step
#pragma unroll
    for (int st = 0; st < S; st++) sum_in[st] = 0.0f;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int ch = 0; ch < nchunks; ch++) {
        // the bound the query's other CTAs have found so far (one broadcast load, used when the chunk is selected)
        uint64_t tau_q = KEY_INF;
        if (q_tau != nullptr) tau_q = *reinterpret_cast<volatile unsigned long long *>(q_tau + q);
        // the chunk after next towards L2 (4 KB per run, one bulk prefetch): the register ring below then only has to
        // cover an L2 hit
#pragma unroll
        for (int st = 0; st < S; st++) {
            const int g0 = G[st] + 16, g1 = min(g0 + 8, gmax + 1);
            if (lane == 0 && g1 > g0) prefetch_l2_bulk(list_bytes + (size_t)g0 * 512, (uint32_t)(g1 - g0) * 512u);
        }
#pragma unroll
        for (int g = 0; g < 8; g++) {
            uint32_t wd[S][4];
#pragma unroll
            for (int st = 0; st < S; st++) {
                const uint4 w4 = cw[st][g & 3];
                cw[st][g & 3] = __ldg(lines + (size_t)min(G[st] + g + 4, gmax) * 32);
                wd[st][0] = w4.x; wd[st][1] = w4.y; wd[st][2] = w4.z; wd[st][3] = w4.w;
            }
            float o[S][4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
#pragma unroll
                for (int st = 0; st < S; st++) {
                    float in = __shfl_up_sync(0xffffffffu, sum_in[st], 1);
                    if (lane == 0) in = 0.0f;
                    float acc = in;
#pragma unroll
                    for (int j = 0; j < PER; j++) acc = __fadd_rn(acc, ring_lut_read(lut_s, wd[st][e], j));
                    sum_in[st] = acc;
                    o[st][e] = acc;
                }
            }
            if (lane == 31) {
#pragma unroll
                for (int st = 0; st < S; st++)
                    reinterpret_cast<float4 *>(stage + (warp * S + st) * 32)[g] = make_float4(o[st][0], o[st][1], o[st][2], o[st][3]);
            }
        }
        __syncwarp();
        const uint64_t bound = min((uint64_t)tau, tau_q);
        bool full = false;                    // did an append of this warp end above the compaction mark?
#pragma unroll
        for (int st = 0; st < S; st++) {
            G[st] += 8;
            const float sum = stage[(warp * S + st) * 32 + lane];
            // lane i holds the row that left the pipeline at step a + 32 ch + i: row a + 32 ch + i - 31
            const long long j = ra[st] + 32ll * ch + lane - 31;
            bool ok = j >= ra[st] && j < rb[st];
            const float dist = __fsqrt_rn(sum);
            if (ok && skip != nullptr && skip[mem ? mem[j] : (uint32_t)j]) ok = false;
            if (ok && threshold > 0.0f && dist > threshold) ok = false;
            uint64_t key = 0;
            if (ok) {
                key = make_key(dist, (uint32_t)(order0 + j));
                ok = key < bound;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, ok);
            if (bal != 0u) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&cnt, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (ok) buf[slot0 + __popc(bal & lt_mask)] = key;
                full |= slot0 + __popc(bal) > limit;
            }
        }
        __syncwarp();
        // one barrier per round: the warp whose append ended highest saw the final fill, so the OR is exact.  Above the
        // mark the buffer may not hold another round (round_keys appends at most): compact, and tell the query's other CTAs
        if (ch + 1 < nchunks && __syncthreads_or(full)) {
            if (!compact_select(buf, C, K, keep_mid, &cnt, &tau, tid, T, sel_hist, sel_red))
                compact_topk(buf, C, K, &cnt, &tau, tid, T, bar);
            // K of this CTA's keys are under tau: no key at or above it is among the query's K best, in any list
            if (q_tau != nullptr && tid == 0) atomicMin(q_tau + q, (unsigned long long)tau);
        }
    }
    // the CTA's answer: exactly the K smallest, in any order (merge_topk_kernel / merge_topk_bigk order them); only
    // score ties across the K-th place need the sort
    if (!compact_select(buf, C, K, K, &cnt, &tau, tid, T, sel_hist, sel_red)) compact_topk(buf, C, K, &cnt, &tau, tid, T, bar);
    const int mcount = cnt;
    if (q_tau != nullptr && tid == 0 && mcount >= K) atomicMin(q_tau + q, (unsigned long long)tau);
    uint64_t *dst = part_keys + part * K;
    for (int i = tid; i < mcount; i += T) dst[i] = buf[i];
    if (tid == 0) part_counts[part] = mcount;
}

// candidate numbers of the final lists -> store positions and ids; for list shards (cm_ivfpq_sharded_*) also the
// candidate's number in the reference's append loop over ALL probed lists, from the global list lengths
__global__ void adc_emit_kernel(const long long *__restrict__ probe_list, const long long *__restrict__ q_off,
                                const long long *__restrict__ list_off, const uint32_t *__restrict__ members, int nprobes,
                                const uint32_t *__restrict__ row_ids, long long out_stride, uint32_t *__restrict__ out_ids,
                                long long *__restrict__ out_pos, const long long *__restrict__ out_counts,
                                const long long *__restrict__ glob_len, uint32_t *__restrict__ out_gno) {
    const int q = blockIdx.x;
    const long long m = out_counts[q];
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        size_t o = (size_t)q * out_stride + i;
        long long c = out_ids[o];
        uint32_t pos;
        if (members) {
            const long long *qo = q_off + (size_t)q * (nprobes + 1);
            const long long *pl = probe_list + (size_t)q * nprobes;
            int lo = 0, hi = nprobes;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (qo[mid] <= c) lo = mid; else hi = mid;
            }
            pos = members[list_off[pl[lo]] + (c - qo[lo])];
            if (out_gno) {
                long long before = 0;
                for (int j = 0; j < lo; j++) before += glob_len[pl[j]];
                out_gno[o] = (uint32_t)(before + (c - qo[lo]));
            }
        } else {
            pos = (uint32_t)c;
        }
        out_ids[o] = row_ids[pos];
        if (out_pos) out_pos[o] = pos;
    }
}

// ivf.cu: per query prefix sums of the probed list lengths (the reference's candidate numbering)
int launch_ivf_offsets(const long long *probe_list, const long long *probe_cnt, const long long *list_off, int nprobes,
                       int64_t nq, long long *q_off, unsigned long long *total, cudaStream_t st);

static int prepare_queries(int metric, int dim, int ld, const float *q_dev, int64_t nq, bool check_zero, float **qp_out,
                           cudaStream_t st) {
    bool fma = rounding_mode() == CM_ROUND_FMA;
    int64_t nq_pad = (nq + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
    float *qp = nullptr;
    int *qflags = nullptr;
    CM_TRY(ws_alloc((void **)&qp, (size_t)nq_pad * ld * 4, st));
    CM_TRY(ws_alloc((void **)&qflags, (size_t)nq * sizeof(int), st));
    if (nq_pad > nq) CM_CUDA(cudaMemsetAsync(qp + (size_t)nq * ld, 0, (size_t)(nq_pad - nq) * ld * 4, st));
    CM_TRY(launch_preprocess_rows(metric, fma, q_dev, nq, dim, dim, qp, ld, qflags, st));
    if (check_zero && metric == CM_COSINE) {
        std::vector<int> hf((size_t)nq);
        CM_CUDA(cudaMemcpyAsync(hf.data(), qflags, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < nq; i++)
            if (hf[(size_t)i]) {
                ws_free(qp, st); ws_free(qflags, st);
                return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)i);
            }
    }
    ws_free(qflags, st);
    *qp_out = qp;
    return CM_OK;
}

static int build_skip(CodeStore &S, const cm_search_params *p, const uint8_t **skip, uint8_t **skip_buf, uint32_t **filt_dev,
                      cudaStream_t st) {
    *skip = nullptr; *skip_buf = nullptr; *filt_dev = nullptr;
    if (p->filter_ids && p->nfilter > 0) {
        std::vector<uint32_t> f(p->filter_ids, p->filter_ids + p->nfilter);
        std::sort(f.begin(), f.end());
        f.erase(std::unique(f.begin(), f.end()), f.end());
        CM_TRY(ws_alloc((void **)filt_dev, f.size() * 4, st));
        CM_TRY(ws_alloc((void **)skip_buf, (size_t)S.n, st));
        CM_CUDA(cudaMemcpyAsync(*filt_dev, f.data(), f.size() * 4, cudaMemcpyHostToDevice, st));
        CM_TRY(launch_build_skip(S.ids, S.deleted, S.n, *filt_dev, (int64_t)f.size(), *skip_buf, st));
        CM_CUDA(cudaStreamSynchronize(st));
        *skip = *skip_buf;
    } else if (S.n_deleted_rows > 0) {
        *skip = S.deleted;
    }
    return CM_OK;
}

// nq independent searchSingleQuery calls (PQ when ix.nlist == 0, else IVFPQ)
static int adc_search_device(PQCore &ix, const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                             uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, cudaStream_t st,
                             bool check_zero, const long long *glob_len = nullptr, uint32_t *out_gno = nullptr) {
    WsScope ws(st);
    if (nq <= 0) return CM_OK;
    const bool ivf = ix.nlist > 0;
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, ivf ? "index must be trained before searching" : "index not trained");
    CodeStore &S = ix.store;
    int nprobes = 1;
    if (ivf) {
        nprobes = p->nprobes;
        if (nprobes <= 0 || nprobes > ix.nlist) nprobes = ix.nlist;
    }
    CM_TRY(ix.sync_csr(st));
    CM_TRY(ix.sync_codebooks_t(st));
    int64_t bound_c = S.n, max_len = S.n;
    if (ivf) {
        bound_c = 0;
        for (int i = 0; i < nprobes && i < (int)ix.sizes_desc.size(); i++) bound_c += ix.sizes_desc[(size_t)i];
        max_len = ix.sizes_desc.empty() ? 0 : ix.sizes_desc[0];
    }
    int64_t k_eff = p->k;
    if (k_eff <= 0 || k_eff > bound_c) k_eff = bound_c;
    if (out_stride < k_eff)
        return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < effective k %lld", (long long)out_stride, (long long)k_eff);
    // k beyond what a CTA can select in shared memory (WithK(0) on a large index): every CTA gets a slice of at most
    // ADC_BIGK_PART codes and returns ALL of their keys; one radix sort per query orders them (flat_bigk.cu)
    const bool bigk = k_eff > 8192;
    float *qp = nullptr;
    if (!ivf && S.n == 0) {      // pq_index_search.go:232: an empty index answers before Preprocess
        CM_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)nq * sizeof(int64_t), st));
        return CM_OK;
    }
    CM_TRY(prepare_queries(ix.metric, ix.dim, ix.ld, q_dev, nq, check_zero, &qp, st));
    ws.adopt(qp);
    if (bound_c == 0 || k_eff == 0) {
        CM_TRY(launch_fill_counts(out_counts, nq, 0, st));       // a kernel: the counts may live on a peer device (list shards)
        return CM_OK;
    }
    bool fma = rounding_mode() == CM_ROUND_FMA;
    int64_t nq_pad = (nq + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;

    long long *probe_list = nullptr, *probe_cnt = nullptr, *q_off = nullptr;
    uint32_t *c_ids = nullptr;
    float *c_sc = nullptr;
    if (ivf) {
        CM_TRY(ws.get(&c_ids, (size_t)nq * nprobes * 4));
        CM_TRY(ws.get(&c_sc, (size_t)nq * nprobes * 4));
        CM_TRY(ws.get(&probe_list, (size_t)nq * nprobes * 8));
        CM_TRY(ws.get(&probe_cnt, (size_t)nq * 8));
        CM_TRY(ws.get(&q_off, (size_t)nq * (nprobes + 1) * 8));
        cm_flat_stats cst{};
        CM_TRY(ix.coarse.search_exact(qp, nq, nq_pad, nprobes, nullptr, 0.0f, nprobes, c_ids, c_sc, (int64_t *)probe_list,
                                      (int64_t *)probe_cnt, st, &cst));
        if (!ix.scanned_total) CM_CUDA(cudaMalloc(&ix.scanned_total, 8));
        CM_CUDA(cudaMemsetAsync(ix.scanned_total, 0, 8, st));
        CM_TRY(launch_ivf_offsets(probe_list, probe_cnt, ix.list_off, nprobes, nq, q_off, ix.scanned_total, st));
    }
    const uint8_t *skip = nullptr;
    uint8_t *skip_buf = nullptr;
    uint32_t *filt_dev = nullptr;
    const int rc_skip = build_skip(S, p, &skip, &skip_buf, &filt_dev, st);
    ws.adopt(skip_buf); ws.adopt(filt_dev);
    CM_TRY(rc_skip);

    const int K = bigk ? ADC_BIGK_PART : (int)k_eff;         // keys a CTA keeps
    // code bytes per row in 16-byte words -> kernel variant and rows per thread per round
    int MW = ((ix.M & 15) == 0 && (ix.M / 16 == 1 || ix.M / 16 == 2 || ix.M / 16 == 4 || ix.M / 16 == 6 || ix.M / 16 == 8)) ? ix.M / 16 : 0;
    if (const char *e = getenv("COMET_B200_ADC_GENERIC")) if (atoi(e)) MW = 0;
    const int R = MW == 0 ? 1 : (MW <= 4 ? 4 : (MW <= 6 ? 3 : 2));
    const bool ring = ix.ring_layout;                // codes_by_list is in the ring layout: adc_ring_kernel
    int C = next_pow2(K + (ring ? ADC_RING_RUNS : R) * ADC_THREADS);         // a round's appends always fit above a compacted buffer
    if (C < 1024) C = 1024;
    const int lut_n = ix.lut_entries();
    size_t smem = (size_t)ix.M * lut_n * 4 + (size_t)((ix.dim + 3) & ~3) * 4 + (size_t)C * 8;
    if (ring) smem = (size_t)ix.M * 256 * 4 + (size_t)(ADC_THREADS / 32) * ADC_RING_RUNS * 32 * 4 + (size_t)C * 8;
    if (smem > max_smem_optin())
        return fail(CM_ERR_UNSUPPORTED, "M=%d x %d table entries + k=%d do not fit shared memory", ix.M, lut_n, K);
    // slices per pair: enough CTAs for a few waves (two CTAs per SM), but never so many that a CTA
    // builds its table for less than one chunk of codes
    int64_t n_slices = (8ll * sm_count() + nq * nprobes - 1) / (nq * nprobes);
    n_slices = std::max<int64_t>(1, std::min<int64_t>(n_slices, (max_len + ADC_CHUNK - 1) / ADC_CHUNK));
    if (const char *e = getenv("COMET_B200_ADC_SLICES")) n_slices = std::max<int64_t>(1, atoll(e));
    if (bigk) n_slices = std::max<int64_t>(1, (max_len + ADC_BIGK_PART - 1) / ADC_BIGK_PART);
    const int64_t parts = (int64_t)nprobes * n_slices;              // per query
    int64_t qgroup = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 29) / (parts * K * 8)));
    if (qgroup * nprobes > 65535) qgroup = std::max<int64_t>(1, 65535 / nprobes);   // grid.y limit
    uint64_t *pk = nullptr;
    int *pc = nullptr;
    using AdcKernel = void (*)(const float *, int, int, int, int, int, int, const float *, const uint8_t *, long long,
                               const float *, int, const long long *, const long long *, const long long *, const uint32_t *,
                               int, const uint8_t *, float, int, int, int, uint64_t *, int *, const long long *, const float4 *);
    CM_TRY(ws.get(&pk, (size_t)qgroup * parts * K * 8));
    CM_TRY(ws.get(&pc, (size_t)qgroup * parts * 4));
    AdcKernel kern = nullptr;
#define CM_ADC_CASE(W) case W: kern = fma ? adc_scan_kernel<true, W> : adc_scan_kernel<false, W>; break;
    switch (MW) {
        CM_ADC_CASE(0) CM_ADC_CASE(1) CM_ADC_CASE(2) CM_ADC_CASE(4) CM_ADC_CASE(6) CM_ADC_CASE(8)
    }
#undef CM_ADC_CASE
    using RingKernel = void (*)(const float *, int, const float4 *, const uint8_t *, const float *, int, const long long *,
                                const long long *, const long long *, const long long *, const uint32_t *, int, const uint8_t *,
                                float, int, int, int, uint64_t *, int *, unsigned long long *, int);
    unsigned long long *q_tau = nullptr;      // per query: the smallest K-th-key bound any of its CTAs has found
    if (ring && !bigk) CM_TRY(ws.get(&q_tau, (size_t)std::min(qgroup, nq) * 8));
    RingKernel rkern = nullptr;
    if (ring) {
        const int per = ix.M / 32, d4 = ix.dsub / 4;
#define CM_RING_CASE(P, D) if (per == P && d4 == D) rkern = fma ? adc_ring_kernel<true, P, D> : adc_ring_kernel<false, P, D>;
        CM_RING_CASE(1, 2) CM_RING_CASE(1, 6) CM_RING_CASE(2, 1) CM_RING_CASE(2, 2) CM_RING_CASE(2, 3)
        CM_RING_CASE(3, 1) CM_RING_CASE(3, 2) CM_RING_CASE(4, 1) CM_RING_CASE(4, 2)
#undef CM_RING_CASE
        if (!rkern) return fail(CM_ERR_UNSUPPORTED, "ring layout without a kernel for M=%d dsub=%d", ix.M, ix.dsub);
    }
    CM_TRY(set_dyn_smem(ring ? (const void *)rkern : (const void *)kern, smem));      // only ever grows: concurrent searches ask for different sizes
    for (int64_t q0 = 0; q0 < nq; q0 += qgroup) {
        int64_t m = std::min(qgroup, nq - q0);
        CM_CUDA(cudaMemsetAsync(pc, 0, (size_t)m * parts * 4, st));
        dim3 grid((unsigned)n_slices, (unsigned)(m * nprobes));
        {
            ProfScope prof(CM_PROF_PQ_SCAN, st);
            if (ring && q_tau) CM_CUDA(cudaMemsetAsync(q_tau, 0xff, (size_t)m * 8, st));
            if (ring)
                rkern<<<grid, ADC_THREADS, smem, st>>>(qp + (size_t)q0 * ix.ld, ix.ld, reinterpret_cast<const float4 *>(ix.codebooks_r),
                                                       ix.codes_by_list, ivf ? ix.coarse.rows : nullptr, ix.coarse.ld,
                                                       ivf ? probe_list + (size_t)q0 * nprobes : nullptr,
                                                       ivf ? q_off + (size_t)q0 * (nprobes + 1) : nullptr, ix.list_off, ix.tile_off,
                                                       ivf ? ix.members : nullptr, nprobes, skip, p->threshold, K, C, (int)n_slices, pk, pc,
                                                       q_tau, (int)m);
            else
            kern<<<grid, ADC_THREADS, smem, st>>>(qp + (size_t)q0 * ix.ld, ix.ld, ix.dim, ix.M, ix.Ksub, ix.dsub, lut_n,
                                                  ix.codebooks, ivf ? ix.codes_by_list : S.codes, (long long)S.n, ivf ? ix.coarse.rows : nullptr,
                                                  ix.coarse.ld, ivf ? probe_list + (size_t)q0 * nprobes : nullptr,
                                                  ivf ? q_off + (size_t)q0 * (nprobes + 1) : nullptr, ix.list_off,
                                                  ivf ? ix.members : nullptr, nprobes, skip, p->threshold, K, C, (int)n_slices, pk, pc,
                                                  (ivf && ix.tiled()) ? ix.tile_off : nullptr,
                                                  ix.dsub == 8 ? reinterpret_cast<const float4 *>(ix.codebooks_t) : nullptr);
            count_launch();
            CM_CUDA(cudaGetLastError());
        }
        CM_TRY(launch_merge_topk(pk, pc, (int)m, (int)parts, K, (int)k_eff, nullptr, out_stride, out_ids + (size_t)q0 * out_stride,
                                 out_scores + (size_t)q0 * out_stride, nullptr, out_counts + q0, st));
        adc_emit_kernel<<<(unsigned)m, 128, 0, st>>>(ivf ? probe_list + (size_t)q0 * nprobes : nullptr,
                                                     ivf ? q_off + (size_t)q0 * (nprobes + 1) : nullptr, ix.list_off,
                                                     ivf ? ix.members : nullptr, nprobes, S.ids, (long long)out_stride,
                                                     out_ids + (size_t)q0 * out_stride,
                                                     out_pos ? (long long *)out_pos + (size_t)q0 * out_stride : nullptr,
                                                     (const long long *)(out_counts + q0), glob_len,
                                                     out_gno ? out_gno + (size_t)q0 * out_stride : nullptr);
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    return CM_OK;
}

// n successive Add calls: PreprocessInPlace, [nearest centroid, residual], encode, append
static int adc_add(PQCore &ix, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists) {
    const bool ivf = ix.nlist > 0;
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, ivf ? "index must be trained before adding" : "index must be trained before adding vectors");
    if (n <= 0) return CM_OK;
    bool fma = rounding_mode() == CM_ROUND_FMA;
    cudaStream_t st;
    CM_TRY(acquire_stream(&st));
    const int64_t slab = std::max<int64_t>(8, (int64_t)(64u << 20) / ((int64_t)ix.ld * 4));
    int64_t sl = std::min(slab, n);
    int64_t sl_pad = (sl + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
    float *raw = nullptr, *pre = nullptr, *a_sc = nullptr;
    int *flags = nullptr;
    uint32_t *a_ids = nullptr;
    long long *a_pos = nullptr, *a_cnt = nullptr;
    int rc = ws_alloc((void **)&raw, (size_t)sl * ix.dim * 4, st);
    if (rc == CM_OK) rc = ws_alloc((void **)&pre, (size_t)sl_pad * ix.ld * 4, st);
    if (rc == CM_OK) rc = ws_alloc((void **)&flags, (size_t)sl * 4, st);
    if (rc == CM_OK && ivf) {
        rc = ws_alloc((void **)&a_ids, (size_t)sl * 4, st);
        if (rc == CM_OK) rc = ws_alloc((void **)&a_sc, (size_t)sl * 4, st);
        if (rc == CM_OK) rc = ws_alloc((void **)&a_pos, (size_t)sl * 8, st);
        if (rc == CM_OK) rc = ws_alloc((void **)&a_cnt, (size_t)sl * 8, st);
    }
    std::vector<int> hflags((size_t)sl);
    std::vector<long long> hpos((size_t)sl);
    int rc_zero = CM_OK;
    for (int64_t i0 = 0; rc == CM_OK && rc_zero == CM_OK && i0 < n; i0 += slab) {
        int64_t m = std::min(slab, n - i0);
        cudaMemcpyAsync(raw, rows + (size_t)i0 * ix.dim, (size_t)m * ix.dim * 4, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(pre, 0, (size_t)sl_pad * ix.ld * 4, st);
        rc = launch_preprocess_rows(ix.metric, fma, raw, m, ix.dim, ix.dim, pre, ix.ld, flags, st);
        if (rc != CM_OK) break;
        int64_t good = m;
        if (ix.metric == CM_COSINE) {
            cudaMemcpyAsync(hflags.data(), flags, (size_t)m * 4, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            for (int64_t i = 0; i < m; i++)
                if (hflags[(size_t)i]) { good = i; break; }
            if (writeback && good > 0)
                cudaMemcpy2DAsync(rows + (size_t)i0 * ix.dim, (size_t)ix.dim * 4, pre, (size_t)ix.ld * 4, (size_t)ix.dim * 4,
                                  (size_t)good, cudaMemcpyDeviceToHost, st);
        }
        if (good < m) rc_zero = fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (row %lld of this Add batch)", (long long)(i0 + good));
        if (good == 0) break;
        if (ivf) {
            int64_t gpad = (good + SCAN_MAX_QB - 1) / SCAN_MAX_QB * SCAN_MAX_QB;
            cm_flat_stats cst{};
            rc = ix.coarse.search_exact(pre, good, gpad, 1, nullptr, 0.0f, 1, a_ids, a_sc, (int64_t *)a_pos, (int64_t *)a_cnt, st, &cst);
            if (rc != CM_OK) break;
        }
        rc = ix.store.reserve(ix.store.n + good);
        if (rc != CM_OK) break;
        long long threads = good * ix.M;
        unsigned blocks = (unsigned)((threads + 127) / 128);
        if (fma)
            pq_encode_kernel<true><<<blocks, 128, 0, st>>>(pre, ix.ld, good, ix.M, ix.Ksub, ix.dsub, ix.codebooks,
                                                          ivf ? ix.coarse.rows : nullptr, ix.coarse.ld, a_pos,
                                                          ix.store.codes + (size_t)ix.store.n * ix.M);
        else
            pq_encode_kernel<false><<<blocks, 128, 0, st>>>(pre, ix.ld, good, ix.M, ix.Ksub, ix.dsub, ix.codebooks,
                                                           ivf ? ix.coarse.rows : nullptr, ix.coarse.ld, a_pos,
                                                           ix.store.codes + (size_t)ix.store.n * ix.M);
        count_launch();
        if (ivf) cudaMemcpyAsync(hpos.data(), a_pos, (size_t)good * 8, cudaMemcpyDeviceToHost, st);
        int64_t n_before = ix.store.n;
        rc = ix.store.commit(ids + i0, good, st);      // synchronises the stream
        ix.csr_dirty = true;
        if (rc != CM_OK) break;
        if (ivf) {
            for (int64_t i = 0; i < good; i++) {
                int32_t l = (int32_t)hpos[(size_t)i];
                ix.lists[(size_t)l].push_back((uint32_t)(n_before + i));
                ix.list_of.push_back(l);
                if (out_lists) out_lists[i0 + i] = l;
            }
            ix.csr_dirty = true;
        }
    }
    ws_free(raw, st); ws_free(pre, st); ws_free(flags, st); ws_free(a_ids, st); ws_free(a_sc, st); ws_free(a_pos, st); ws_free(a_cnt, st);
    cudaError_t e = cudaStreamSynchronize(st);
    release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return fail(CM_ERR_CUDA, "add: %s", cudaGetErrorString(e));
    return rc != CM_OK ? rc : rc_zero;
}

// PQIndex.ReadFrom (pq_index.go:652-846) / IVFPQIndex.ReadFrom: restore codes (and list membership) as stored
static int adc_load_codes(PQCore &ix, const uint32_t *ids, const uint8_t *codes, const int32_t *list_of, int64_t n) {
    const bool ivf = ix.nlist > 0;
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, "index must be trained (codebooks loaded) before codes are restored");
    if (n <= 0) return CM_OK;
    if (ivf) {
        if (!list_of) return fail(CM_ERR_INVALID_ARG, "list_of is required for an IVFPQ index");
        for (int64_t i = 0; i < n; i++)
            if (list_of[i] < 0 || list_of[i] >= ix.nlist) return fail(CM_ERR_INVALID_ARG, "row %lld: list %d out of range", (long long)i, list_of[i]);
    }
    cudaStream_t st;
    CM_TRY(acquire_stream(&st));
    int rc = ix.store.reserve(ix.store.n + n);
    int64_t n_before = ix.store.n;
    if (rc == CM_OK) {
        cudaError_t e = cudaMemcpyAsync(ix.store.codes + (size_t)n_before * ix.M, codes, (size_t)n * ix.M, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) rc = fail(CM_ERR_CUDA, "load_codes: %s", cudaGetErrorString(e));
    }
    if (rc == CM_OK) rc = ix.store.commit(ids, n, st);
    ix.csr_dirty = true;
    if (rc == CM_OK && ivf) {
        for (int64_t i = 0; i < n; i++) {
            ix.lists[(size_t)list_of[i]].push_back((uint32_t)(n_before + i));
            ix.list_of.push_back(list_of[i]);
        }
        ix.csr_dirty = true;
    }
    cudaStreamSynchronize(st);
    release_stream(st);
    return rc;
}

static int adc_flush(PQCore &ix) {
    if (ix.store.deleted_ids.empty()) return CM_OK;
    std::vector<int64_t> new_pos;
    int64_t n_old = ix.store.n;
    CM_TRY(ix.store.flush(&new_pos));
    ix.csr_dirty = true;
    if (ix.nlist > 0) {
        std::vector<int32_t> nlo((size_t)ix.store.n);
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
    }
    return CM_OK;
}

static int core_create(int dim, int metric, int nlist, int M, int nbits, bool ivf, PQCore *ix) {
    // pq_index.go:135-170 / ivfpq_index.go:114-160 constructor checks
    if (dim <= 0) return fail(CM_ERR_INVALID_ARG, "dimension must be positive");
    if (ivf && nlist <= 0) return fail(CM_ERR_INVALID_ARG, "nlist must be positive");
    if (M <= 0) return fail(CM_ERR_INVALID_ARG, "M must be positive");
    if (dim % M != 0) return fail(CM_ERR_INVALID_ARG, "dimension %d must be divisible by M %d", dim, M);
    if (nbits <= 0 || nbits > 16) return fail(CM_ERR_INVALID_ARG, "Nbits must be between 1 and 16");
    if (metric < 0 || metric > 2) return fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    CM_TRY(ensure_device());
    ix->dim = dim; ix->metric = metric; ix->M = M; ix->nbits = nbits; ix->Ksub = 1 << nbits; ix->dsub = dim / M;
    ix->ld = (dim + SCAN_CHUNK - 1) / SCAN_CHUNK * SCAN_CHUNK;
    ix->nlist = ivf ? nlist : 0;
    ix->store.M = M;
    cudaGetDevice(&ix->device);
    if (ivf) {
        ix->coarse.dim = dim; ix->coarse.ld = ix->ld; ix->coarse.metric = metric; ix->coarse.device = ix->device;
        ix->coarse.raw_rows = true;
        ix->lists.resize((size_t)nlist);
    }
    return CM_OK;
}

static int core_set_trained(PQCore &ix, const float *centroids, const float *codebooks) {
    if (ix.store.n > 0) return fail(CM_ERR_INVALID_ARG, "cannot replace the trained state of a non-empty index");
    size_t cb_bytes = (size_t)ix.M * ix.Ksub * ix.dsub * 4;
    if (!ix.codebooks) CM_CUDA(cudaMalloc(&ix.codebooks, cb_bytes));
    CM_CUDA(cudaMemcpy(ix.codebooks, codebooks, cb_bytes, cudaMemcpyHostToDevice));
    if (ix.nlist > 0) {
        cudaStream_t st;
        CM_TRY(acquire_stream(&st));
        ix.coarse.n = 0; ix.coarse.ids_host_mirror.clear();
        float *stage = nullptr;
        size_t bytes = (size_t)ix.nlist * ix.dim * 4;
        int rc = ws_alloc((void **)&stage, bytes, st);
        if (rc == CM_OK) {
            cudaMemcpyAsync(stage, centroids, bytes, cudaMemcpyHostToDevice, st);
            std::vector<uint32_t> ids((size_t)ix.nlist);
            std::iota(ids.begin(), ids.end(), 0u);
            rc = ix.coarse.add_from_device(ids.data(), stage, ix.nlist, nullptr, st);
        }
        ws_free(stage, st);
        cudaStreamSynchronize(st);
        release_stream(st);
        CM_TRY(rc);
    }
    ix.trained = true;
    ix.cbt_dirty = true;
    return CM_OK;
}

// PQIndex.Train (pq_index.go:193-247) / IVFPQIndex.Train (ivfpq_index.go:180-259) on the device
static int core_train(PQCore &ix, const float *rows, int64_t n) {
    const bool ivf = ix.nlist > 0;
    if (ivf && n < (int64_t)ix.nlist * 10) return fail(CM_ERR_TOO_FEW, "need at least %d vectors for training", ix.nlist * 10);
    if (!ivf && n < ix.Ksub) return fail(CM_ERR_TOO_FEW, "need at least %d vectors for training", ix.Ksub);
    if (ix.store.n > 0) return fail(CM_ERR_UNSUPPORTED, "retraining a non-empty index is not supported");
    cudaStream_t st;
    CM_TRY(acquire_stream(&st));
    float *x = nullptr, *res = nullptr;
    long long *assign = nullptr;
    size_t cb_bytes = (size_t)ix.M * ix.Ksub * ix.dsub * 4;
    int rc = upload_training_rows(rows, n, ix.dim, ix.ld, &x, st);
    if (rc == CM_OK && !ix.codebooks) {
        cudaError_t e = cudaMalloc(&ix.codebooks, cb_bytes);
        if (e != cudaSuccess) rc = fail(CM_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    const float *sub_src = x;
    if (rc == CM_OK && ivf) {
        rc = ws_alloc((void **)&assign, (size_t)n * 8, st);
        if (rc == CM_OK) rc = ws_alloc((void **)&res, (size_t)n * ix.ld * 4, st);
        if (rc == CM_OK) rc = kmeans_full(ix.coarse, x, n, ix.nlist, 20, assign, st);
        if (rc == CM_OK) rc = launch_residuals(x, n, ix.dim, ix.ld, ix.coarse.rows, assign, res, st);
        sub_src = res;
    }
    for (int m = 0; rc == CM_OK && m < ix.M; m++)
        rc = kmeans_subspace(sub_src, n, ix.ld, m * ix.dsub, ix.dsub, ix.Ksub, 20, ix.codebooks + (size_t)m * ix.Ksub * ix.dsub, st);
    ws_free(x, st); ws_free(res, st); ws_free(assign, st);
    cudaError_t e = cudaStreamSynchronize(st);
    release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) rc = fail(CM_ERR_CUDA, "train: %s", cudaGetErrorString(e));
    if (rc == CM_OK) { ix.trained = true; ix.cbt_dirty = true; }
    return rc;
}

static int core_get_trained(const PQCore &ix, float *centroids, float *codebooks) {
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, "index must be trained");
    if (codebooks) CM_CUDA(cudaMemcpy(codebooks, ix.codebooks, (size_t)ix.M * ix.Ksub * ix.dsub * 4, cudaMemcpyDeviceToHost));
    if (centroids && ix.nlist > 0)
        CM_CUDA(cudaMemcpy2D(centroids, (size_t)ix.dim * 4, ix.coarse.rows, (size_t)ix.coarse.ld * 4, (size_t)ix.dim * 4, (size_t)ix.nlist,
                             cudaMemcpyDeviceToHost));
    return CM_OK;
}

// ---- wire formats "PQIX" (pq_index.go:509-846) and "IVPQ" (ivfpq_index.go:544-960) -----------------------------
// back to the state NewPQIndex / NewIVFPQIndex leaves: untrained, no codes
static int core_reset(PQCore &ix) {
    if (ix.store.deleted && ix.store.cap > 0) CM_CUDA(cudaMemset(ix.store.deleted, 0, (size_t)ix.store.cap));
    ix.store.n = 0;
    ix.csr_dirty = true;
    ix.store.n_deleted_rows = 0;
    ix.store.ids_host.clear();
    ix.store.deleted_ids.clear();
    if (ix.nlist > 0) {
        CM_TRY(ix.coarse.reset());
        for (auto &l : ix.lists) l.clear();
        ix.list_of.clear();
        ix.csr_dirty = true;
    }
    ix.trained = false;
    return CM_OK;
}

// WriteTo of both index types: Flush; header; [nlist]; M, Nbits, Ksub, dsub; trained flag; [centroids]; codebooks
// (size + data per sub-quantiser); then PQ: vector count, per vector (ID, code) in arrival order -- IVFPQ: list count,
// per list (size, per vector ID + code); roaring blob.
static int core_save(PQCore &ix, wire::Sink &s) {
    const bool ivf = ix.nlist > 0;
    CM_TRY(adc_flush(ix));
    CM_TRY(wire::write_header(s, ivf ? "IVPQ" : "PQIX", ix.dim, ix.metric));
    if (ivf) CM_WIRE_PUT(s.u32((uint32_t)ix.nlist), "nlist");
    CM_WIRE_PUT(s.u32((uint32_t)ix.M), "M");
    CM_WIRE_PUT(s.u32((uint32_t)ix.nbits), "Nbits");
    CM_WIRE_PUT(s.u32((uint32_t)ix.Ksub), "Ksub");
    CM_WIRE_PUT(s.u32((uint32_t)ix.dsub), "dsub");
    CM_WIRE_PUT(s.u8(ix.trained ? 1 : 0), "trained flag");
    if (ix.trained) {
        const size_t cb = (size_t)ix.Ksub * ix.dsub;
        std::vector<float> cent(ivf ? (size_t)ix.nlist * ix.dim : 0), books((size_t)ix.M * cb);
        CM_TRY(core_get_trained(ix, ivf ? cent.data() : nullptr, books.data()));
        for (int l = 0; ivf && l < ix.nlist; l++) {
            CM_WIRE_PUT(s.u32((uint32_t)ix.dim), "centroid size");
            CM_WIRE_PUT(s.put(&cent[(size_t)l * ix.dim], (size_t)ix.dim * 4), "centroid data");
        }
        for (int m = 0; m < ix.M; m++) {
            CM_WIRE_PUT(s.u32((uint32_t)cb), "codebook size");
            CM_WIRE_PUT(s.put(&books[(size_t)m * cb], cb * 4), "codebook data");
        }
    }
    const int64_t n = ix.store.n;
    std::vector<uint8_t> codes((size_t)n * ix.M);
    if (n > 0) CM_CUDA(cudaMemcpy(codes.data(), ix.store.codes, codes.size(), cudaMemcpyDeviceToHost));
    const size_t per = 4 + (size_t)ix.M;
    std::vector<uint8_t> rec;
    auto put_rows = [&](const uint32_t *pos, int64_t first, int64_t m) {      // pos == NULL: positions first, first + 1, ...
        rec.resize((size_t)m * per);
        for (int64_t i = 0; i < m; i++) {
            const int64_t at = pos ? (int64_t)pos[i] : first + i;
            memcpy(&rec[(size_t)i * per], &ix.store.ids_host[(size_t)at], 4);
            memcpy(&rec[(size_t)i * per + 4], &codes[(size_t)at * ix.M], (size_t)ix.M);
        }
        return s.put(rec.data(), rec.size());
    };
    if (!ivf) {
        CM_WIRE_PUT(s.u32((uint32_t)n), "vector count");
        CM_WIRE_PUT(put_rows(nullptr, 0, n), "vector codes");
    } else {
        CM_WIRE_PUT(s.u32((uint32_t)ix.lists.size()), "list count");
        for (const std::vector<uint32_t> &L : ix.lists) {
            CM_WIRE_PUT(s.u32((uint32_t)L.size()), "list size");
            CM_WIRE_PUT(put_rows(L.data(), 0, (int64_t)L.size()), "list codes");
        }
    }
    CM_WIRE_PUT(wire::write_empty_bitmap(s), "bitmap");
    return CM_OK;
}

// ReadFrom of both index types: decode and validate everything, then replace the index state.
static int core_load(PQCore &ix, wire::Source &s) {
    const bool ivf = ix.nlist > 0;
    CM_TRY(wire::read_header(s, ivf ? "IVPQ" : "PQIX", ix.dim, ix.metric));
    uint32_t nlist = 0, M = 0, nbits = 0, Ksub = 0, dsub = 0;
    uint8_t trained = 0;
    if (ivf) CM_WIRE_GET(s.u32(&nlist), "nlist");
    CM_WIRE_GET(s.u32(&M), "M");
    CM_WIRE_GET(s.u32(&nbits), "Nbits");
    CM_WIRE_GET(s.u32(&Ksub), "Ksub");
    CM_WIRE_GET(s.u32(&dsub), "dsub");
    if (ivf && (int64_t)nlist != ix.nlist) return fail(CM_ERR_INVALID_ARG, "parameter nlist mismatch: index has nlist=%d, serialized data has nlist=%u", ix.nlist, nlist);
    if ((int64_t)M != ix.M) return fail(CM_ERR_INVALID_ARG, "parameter M mismatch: index has M=%d, serialized data has M=%u", ix.M, M);
    if ((int64_t)nbits != ix.nbits) return fail(CM_ERR_INVALID_ARG, "parameter Nbits mismatch: index has Nbits=%d, serialized data has Nbits=%u", ix.nbits, nbits);
    if ((int64_t)Ksub != ix.Ksub) return fail(CM_ERR_INVALID_ARG, "parameter Ksub mismatch: index has Ksub=%d, serialized data has Ksub=%u", ix.Ksub, Ksub);
    if ((int64_t)dsub != ix.dsub) return fail(CM_ERR_INVALID_ARG, "parameter dsub mismatch: index has dsub=%d, serialized data has dsub=%u", ix.dsub, dsub);
    CM_WIRE_GET(s.u8(&trained), "trained flag");
    const size_t cb = (size_t)ix.Ksub * ix.dsub;
    std::vector<float> cent, books;
    if (trained == 1) {
        if (ivf) {
            cent.resize((size_t)ix.nlist * ix.dim);
            for (int l = 0; l < ix.nlist; l++) {
                uint32_t sz = 0;
                CM_WIRE_GET(s.u32(&sz), "centroid size");
                if ((int64_t)sz != ix.dim) return fail(CM_ERR_DIM_MISMATCH, "centroid %d has dimension %u, expected %d", l, sz, ix.dim);
                CM_WIRE_GET(s.get(&cent[(size_t)l * ix.dim], (size_t)ix.dim * 4), "centroid data");
            }
        }
        books.resize((size_t)ix.M * cb);
        for (int m = 0; m < ix.M; m++) {
            uint32_t sz = 0;
            CM_WIRE_GET(s.u32(&sz), "codebook size");
            if ((size_t)sz != cb) return fail(CM_ERR_INVALID_ARG, "codebook %d has %u values, expected %zu", m, sz, cb);
            CM_WIRE_GET(s.get(&books[(size_t)m * cb], cb * 4), "codebook data");
        }
    }
    std::vector<uint32_t> ids;
    std::vector<int32_t> list_of;
    std::vector<uint8_t> codes;
    auto get_rows = [&](uint32_t count, int32_t list) {
        const size_t at = ids.size();
        ids.resize(at + count);
        codes.resize((at + count) * (size_t)ix.M);
        if (ivf) list_of.resize(at + count, list);
        for (uint32_t i = 0; i < count; i++)
            if (!s.u32(&ids[at + i]) || !s.get(&codes[(at + i) * (size_t)ix.M], (size_t)ix.M)) return false;
        return true;
    };
    if (!ivf) {
        uint32_t count = 0;
        CM_WIRE_GET(s.u32(&count), "vector count");
        CM_WIRE_GET(get_rows(count, 0), "vector codes");
    } else {
        uint32_t list_count = 0;
        CM_WIRE_GET(s.u32(&list_count), "list count");
        if (list_count > (uint32_t)ix.nlist) return fail(CM_ERR_INVALID_ARG, "serialized data has %u lists, index has nlist=%d", list_count, ix.nlist);
        for (uint32_t l = 0; l < list_count; l++) {
            uint32_t sz = 0;
            CM_WIRE_GET(s.u32(&sz), "list size");
            CM_WIRE_GET(get_rows(sz, (int32_t)l), "list codes");
        }
    }
    std::vector<uint32_t> dead;
    CM_TRY(wire::read_bitmap(s, &dead));
    if (!ids.empty() && trained != 1) return fail(CM_ERR_NOT_TRAINED, "serialized data holds codes but no codebooks");
    CM_TRY(core_reset(ix));
    if (trained == 1) CM_TRY(core_set_trained(ix, ivf ? cent.data() : nullptr, books.data()));
    if (!ids.empty()) CM_TRY(adc_load_codes(ix, ids.data(), codes.data(), ivf ? list_of.data() : nullptr, (int64_t)ids.size()));
    for (uint32_t id : dead)
        if (ix.store.remove(id) != CM_OK) ix.store.deleted_ids.insert(id);
    return CM_OK;
}

static int core_search_host(PQCore &ix, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                            uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts) {
    if (!ix.trained) return fail(CM_ERR_NOT_TRAINED, ix.nlist > 0 ? "index must be trained before searching" : "index not trained");
    if (dim != ix.dim) return fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", ix.dim, dim);
    if (nq <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(ix.device));
    cudaStream_t st;
    CM_TRY(acquire_stream(&st));
    float *dq = nullptr, *dsc = nullptr;
    uint32_t *dids = nullptr;
    int64_t *dpos = nullptr, *dcnt = nullptr;
    size_t no = (size_t)nq * (size_t)(out_stride > 0 ? out_stride : 1);
    int rc = ws_alloc((void **)&dq, (size_t)nq * dim * 4, st);
    if (rc == CM_OK) rc = ws_alloc((void **)&dids, no * 4, st);
    if (rc == CM_OK) rc = ws_alloc((void **)&dsc, no * 4, st);
    if (rc == CM_OK && out_pos) rc = ws_alloc((void **)&dpos, no * 8, st);
    if (rc == CM_OK) rc = ws_alloc((void **)&dcnt, (size_t)nq * 8, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dq, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st);
        rc = adc_search_device(ix, dq, nq, p, out_stride, dids, dsc, dpos, dcnt, st, true);
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(out_ids, dids, no * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_scores, dsc, no * 4, cudaMemcpyDeviceToHost, st);
        if (out_pos) cudaMemcpyAsync(out_pos, dpos, no * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, dcnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, st);
    }
    ws_free(dq, st); ws_free(dids, st); ws_free(dsc, st); ws_free(dpos, st); ws_free(dcnt, st);
    cudaError_t e = cudaStreamSynchronize(st);
    release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return fail(CM_ERR_CUDA, "search: %s", cudaGetErrorString(e));
    return rc;
}

}  // namespace cm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct cm_pq { cm::PQCore ix; };
struct cm_ivfpq { cm::PQCore ix; };

extern "C" {

int cm_pq_create(int dim, int metric, int M, int nbits, cm_pq **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    cm_pq *h = new cm_pq();
    int rc = cm::core_create(dim, metric, 0, M, nbits, false, &h->ix);
    if (rc != CM_OK) { delete h; return rc; }
    *out = h;
    return CM_OK;
}
int cm_pq_destroy(cm_pq *h) { delete h; return CM_OK; }
int cm_pq_set_codebooks(cm_pq *h, const float *codebooks) {
    if (!h || !codebooks) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_set_trained(h->ix, nullptr, codebooks);
}
int cm_pq_train(cm_pq *h, const float *rows, int64_t n) {
    if (!h || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_train(h->ix, rows, n);
}
int cm_pq_get_codebooks(const cm_pq *h, float *out) {
    if (!h || !out) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_get_trained(h->ix, nullptr, out);
}
int cm_pq_trained(const cm_pq *h) { return h && h->ix.trained ? 1 : 0; }
int64_t cm_pq_size(const cm_pq *h) { return h ? h->ix.store.n : 0; }
int cm_pq_add(cm_pq *h, const uint32_t *ids, float *rows, int64_t n, int writeback) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_add(h->ix, ids, rows, n, writeback, nullptr);
}
int cm_pq_load_codes(cm_pq *h, const uint32_t *ids, const uint8_t *codes, int64_t n) {
    if (!h || (n > 0 && (!ids || !codes))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_load_codes(h->ix, ids, codes, nullptr, n);
}
int cm_pq_get_codes(const cm_pq *h, int64_t first, int64_t n, uint8_t *out) {
    if (!h || !out || first < 0 || first + n > h->ix.store.n) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    CM_CUDA(cudaMemcpy(out, h->ix.store.codes + (size_t)first * h->ix.M, (size_t)n * h->ix.M, cudaMemcpyDeviceToHost));
    return CM_OK;
}
int cm_pq_get_ids(const cm_pq *h, int64_t first, int64_t n, uint32_t *out) {       // node IDs by store position
    if (!h || first < 0 || n < 0 || first + n > h->ix.store.n || (n > 0 && !out)) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n > 0) memcpy(out, h->ix.store.ids_host.data() + first, (size_t)n * 4);
    return CM_OK;
}
int cm_pq_remove(cm_pq *h, uint32_t id) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.store.remove(id);
}
int cm_pq_flush(cm_pq *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_flush(h->ix);
}
int cm_pq_search(cm_pq *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                 uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    return cm::core_search_host(h->ix, queries, nq, dim, p, out_stride, out_ids, out_scores, out_pos, out_counts);
}
int cm_pq_search_device(cm_pq *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                        int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                        int64_t *out_counts_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_search_device(h->ix, queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                                 out_counts_dev, (cudaStream_t)stream, false);
}

int cm_ivfpq_create(int dim, int metric, int nlist, int M, int nbits, cm_ivfpq **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    cm_ivfpq *h = new cm_ivfpq();
    int rc = cm::core_create(dim, metric, nlist, M, nbits, true, &h->ix);
    if (rc != CM_OK) { delete h; return rc; }
    *out = h;
    return CM_OK;
}
int cm_ivfpq_destroy(cm_ivfpq *h) { delete h; return CM_OK; }
int cm_ivfpq_set_trained(cm_ivfpq *h, const float *centroids, const float *codebooks) {
    if (!h || !centroids || !codebooks) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_set_trained(h->ix, centroids, codebooks);
}
int cm_ivfpq_train(cm_ivfpq *h, const float *rows, int64_t n) {
    if (!h || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_train(h->ix, rows, n);
}
int cm_ivfpq_get_trained(const cm_ivfpq *h, float *centroids, float *codebooks) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::core_get_trained(h->ix, centroids, codebooks);
}
int cm_ivfpq_trained(const cm_ivfpq *h) { return h && h->ix.trained ? 1 : 0; }
int64_t cm_ivfpq_size(const cm_ivfpq *h) { return h ? h->ix.store.n : 0; }
int cm_ivfpq_default_nprobes(const cm_ivfpq *h) {   // ivfpq_index.go:446 int(sqrt(nlist))
    if (!h) return 0;
    int r = (int)std::sqrt((double)h->ix.nlist);
    while ((long long)(r + 1) * (r + 1) <= h->ix.nlist) r++;
    while ((long long)r * r > h->ix.nlist) r--;
    return r;
}
int cm_ivfpq_add(cm_ivfpq *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_add(h->ix, ids, rows, n, writeback, out_lists);
}
int cm_ivfpq_load_codes(cm_ivfpq *h, const uint32_t *ids, const uint8_t *codes, const int32_t *list_of, int64_t n) {
    if (!h || (n > 0 && (!ids || !codes || !list_of))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_load_codes(h->ix, ids, codes, list_of, n);
}
int cm_ivfpq_get_codes(const cm_ivfpq *h, int64_t first, int64_t n, uint8_t *out) {
    if (!h || !out || first < 0 || first + n > h->ix.store.n) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    CM_CUDA(cudaMemcpy(out, h->ix.store.codes + (size_t)first * h->ix.M, (size_t)n * h->ix.M, cudaMemcpyDeviceToHost));
    return CM_OK;
}
int cm_ivfpq_get_ids(const cm_ivfpq *h, int64_t first, int64_t n, uint32_t *out) {   // node IDs by store position
    if (!h || first < 0 || n < 0 || first + n > h->ix.store.n || (n > 0 && !out)) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n > 0) memcpy(out, h->ix.store.ids_host.data() + first, (size_t)n * 4);
    return CM_OK;
}
int64_t cm_ivfpq_last_scanned(const cm_ivfpq *h) {
    if (!h || !h->ix.scanned_total) return 0;
    unsigned long long v = 0;
    cudaSetDevice(h->ix.device);
    if (cudaMemcpy(&v, h->ix.scanned_total, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}
int cm_ivfpq_remove(cm_ivfpq *h, uint32_t id) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return h->ix.store.remove(id);
}
int cm_ivfpq_flush(cm_ivfpq *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_flush(h->ix);
}
int cm_ivfpq_search(cm_ivfpq *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                    uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    return cm::core_search_host(h->ix, queries, nq, dim, p, out_stride, out_ids, out_scores, out_pos, out_counts);
}
int cm_ivfpq_search_device(cm_ivfpq *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                           int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                           int64_t *out_counts_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!h->ix.trained) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::adc_search_device(h->ix, queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                                 out_counts_dev, (cudaStream_t)stream, false);
}

#define CM_WIRE_ABI(NAME, TYPE)                                                                                      \
    int NAME##_save(TYPE *h, uint8_t *buf, int64_t cap, int64_t *bytes) {                                           \
        if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");                                                 \
        CM_CUDA(cudaSetDevice(h->ix.device));                                                                       \
        return cm::wire::save_to_buffer([&](cm::wire::Sink &s) { return cm::core_save(h->ix, s); }, buf, cap, bytes); \
    }                                                                                                               \
    int NAME##_load(TYPE *h, const uint8_t *buf, int64_t len, int64_t *consumed) {                                  \
        if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");                                                 \
        CM_CUDA(cudaSetDevice(h->ix.device));                                                                       \
        return cm::wire::load_from_buffer([&](cm::wire::Source &s) { return cm::core_load(h->ix, s); }, buf, len, consumed); \
    }                                                                                                               \
    int NAME##_save_file(TYPE *h, const char *path) {                                                               \
        if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");                                                 \
        CM_CUDA(cudaSetDevice(h->ix.device));                                                                       \
        return cm::wire::save_to_file([&](cm::wire::Sink &s) { return cm::core_save(h->ix, s); }, path);            \
    }                                                                                                               \
    int NAME##_load_file(TYPE *h, const char *path) {                                                               \
        if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");                                                 \
        CM_CUDA(cudaSetDevice(h->ix.device));                                                                       \
        return cm::wire::load_from_file([&](cm::wire::Source &s) { return cm::core_load(h->ix, s); }, path);        \
    }
CM_WIRE_ABI(cm_pq, cm_pq)
CM_WIRE_ABI(cm_ivfpq, cm_ivfpq)
#undef CM_WIRE_ABI

// ================================================================================================
// IVFPQ list shards over the GPUs of one box, ONE host process: cm_ivfpq_sharded_*.  Every shard holds the
// centroids and the residual codebooks (replicated: they are small) and the code lists it owns; Add encodes once on
// devices[0] and routes the codes.  Search driver, global candidate numbers and merge: list_shards.cuh.
// ================================================================================================
struct cm_ivfpq_sharded {
    cm::ListShards ls;
    int M = 0, nbits = 0;
    std::vector<cm_ivfpq *> shard;
    cm_ivfpq *assigner = nullptr;          // devices[0]: trained state only; Add preprocesses, assigns and encodes here
};

static int ivfpqs_clear_vectors(cm::PQCore &ix) {          // drop the codes, keep the trained state
    if (ix.store.deleted && ix.store.cap > 0) CM_CUDA(cudaMemset(ix.store.deleted, 0, (size_t)ix.store.cap));
    ix.store.n = 0;
    ix.csr_dirty = true;
    ix.store.n_deleted_rows = 0;
    ix.store.ids_host.clear();
    ix.store.deleted_ids.clear();
    for (auto &l : ix.lists) l.clear();
    ix.list_of.clear();
    ix.csr_dirty = true;
    return CM_OK;
}

int cm_ivfpq_sharded_destroy(cm_ivfpq_sharded *h) {
    if (!h) return CM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    h->ls.workers.clear();
    for (size_t r = 0; r < h->shard.size(); r++)
        if (h->shard[r]) { cudaSetDevice(h->ls.dev[r]); cm_ivfpq_destroy(h->shard[r]); }
    if (h->assigner) { cudaSetDevice(h->ls.dev[0]); cm_ivfpq_destroy(h->assigner); }
    h->ls.destroy();
    cudaSetDevice(prev);
    delete h;
    return CM_OK;
}

int cm_ivfpq_sharded_create(int dim, int metric, int nlist, int M, int nbits, const int *devices, int n_devices,
                            cm_ivfpq_sharded **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices <= 0) return cm::fail(CM_ERR_INVALID_ARG, "need 1..64 devices");
    int prev = 0;
    cudaGetDevice(&prev);
    cm_ivfpq_sharded *h = new cm_ivfpq_sharded();
    h->M = M; h->nbits = nbits;
    // the first shard validates the parameters (NewIVFPQIndex, ivfpq_index.go:114-160) before anything else is set up
    h->shard.assign((size_t)n_devices, nullptr);
    int rc = CM_OK;
    for (int r = 0; r < n_devices && rc == CM_OK; r++) {
        if (cudaSetDevice(devices[r]) != cudaSuccess) { rc = cm::fail(CM_ERR_INVALID_ARG, "no CUDA device %d", devices[r]); break; }
        rc = cm_ivfpq_create(dim, metric, nlist, M, nbits, &h->shard[(size_t)r]);
    }
    if (rc == CM_OK) {
        cudaSetDevice(devices[0]);
        rc = cm_ivfpq_create(dim, metric, nlist, M, nbits, &h->assigner);
    }
    if (rc == CM_OK) rc = h->ls.init(dim, nlist, metric, devices, n_devices);
    cudaSetDevice(prev);
    if (rc != CM_OK) {
        std::string msg = cm_last_error();
        if (h->ls.dev.empty()) {          // init never ran: free the shards by hand
            for (size_t r = 0; r < h->shard.size(); r++) if (h->shard[r]) { cudaSetDevice(devices[r]); cm_ivfpq_destroy(h->shard[r]); }
            if (h->assigner) { cudaSetDevice(devices[0]); cm_ivfpq_destroy(h->assigner); }
            cudaSetDevice(prev);
            delete h;
        } else {
            cm_ivfpq_sharded_destroy(h);
        }
        return cm::fail(rc, "%s", msg.c_str());
    }
    *out = h;
    return CM_OK;
}

int cm_ivfpq_sharded_shards(const cm_ivfpq_sharded *h) { return h ? h->ls.W() : 0; }
int cm_ivfpq_sharded_trained(const cm_ivfpq_sharded *h) { return h && h->assigner && h->assigner->ix.trained ? 1 : 0; }
int64_t cm_ivfpq_sharded_size(const cm_ivfpq_sharded *h) {
    int64_t n = 0;
    if (h) for (cm_ivfpq *s : h->shard) n += cm_ivfpq_size(s);
    return n;
}
int cm_ivfpq_sharded_default_nprobes(const cm_ivfpq_sharded *h) { return h ? cm_ivfpq_default_nprobes(h->assigner) : 0; }
int cm_ivfpq_sharded_owner(const cm_ivfpq_sharded *h, int list) { return (h && list >= 0 && list < h->ls.nlist) ? h->ls.owner[(size_t)list] : -1; }
int cm_ivfpq_sharded_shard_size(const cm_ivfpq_sharded *h, int shard, int64_t *rows) {
    if (!h || shard < 0 || shard >= h->ls.W() || !rows) return cm::fail(CM_ERR_INVALID_ARG, "bad shard");
    *rows = cm_ivfpq_size(h->shard[(size_t)shard]);
    return CM_OK;
}

int cm_ivfpq_sharded_set_trained(cm_ivfpq_sharded *h, const float *centroids, const float *codebooks) {
    if (!h || !centroids || !codebooks) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (cm_ivfpq_sharded_size(h) > 0) return cm::fail(CM_ERR_INVALID_ARG, "cannot replace the trained state of a non-empty index");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = cm_ivfpq_set_trained(h->assigner, centroids, codebooks);
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) rc = cm_ivfpq_set_trained(h->shard[r], centroids, codebooks);
    cudaSetDevice(prev);
    return rc;
}
int cm_ivfpq_sharded_get_trained(const cm_ivfpq_sharded *h, float *centroids, float *codebooks) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    return cm_ivfpq_get_trained(h->assigner, centroids, codebooks);
}
int cm_ivfpq_sharded_train(cm_ivfpq_sharded *h, const float *rows, int64_t n) {     // IVFPQIndex.Train (ivfpq_index.go:180-259) on devices[0]
    if (!h || (n > 0 && !rows)) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (cm_ivfpq_sharded_size(h) > 0) return cm::fail(CM_ERR_UNSUPPORTED, "retraining a non-empty index is not supported");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = cm_ivfpq_train(h->assigner, rows, n);
    const cm::PQCore &a = h->assigner->ix;
    std::vector<float> cent((size_t)a.nlist * a.dim), books((size_t)a.M * a.Ksub * a.dsub);
    if (rc == CM_OK) rc = cm_ivfpq_get_trained(h->assigner, cent.data(), books.data());
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) rc = cm_ivfpq_set_trained(h->shard[r], cent.data(), books.data());
    cudaSetDevice(prev);
    return rc;
}

// n successive IVFPQIndex.Add calls (ivfpq_index.go:279-319): preprocess, nearest centroid, residual, encode -- once, on
// devices[0]; every (id, code) then goes to the shard that owns its list, lists keep arrival order.
int cm_ivfpq_sharded_add(cm_ivfpq_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists) {
    if (!h || (n > 0 && (!ids || !rows))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivfpq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before adding");
    if (n <= 0) return CM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cm::ListShards &ls = h->ls;
    const int W = ls.W(), dim = ls.dim, M = h->M;
    const int64_t slab = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)dim * 4));
    std::vector<int32_t> lists((size_t)std::min(slab, n));
    std::vector<uint8_t> codes;
    std::vector<std::vector<uint32_t>> sub_ids((size_t)W);
    std::vector<std::vector<int32_t>> sub_lists((size_t)W);
    std::vector<std::vector<uint8_t>> sub_codes((size_t)W);
    int rc = CM_OK, rc_zero = CM_OK;
    std::string zero_msg;
    for (int64_t i0 = 0; i0 < n && rc == CM_OK && rc_zero == CM_OK; i0 += slab) {
        const int64_t m = std::min(slab, n - i0);
        int rc_a = cm_ivfpq_add(h->assigner, ids + i0, rows + (size_t)i0 * dim, m, writeback, lists.data());
        const int64_t good = cm_ivfpq_size(h->assigner);
        if (rc_a == CM_ERR_ZERO_VECTOR) { rc_zero = rc_a; zero_msg = cm_last_error(); } else rc = rc_a;
        codes.resize((size_t)std::max<int64_t>(good, 1) * M);
        if (rc == CM_OK && good > 0) rc = cm_ivfpq_get_codes(h->assigner, 0, good, codes.data());
        if (rc == CM_OK) { cudaSetDevice(ls.dev[0]); rc = ivfpqs_clear_vectors(h->assigner->ix); }
        for (int r = 0; r < W; r++) { sub_ids[(size_t)r].clear(); sub_lists[(size_t)r].clear(); sub_codes[(size_t)r].clear(); }
        for (int64_t i = 0; i < good && rc == CM_OK; i++) {
            const int32_t l = lists[(size_t)i];
            const int r = ls.owner[(size_t)l];
            sub_ids[(size_t)r].push_back(ids[i0 + i]);
            sub_lists[(size_t)r].push_back(l);
            sub_codes[(size_t)r].insert(sub_codes[(size_t)r].end(), codes.begin() + (size_t)i * M, codes.begin() + (size_t)(i + 1) * M);
            ls.glob_len[(size_t)l]++;
            if (out_lists) out_lists[i0 + i] = l;
        }
        for (int r = 0; r < W && rc == CM_OK; r++)
            if (!sub_ids[(size_t)r].empty())
                rc = cm_ivfpq_load_codes(h->shard[(size_t)r], sub_ids[(size_t)r].data(), sub_codes[(size_t)r].data(), sub_lists[(size_t)r].data(),
                                         (int64_t)sub_ids[(size_t)r].size());
        ls.len_dirty = true;
    }
    cudaSetDevice(prev);
    if (rc != CM_OK) return rc;
    return rc_zero == CM_OK ? CM_OK : cm::fail(CM_ERR_ZERO_VECTOR, "%s", zero_msg.c_str());
}

int cm_ivfpq_sharded_remove(cm_ivfpq_sharded *h, uint32_t id) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    bool found = false, already = false;
    for (cm_ivfpq *s : h->shard) {
        int rc = cm_ivfpq_remove(s, id);
        if (rc == CM_OK) found = true;
        else if (rc == CM_ERR_NOT_FOUND && strstr(cm_last_error(), "already")) already = true;
        else if (rc != CM_ERR_NOT_FOUND) { cudaSetDevice(prev); return rc; }
    }
    cudaSetDevice(prev);
    if (found) return CM_OK;
    return cm::fail(CM_ERR_NOT_FOUND, already ? "vector with ID %u already deleted" : "vector with ID %u not found", id);
}

int cm_ivfpq_sharded_flush(cm_ivfpq_sharded *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    std::fill(h->ls.glob_len.begin(), h->ls.glob_len.end(), 0);
    for (size_t r = 0; r < h->shard.size() && rc == CM_OK; r++) {
        rc = cm_ivfpq_flush(h->shard[r]);
        for (int l = 0; l < h->ls.nlist; l++) h->ls.glob_len[(size_t)l] += (long long)h->shard[r]->ix.lists[(size_t)l].size();
    }
    h->ls.len_dirty = true;
    cudaSetDevice(prev);
    return rc;
}

// greedy-by-length list assignment; lists that change owner move whole (codes and ids), in their order
int cm_ivfpq_sharded_rebalance(cm_ivfpq_sharded *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    cm::ListShards &ls = h->ls;
    const int W = ls.W(), M = h->M;
    const std::vector<int> want = ls.greedy_plan();
    for (cm_ivfpq *s : h->shard)
        if (!s->ix.store.deleted_ids.empty()) return cm::fail(CM_ERR_UNSUPPORTED, "flush before rebalancing");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = CM_OK;
    for (int r = 0; r < W && rc == CM_OK; r++) {
        cm::PQCore &ix = h->shard[(size_t)r]->ix;
        const int64_t n = ix.store.n;
        bool any = false;
        for (int l = 0; l < ls.nlist && !any; l++) any = ls.owner[(size_t)l] == r && want[(size_t)l] != r && !ix.lists[(size_t)l].empty();
        if (!any) continue;
        std::vector<uint8_t> all((size_t)n * M);
        rc = cm_ivfpq_get_codes(h->shard[(size_t)r], 0, n, all.data());
        if (rc != CM_OK) break;
        const std::vector<uint32_t> ids_all = ix.store.ids_host;
        const std::vector<int32_t> lo_all = ix.list_of;
        const std::vector<std::vector<uint32_t>> lists_all = ix.lists;
        // target t receives the positions of the lists moving to it (list by list, each in its order); r keeps the rest
        for (int t = 0; t < W && rc == CM_OK; t++) {
            std::vector<uint32_t> ids;
            std::vector<int32_t> lo;
            std::vector<uint8_t> codes;
            if (t == r) {
                for (int64_t i = 0; i < n; i++) {
                    const int l = lo_all[(size_t)i];
                    if (ls.owner[(size_t)l] == r && want[(size_t)l] != r) continue;
                    ids.push_back(ids_all[(size_t)i]); lo.push_back(l);
                    codes.insert(codes.end(), all.begin() + (size_t)i * M, all.begin() + (size_t)(i + 1) * M);
                }
                cudaSetDevice(ls.dev[(size_t)r]);          // the shard's own device: its buffers are not mapped on the others
                rc = ivfpqs_clear_vectors(ix);
            } else {
                for (int l = 0; l < ls.nlist; l++)
                    if (ls.owner[(size_t)l] == r && want[(size_t)l] == t)
                        for (uint32_t pos : lists_all[(size_t)l]) {
                            ids.push_back(ids_all[pos]); lo.push_back(l);
                            codes.insert(codes.end(), all.begin() + (size_t)pos * M, all.begin() + (size_t)(pos + 1) * M);
                        }
            }
            if (rc == CM_OK && !ids.empty())
                rc = cm_ivfpq_load_codes(h->shard[(size_t)t], ids.data(), codes.data(), lo.data(), (int64_t)ids.size());
        }
    }
    if (rc == CM_OK) ls.owner = want;
    cudaSetDevice(prev);
    return rc;
}

static cm::ListShards::ShardSearch ivfpqs_search_fn(cm_ivfpq_sharded *h) {
    return [h](int r, const float *q, int64_t nq, const cm_search_params *p, int64_t K, uint32_t *o_ids, float *o_sc, int64_t *o_cnt,
               cudaStream_t s, const long long *glob_len, uint32_t *o_gno) {
        return cm::adc_search_device(h->shard[(size_t)r]->ix, q, nq, p, K, o_ids, o_sc, nullptr, o_cnt, s, false, glob_len, o_gno);
    };
}

int cm_ivfpq_sharded_search_device(cm_ivfpq_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                   int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev,
                                   void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivfpq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ls.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ls.dim, dim);
    if (nq <= 0) return CM_OK;
    return h->ls.search_device(ivfpqs_search_fn(h), queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_counts_dev,
                               (cudaStream_t)stream);
}

// nq independent searchSingleQuery calls (ivfpq_index_search.go:231-390) against the whole list-sharded index
int cm_ivfpq_sharded_search(cm_ivfpq_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                            int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (!cm_ivfpq_sharded_trained(h)) return cm::fail(CM_ERR_NOT_TRAINED, "index must be trained before searching");
    if (dim != h->ls.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ls.dim, dim);
    if (nq <= 0) return CM_OK;
    int rc = h->ls.search_host(ivfpqs_search_fn(h), queries, nq, p, out_stride, out_ids, out_scores, out_counts);
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t r = 0; r < h->shard.size(); r++) h->ls.last_scanned[r] = cm_ivfpq_last_scanned(h->shard[r]);
    cudaSetDevice(prev);
    return rc;
}

int cm_ivfpq_sharded_last_scanned(const cm_ivfpq_sharded *h, int64_t *per_shard) {
    if (!h || !per_shard) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    for (size_t r = 0; r < h->shard.size(); r++) per_shard[r] = h->ls.last_scanned[r];
    return CM_OK;
}

}  // extern "C"
