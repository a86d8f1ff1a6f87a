// hnsw.cu -- HNSWIndex device state, warp-per-query graph search, device-side insertion and the
// cm_hnsw_* entry points.
//
// Replaces hnswIndexSearch.searchSingleQuery (hnsw_index_search.go:248-354), HNSWIndex.searchLayer
// (hnsw_index.go:565-629) and HNSWIndex.Add / insertNode / selectNeighbors / pruneConnections
// (hnsw_index.go:228-288, 493-552, 637-694): greedy descent through the upper layers (K10), beam
// search with Go's container/heap semantics (K11: Push = append + sift-up, Pop = swap(0, n-1) +
// sift-down, strict comparisons -- the order of equal keys decides which node is expanded next, so
// the heaps are reproduced operation by operation), document filter and threshold applied after the
// traversal.  Results AND graphs are bit-identical to a sequential CPU replay of the reference given
// the same level draws (the reference draws levels from an unseeded global RNG, hnsw_index.go:474-484,
// so the caller supplies them).
//
// Two reference behaviours are reproduced on purpose (SURVEY 2.1, quirks b and the prune below):
//   * the entry point is the first node ever inserted and is never promoted (hnsw_index.go:273-278);
//   * while a node is being inserted it is not yet in idx.nodes, so pruneConnections of a neighbour
//     whose list overflowed skips -- i.e. drops -- the very back-edge that was just appended
//     (hnsw_index.go:540-545, 668-671).
//
// Device layout (replaces map[uint32]*hnswNode, hnsw_index.go:50-61, 111-155), by slot = insertion order:
//   rows fp32 [cap][ld], ids u32 [cap], deleted u8 [cap], levels i32 [cap]
//   adj0 u32 [cap][E0] + deg0 i32 [cap]            layer 0, E0 >= 2M + 1 (one slot for append-then-prune)
//   up_of i32 [cap] -> upper block or -1;  adjU u32 [cap_up][16][EU] + degU i32 [cap_up][16], EU >= M + 1
// One warp per query (or per insertion): the 32 lanes evaluate up to 32 neighbour distances at once --
// each lane walks its own row in the reference's sequential order -- and lane 0 replays the heap
// operations in edge order.  Visited set: one bit per node (global memory, atomicOr).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "flat_index.cuh"
#include "flat_kernels.cuh"
#include "wire.cuh"

namespace cm {

static constexpr int HNSW_WARPS = 4;
static constexpr int HNSW_MAX_LEVELS = 17;      // levels 0..16 (hnsw_index.go:474-484 caps the draw at 16)

struct HCand { float d; uint32_t slot; };

// what the kernels see of the graph
struct GraphView {
    const float *rows; int ld;
    const uint32_t *ids; const uint8_t *deleted; const int *levels;
    uint32_t *adj0; int *deg0; int E0;
    const int *up_of; uint32_t *adjU; int *degU; int EU;

    __device__ __forceinline__ uint32_t *edges(long long slot, int layer, int **deg_out) const {
        if (layer == 0) { *deg_out = deg0 + slot; return adj0 + (size_t)slot * E0; }
        int u = up_of[slot];
        *deg_out = degU + (size_t)u * HNSW_MAX_LEVELS + layer;
        return adjU + ((size_t)u * HNSW_MAX_LEVELS + layer) * EU;
    }
};

struct HNSWIndex {
    int dim = 0, ld = 0, metric = 0, m = 0, efc = 0, efs = 0, device = 0;
    int64_t n = 0, cap = 0, n_up = 0, cap_up = 0;
    int E0 = 0, EU = 0;
    float *rows = nullptr;
    uint32_t *ids = nullptr;
    uint8_t *deleted = nullptr;
    int *levels = nullptr, *deg0 = nullptr, *up_of = nullptr, *degU = nullptr;
    uint32_t *adj0 = nullptr, *adjU = nullptr;
    long long entry_slot = -1;
    int max_level = -1;
    std::unordered_map<uint32_t, int64_t> slot_of;
    std::unordered_set<uint32_t> deleted_ids;
    std::vector<int> levels_host;

    void free_dev() {
        cudaFree(rows); cudaFree(ids); cudaFree(deleted); cudaFree(levels); cudaFree(deg0); cudaFree(up_of); cudaFree(degU);
        cudaFree(adj0); cudaFree(adjU);
        rows = nullptr; ids = nullptr; deleted = nullptr; levels = nullptr; deg0 = nullptr; up_of = nullptr; degU = nullptr;
        adj0 = nullptr; adjU = nullptr;
        cap = cap_up = 0;
    }
    ~HNSWIndex() { free_dev(); }
    GraphView view() const { return GraphView{rows, ld, ids, deleted, levels, adj0, deg0, E0, up_of, adjU, degU, EU}; }
    int reserve(int64_t want, int64_t want_up);
};

template <typename T>
static int grow(T **p, int64_t old_n, int64_t new_cap, size_t per, bool zero) {
    T *np = nullptr;
    CM_CUDA(cudaMalloc(&np, (size_t)new_cap * per * sizeof(T)));
    if (zero) CM_CUDA(cudaMemset(np, 0, (size_t)new_cap * per * sizeof(T)));
    if (old_n > 0 && *p) CM_CUDA(cudaMemcpy(np, *p, (size_t)old_n * per * sizeof(T), cudaMemcpyDeviceToDevice));
    cudaFree(*p);
    *p = np;
    return CM_OK;
}

int HNSWIndex::reserve(int64_t want, int64_t want_up) {
    if (want > cap) {
        int64_t nc = cap ? cap : 1024;
        while (nc < want) nc = nc + nc / 2 + 1024;
        CM_TRY(grow(&rows, n, nc, (size_t)ld, true));
        CM_TRY(grow(&ids, n, nc, 1, true));
        CM_TRY(grow(&deleted, n, nc, 1, true));
        CM_TRY(grow(&levels, n, nc, 1, true));
        CM_TRY(grow(&deg0, n, nc, 1, true));
        CM_TRY(grow(&up_of, n, nc, 1, true));
        CM_TRY(grow(&adj0, n, nc, (size_t)E0, true));
        cap = nc;
    }
    if (want_up > cap_up) {
        int64_t nc = cap_up ? cap_up : 256;
        while (nc < want_up) nc = nc + nc / 2 + 256;
        CM_TRY(grow(&degU, n_up, nc, HNSW_MAX_LEVELS, true));
        CM_TRY(grow(&adjU, n_up, nc, (size_t)HNSW_MAX_LEVELS * EU, true));
        cap_up = nc;
    }
    return CM_OK;
}

// ---- Go container/heap on (distance, slot) pairs; IS_MAX: less(i, j) = d[i] > d[j], else d[i] < d[j] ----
template <bool IS_MAX>
__device__ __forceinline__ bool h_less(const HCand *a, int i, int j) {
    return IS_MAX ? (a[i].d > a[j].d) : (a[i].d < a[j].d);
}
template <bool IS_MAX>
__device__ __forceinline__ void h_push(HCand *a, int &n, HCand c) {
    a[n++] = c;
    int j = n - 1;
    for (;;) {                                   // heap.up
        int i = (j - 1) / 2;
        if (i == j || !h_less<IS_MAX>(a, j, i)) break;
        HCand t = a[i]; a[i] = a[j]; a[j] = t;
        j = i;
    }
}
template <bool IS_MAX>
__device__ __forceinline__ HCand h_pop(HCand *a, int &n) {
    int last = n - 1;
    HCand t = a[0]; a[0] = a[last]; a[last] = t;
    int i = 0;
    for (;;) {                                   // heap.down(0, last)
        int j1 = 2 * i + 1;
        if (j1 >= last || j1 < 0) break;
        int j = j1, j2 = j1 + 1;
        if (j2 < last && h_less<IS_MAX>(a, j2, j1)) j = j2;
        if (!h_less<IS_MAX>(a, j, i)) break;
        HCand u = a[i]; a[i] = a[j]; a[j] = u;
        i = j;
    }
    n = last;
    return a[last];
}

// heap.Push of Go's container/heap by the whole warp, same result as the serial h_push: sift-up compares the NEW element
// with its successive parents and stops at the first one it is not less than -- so every ancestor on the path to the
// root is loaded by its own lane at once, a ballot finds how far the element climbs, the ancestors below that point move
// down one level and the element lands.  One round of loads instead of a dependent chain of log2(n) (lane 0 replaying the
// heaps serially was most of an expansion's time).  Must be called by all 32 lanes with identical arguments.
template <bool IS_MAX>
__device__ __forceinline__ void h_push_warp(HCand *a, int &n, HCand c, int lane) {
    const int j = n;
    const int D = 31 - __clz(j + 1);                      // ancestors of index j: ((j + 1) >> l) - 1, l = 1 .. D
    HCand anc{0.0f, 0u};
    bool lt = false;
    if (lane >= 1 && lane <= D) {
        anc = a[((j + 1) >> lane) - 1];
        lt = IS_MAX ? (c.d > anc.d) : (c.d < anc.d);      // less(new, ancestor)
    }
    const unsigned m = __ballot_sync(0xffffffffu, lt);
    __syncwarp();                                         // every ancestor is read before any slot is rewritten
    const int up = __ffs(~(m >> 1)) - 1;                  // leading run of "less than the parent": levels climbed
    if (lane >= 1 && lane <= up) a[((j + 1) >> (lane - 1)) - 1] = anc;
    if (lane == 0) a[((j + 1) >> up) - 1] = c;
    __syncwarp();
    n = j + 1;
}

// Distance.Calculate(a = vector in shared memory, b = row) by ONE lane in the reference's order
template <int METRIC, bool FMA>
__device__ __forceinline__ float row_distance(const float *__restrict__ row, const float *__restrict__ q_s, int ld) {
    const float4 *x = reinterpret_cast<const float4 *>(row);
    const float4 *q = reinterpret_cast<const float4 *>(q_s);
    float acc = 0.0f;
#pragma unroll 4
    for (int j = 0; j < ld / 4; j++) {
        float4 xv = __ldg(x + j), qv = q[j];
        acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
        acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
        acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
        acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
    }
    return metric_finish<METRIC>(acc);
}
// same for rows that were written earlier in the SAME kernel (insertion): no read-only path
template <int METRIC, bool FMA>
__device__ __forceinline__ float row_distance_rw(const float *row, const float *q_s, int ld) {
    const float4 *x = reinterpret_cast<const float4 *>(row);
    const float4 *q = reinterpret_cast<const float4 *>(q_s);
    float acc = 0.0f;
#pragma unroll 4
    for (int j = 0; j < ld / 4; j++) {
        float4 xv = x[j], qv = q[j];
        acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
        acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
        acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
        acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
    }
    return metric_finish<METRIC>(acc);
}

// Distances of up to 32 rows (one per lane, `need` = this lane's row counts) to the vector in q_s, each in the
// reference's order, with the rows STAGED through shared memory: per step the warp copies the next 128 bytes of every
// needed row with 16-byte cp.async (a quarter-warp per row: coalesced, two steps in flight) and each lane then walks
// its own row's piece.  The per-lane `row_distance` walk keeps only 64 bytes per lane in flight and re-fetches lines
// through L1 when many warps share an SM.  Must be called by all 32 lanes; rows must not be written by this kernel.
template <int METRIC, bool FMA, int NS>
__device__ __forceinline__ float warp_row_distances(const float *__restrict__ rows, int ld, uint32_t nb, bool need,
                                                    const float *__restrict__ q_s, uint8_t *stage, int lane) {
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0u) return 0.0f;
    // the eight (row, piece) pairs this lane copies every step: rows p*4 + lane/8, piece lane%8
    const int piece = lane & 7;
    const float *src[8];
    uint32_t dst_off[8];
    bool cp[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int r = p * 4 + (lane >> 3);
        const uint32_t nbr = __shfl_sync(0xffffffffu, nb, r);
        cp[p] = (mask >> r) & 1u;
        src[p] = rows + (size_t)nbr * ld + piece * 4;
        dst_off[p] = (uint32_t)(r * 128 + ((piece ^ (r & 7)) << 4));
    }
    const int n_ch = ld / 32;
    const uint32_t stage_u = smem_u32(stage);
    // NS - 1 steps (4 KB each: 128 bytes of every needed row) in flight: an expansion is a chain of ld / 32 dependent
    // steps, each an HBM round trip, and with few queries per SM nothing else hides that latency.  Every iteration
    // commits exactly one (possibly empty) group so the wait below is the same constant throughout.
    auto issue = [&](int c) {
        if (c < n_ch) {
            const uint32_t base = stage_u + (uint32_t)(c % NS) * 4096u;
#pragma unroll
            for (int p = 0; p < 8; p++)
                if (cp[p]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + dst_off[p]), "l"(src[p] + c * 32) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int c = 0; c < NS - 1; c++) issue(c);
    float acc = 0.0f;
    for (int c = 0; c < n_ch; c++) {
        issue(c + NS - 1);                       // into the slot the previous iteration finished reading
        asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
        __syncwarp();
        if (need) {
            const uint8_t *sp = stage + (size_t)(c % NS) * 4096 + lane * 128;
            const float *qc = q_s + c * 32;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 xv = *reinterpret_cast<const float4 *>(sp + ((j ^ (lane & 7)) << 4));
                const float4 qv = *reinterpret_cast<const float4 *>(qc + j * 4);
                acc = metric_step<METRIC, FMA>(acc, qv.x, xv.x);
                acc = metric_step<METRIC, FMA>(acc, qv.y, xv.y);
                acc = metric_step<METRIC, FMA>(acc, qv.z, xv.z);
                acc = metric_step<METRIC, FMA>(acc, qv.w, xv.w);
            }
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return metric_finish<METRIC>(acc);
}

// per-warp scratch in shared memory
struct WarpScratch {
    float *q_s;        // [ld] the query / the vector being inserted
    HCand *res;        // [ef + 1] result max-heap
    float *nb_d;       // [32]
    uint32_t *nb_slot; // [32]
    uint32_t *nb_new;  // [32]
    uint8_t *stage;    // [stages][32 rows][128 B] row pieces of one expansion's neighbours (16-byte pieces XOR-swizzled)
};
static constexpr int HNSW_STAGE_STEP = 32 * 128;       // one step: 128 bytes of 32 rows
__host__ __device__ inline size_t warp_scratch_head(int ld, int ef) {
    return (((size_t)ld * 4 + (size_t)(ef + 1) * sizeof(HCand) + 32 * 12) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t warp_scratch_bytes(int ld, int ef, int stages = 2) {
    return warp_scratch_head(ld, ef) + (size_t)stages * HNSW_STAGE_STEP;
}
__device__ __forceinline__ WarpScratch carve(uint8_t *base, int ld, int ef) {
    WarpScratch w;
    w.q_s = reinterpret_cast<float *>(base);
    w.res = reinterpret_cast<HCand *>(w.q_s + ld);
    w.nb_d = reinterpret_cast<float *>(w.res + (ef + 1));
    w.nb_slot = reinterpret_cast<uint32_t *>(w.nb_d + 32);
    w.nb_new = w.nb_slot + 32;
    w.stage = base + warp_scratch_head(ld, ef);
    return w;
}

// greedy descent from `curr` through layers hi .. lo+1 (hnsw_index_search.go:271-296 / hnsw_index.go:498-521)
template <int METRIC, bool FMA>
__device__ __forceinline__ void greedy_descend(const GraphView &G, const float *q_s, int hi, int lo, long long &curr,
                                               float &curr_dist, long long &evals, int lane) {
    for (int lc = hi; lc > lo; lc--) {
        bool changed = true;
        while (changed) {
            changed = false;
            if (lc > G.levels[curr]) break;                     // lc < len(node.Edges)
            int *degp;
            const uint32_t *E = G.edges(curr, lc, &degp);
            const int deg = *degp;
            for (int b = 0; b < deg; b += 32) {
                int j = b + lane;
                uint32_t nb = 0;
                bool valid = false;
                float d = 0.0f;
                if (j < deg) {
                    nb = E[j];
                    valid = G.deleted[nb] == 0;
                    if (valid) d = row_distance_rw<METRIC, FMA>(G.rows + (size_t)nb * G.ld, q_s, G.ld);
                }
                int cnt = min(32, deg - b);
                for (int t = 0; t < cnt; t++) {
                    float dt = __shfl_sync(0xffffffffu, d, t);
                    bool vt = __shfl_sync(0xffffffffu, (int)valid, t) != 0;
                    uint32_t nt = __shfl_sync(0xffffffffu, nb, t);
                    if (vt) {
                        evals++;
                        if (dt < curr_dist) { curr_dist = dt; curr = nt; changed = true; }
                    }
                }
            }
        }
    }
}

// searchLayer(query, entry, ef, layer), hnsw_index.go:565-629.  Leaves the results ASCENDING in `sorted`
// (= the cands buffer, whose heap is dead by then) and returns their number, or -1 on candidate-heap overflow.
// touched (optional): every slot whose visited bit was set is appended, so the caller can clear the bits.
template <int METRIC, bool FMA, bool STAGED, int NS = 2>
__device__ __forceinline__ int search_layer(const GraphView &G, const WarpScratch &W, long long entry, int ef, int layer,
                                            uint32_t *vis, HCand *cands, int cand_cap, uint32_t *touched, int *n_touched,
                                            long long &evals, long long &expansions, int lane) {
    int n_c = 0, n_r = 0;                 // identical on every lane: pushes are warp-cooperative, pops run on lane 0
    bool overflow = false;
    if (G.deleted[entry] == 0) {
        float d = 0.0f;
        if (lane == 0) d = row_distance_rw<METRIC, FMA>(G.rows + (size_t)entry * G.ld, W.q_s, G.ld);
        d = __shfl_sync(0xffffffffu, d, 0);
        const HCand c{d, (uint32_t)entry};
        h_push_warp<false>(cands, n_c, c, lane);
        h_push_warp<true>(W.res, n_r, c, lane);
        evals++;
    }
    if (lane == 0) {
        atomicOr(&vis[entry >> 5], 1u << (entry & 31));
        if (touched) touched[(*n_touched)++] = (uint32_t)entry;
    }
    __syncwarp();
    for (;;) {
        uint32_t cur_slot = 0;
        int go = 0;
        if (n_c > 0) {
            if (lane == 0) {
                int nn = n_c;
                HCand cur = h_pop<false>(cands, nn);
                if (!(n_r >= ef && cur.d > W.res[0].d)) { go = 1; cur_slot = cur.slot; }
            }
            n_c--;
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) break;
        cur_slot = __shfl_sync(0xffffffffu, cur_slot, 0);
        expansions++;
        int deg = 0, width = 0;
        uint32_t *E = nullptr;
        if (layer == 0 || layer <= G.levels[cur_slot]) {            // layer < len(node.Edges); every node has layer 0
            int *degp;
            E = G.edges(cur_slot, layer, &degp);
            width = layer == 0 ? G.E0 : G.EU;                       // the adjacency row is that wide whatever its degree:
            deg = *degp;                                            // degree and neighbours are fetched together below
        }
        for (int b = 0; b < width; b += 32) {
            int j = b + lane;
            uint32_t nb = 0, isnew = 0;
            float d = 0.0f;
            const uint32_t nb_raw = j < width ? E[j] : 0u;
            if (b >= deg) break;                                    // uniform: deg is the same on every lane
            if (j < deg) {
                nb = nb_raw;
                const uint32_t bit = 1u << (nb & 31);
                if (STAGED) {
                    // search: the deleted flag and the visited bit travel together; marking a deleted node visited changes
                    // nothing (it is skipped whenever it is met) and the insertion path, which records what it touched, keeps
                    // the reference's order of tests
                    const uint8_t del = G.deleted[nb];
                    const uint32_t old = atomicOr(&vis[nb >> 5], bit);
                    isnew = del == 0 && (old & bit) == 0u;
                } else if (G.deleted[nb] == 0) {
                    const uint32_t old = atomicOr(&vis[nb >> 5], bit);
                    isnew = (old & bit) == 0u;
                }
                if (!STAGED && isnew) d = row_distance_rw<METRIC, FMA>(G.rows + (size_t)nb * G.ld, W.q_s, G.ld);
            }
            if (STAGED) d = warp_row_distances<METRIC, FMA, NS>(G.rows, G.ld, nb, isnew != 0u, W.q_s, W.stage, lane);
            W.nb_d[lane] = d; W.nb_slot[lane] = nb; W.nb_new[lane] = isnew;
            evals += __popc(__ballot_sync(0xffffffffu, isnew != 0u));
            __syncwarp();
            const int cnt = min(32, deg - b);
            for (int t = 0; t < cnt; t++) {                         // every lane walks the neighbours in order
                if (!W.nb_new[t]) continue;
                if (touched && lane == 0) touched[(*n_touched)++] = W.nb_slot[t];
                const float dt = W.nb_d[t];
                if (n_r < ef || dt < W.res[0].d) {
                    const HCand c{dt, W.nb_slot[t]};
                    if (n_c >= cand_cap) { overflow = true; break; }
                    h_push_warp<false>(cands, n_c, c, lane);
                    h_push_warp<true>(W.res, n_r, c, lane);
                    if (n_r > ef) {
                        if (lane == 0) { int nn = n_r; (void)h_pop<true>(W.res, nn); }
                        n_r--;
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
            if (overflow) break;
        }
        if (overflow) break;
    }
    int n = -1;
    if (lane == 0 && !overflow) {
        n = n_r;
        int nn = n_r;
        for (int i = n - 1; i >= 0; i--) cands[i] = h_pop<true>(W.res, nn);      // :622-626 ascending
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();
    return n;
}

template <int METRIC, bool FMA, int NS>
__global__ void __launch_bounds__(HNSW_WARPS * 32) hnsw_search_kernel(
    GraphView G, long long entry_slot, int max_level, const float *__restrict__ queries, int nq, int ef, long long k_req,
    float threshold, const uint8_t *__restrict__ doc_skip, uint32_t *__restrict__ visited, long long vis_words,
    HCand *__restrict__ cand_heaps, int cand_cap, long long out_stride, uint32_t *__restrict__ out_ids,
    float *__restrict__ out_scores, long long *__restrict__ out_pos, long long *__restrict__ out_counts,
    long long *__restrict__ work /* [nq][2]: distance evaluations, expansions */) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * HNSW_WARPS + warp;
    if (q >= nq) return;
    WarpScratch W = carve(smem + (size_t)warp * warp_scratch_bytes(G.ld, ef, NS), G.ld, ef);
    for (int j = lane; j < G.ld; j += 32) W.q_s[j] = queries[(size_t)q * G.ld + j];
    __syncwarp();
    HCand *cands = cand_heaps + (size_t)q * cand_cap;
    uint32_t *vis = visited + (size_t)q * vis_words;
    long long evals = 0, expansions = 0;

    // phase 1: greedy descent, hnsw_index_search.go:271-296
    long long curr = entry_slot;
    float curr_dist = 0.0f;
    if (lane == 0) curr_dist = row_distance_rw<METRIC, FMA>(G.rows + (size_t)curr * G.ld, W.q_s, G.ld);
    curr_dist = __shfl_sync(0xffffffffu, curr_dist, 0);
    evals++;
    greedy_descend<METRIC, FMA>(G, W.q_s, max_level, 0, curr, curr_dist, evals, lane);

    // phase 2: searchLayer(query, curr, ef, 0)
    int n = search_layer<METRIC, FMA, true, NS>(G, W, curr, ef, 0, vis, cands, cand_cap, nullptr, nullptr, evals, expansions, lane);

    // results: post-filter (hnsw_index_search.go:321-335), already ascending, first k
    if (lane == 0) {
        long long count = -1;
        if (n >= 0) {
            HCand *sorted = cands;
            long long kept = 0;
            for (int i = 0; i < n; i++) {
                uint32_t s = sorted[i].slot;
                if (doc_skip != nullptr && doc_skip[s]) continue;                    // docFilter.ShouldSkip
                if (threshold > 0.0f && sorted[i].d > threshold) continue;
                sorted[kept++] = sorted[i];
            }
            long long k = (k_req <= 0 || k_req > kept) ? kept : k_req;               // sanitizeK(k, len(results))
            if (k > out_stride) k = out_stride;
            for (long long i = 0; i < k; i++) {
                size_t o = (size_t)q * out_stride + i;
                out_ids[o] = G.ids[sorted[i].slot];
                out_scores[o] = sorted[i].d;
                if (out_pos) out_pos[o] = sorted[i].slot;
            }
            count = k;
        }
        out_counts[q] = count;
        if (work) { work[(size_t)q * 2] = evals; work[(size_t)q * 2 + 1] = expansions; }
    }
}

// ------------------------------------------------------------------------------------------------
// insertion: one warp inserts the slots [s0, s0 + m) one after the other (each insertion sees the graph
// the previous ones left -- the reference holds idx.mu for the whole Add)
// ------------------------------------------------------------------------------------------------
// pruneConnections(nb, layer, M) (hnsw_index.go:658-694) while `skip` (the node being inserted) is not in
// idx.nodes yet: its edge is dropped, the others are re-sorted by distance to nb (stable) and cut to M.
template <int METRIC, bool FMA>
__device__ __forceinline__ void prune(const GraphView &G, const WarpScratch &W, float *nbrow_s, long long nb, int layer, int M,
                                      uint32_t skip, HCand *tmp, int lane) {
    int *degp;
    uint32_t *E = G.edges(nb, layer, &degp);
    const int deg = *degp;
    for (int j = lane; j < G.ld; j += 32) nbrow_s[j] = G.rows[(size_t)nb * G.ld + j];
    __syncwarp();
    for (int b = 0; b < deg; b += 32) {
        int j = b + lane;
        if (j < deg) {
            uint32_t e = E[j];
            float d = 0.0f;
            if (e != skip) d = row_distance_rw<METRIC, FMA>(G.rows + (size_t)e * G.ld, nbrow_s, G.ld);
            tmp[j] = HCand{d, e};
        }
    }
    __syncwarp();
    if (lane == 0) {
        int n = 0;
        for (int j = 0; j < deg; j++) {                         // stable insertion sort of the known nodes
            HCand c = tmp[j];
            if (c.slot == skip) continue;
            int i = n++;
            while (i > 0 && tmp[deg + i - 1].d > c.d) { tmp[deg + i] = tmp[deg + i - 1]; i--; }
            tmp[deg + i] = c;
        }
        int keep = n < M ? n : M;
        for (int i = 0; i < keep; i++) E[i] = tmp[deg + i].slot;
        *degp = keep;
    }
    __syncwarp();
}

template <int METRIC, bool FMA>
__global__ void __launch_bounds__(32) hnsw_insert_kernel(GraphView G, long long s0, int m, long long entry_slot, int max_level,
                                                         int M, int efc, uint32_t *__restrict__ vis, HCand *__restrict__ cands,
                                                         int cand_cap, uint32_t *__restrict__ touched, HCand *__restrict__ tmp,
                                                         int *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x;
    WarpScratch W = carve(smem, G.ld, efc);
    float *nbrow_s = reinterpret_cast<float *>(smem + warp_scratch_bytes(G.ld, efc));
    int max_l = max_level;
    for (int i = 0; i < m; i++) {
        const long long s = s0 + i;
        const int L = G.levels[s];
        if (L > max_l) max_l = L;                               // hnsw_index.go:266-268: before insertNode
        if (s == 0 && entry_slot < 0) continue;                 // :273-278 the first node only becomes the entry point
        const long long entry = entry_slot < 0 ? 0 : entry_slot;
        for (int j = lane; j < G.ld; j += 32) W.q_s[j] = G.rows[(size_t)s * G.ld + j];
        __syncwarp();
        long long evals = 0, expansions = 0;
        long long curr = entry;
        float curr_dist = 0.0f;
        if (lane == 0) curr_dist = row_distance_rw<METRIC, FMA>(G.rows + (size_t)curr * G.ld, W.q_s, G.ld);
        curr_dist = __shfl_sync(0xffffffffu, curr_dist, 0);
        greedy_descend<METRIC, FMA>(G, W.q_s, max_l, L, curr, curr_dist, evals, lane);
        for (int lc = L; lc >= 0; lc--) {
            int n_touched = 0;
            int nc = search_layer<METRIC, FMA, false>(G, W, curr, efc, lc, vis, cands, cand_cap, touched, &n_touched, evals, expansions, lane);
            n_touched = __shfl_sync(0xffffffffu, n_touched, 0);
            for (int t = lane; t < n_touched; t += 32) vis[touched[t] >> 5] = 0u;   // a fresh visited set per searchLayer
            __syncwarp();
            if (nc < 0) { if (lane == 0) *status = 1; return; }
            const int Mx = lc == 0 ? 2 * M : M;
            const int nsel = nc < Mx ? nc : Mx;                 // selectNeighbors: candidates are already ascending
            int *my_degp;
            uint32_t *my_E = G.edges(s, lc, &my_degp);
            for (int t = 0; t < nsel; t++) {
                const uint32_t nb = cands[t].slot;
                if (lane == 0) { my_E[*my_degp] = nb; (*my_degp)++; }
                if (lc <= G.levels[nb]) {
                    int *degp;
                    uint32_t *E = G.edges(nb, lc, &degp);
                    if (lane == 0) { E[*degp] = (uint32_t)s; (*degp)++; }
                    __syncwarp();
                    if (*degp > Mx) prune<METRIC, FMA>(G, W, nbrow_s, nb, lc, Mx, (uint32_t)s, tmp, lane);
                }
                __syncwarp();
            }
            if (nc > 0) curr = cands[0].slot;                   // :548-550
        }
    }
}

static int hnsw_search_device(HNSWIndex &ix, const float *q_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                              uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, int64_t *work,
                              cudaStream_t st, bool check_zero) {
    WsScope ws(st);
    if (nq <= 0) return CM_OK;
    if (ix.n == 0 || ix.max_level == -1) {                                           // hnsw_index_search.go:258-260
        CM_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)nq * sizeof(int64_t), st));
        return CM_OK;
    }
    int ef = p->ef_search > 0 ? p->ef_search : ix.efs;                               // :302-305
    int64_t k_bound = (p->k <= 0 || p->k > ef) ? ef : p->k;
    if (out_stride < std::min<int64_t>(k_bound, ix.n))
        return fail(CM_ERR_BUFFER_TOO_SMALL, "out_stride %lld < min(k, ef, n) = %lld", (long long)out_stride,
                    (long long)std::min<int64_t>(k_bound, ix.n));
    bool fma = rounding_mode() == CM_ROUND_FMA;
    const int ld = ix.ld;
    float *qp = nullptr;
    int *qflags = nullptr;
    CM_TRY(ws.get(&qp, (size_t)nq * ld * 4));
    CM_TRY(ws.get(&qflags, (size_t)nq * sizeof(int)));
    CM_TRY(launch_preprocess_rows(ix.metric, fma, q_dev, nq, ix.dim, ix.dim, qp, ld, qflags, st));
    if (check_zero && ix.metric == CM_COSINE) {
        std::vector<int> hf((size_t)nq);
        CM_CUDA(cudaMemcpyAsync(hf.data(), qflags, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < nq; i++)
            if (hf[(size_t)i]) {
                return fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (query %lld)", (long long)i);
            }
    }
    // document filter -> per-slot mask (soft deletes are handled inside the traversal, like the reference)
    uint8_t *doc_skip = nullptr;
    uint32_t *filt_dev = nullptr;
    if (p->filter_ids && p->nfilter > 0) {
        std::vector<uint32_t> f(p->filter_ids, p->filter_ids + p->nfilter);
        std::sort(f.begin(), f.end());
        f.erase(std::unique(f.begin(), f.end()), f.end());
        CM_TRY(ws.get(&filt_dev, f.size() * 4));
        CM_TRY(ws.get(&doc_skip, (size_t)ix.n));
        CM_CUDA(cudaMemcpyAsync(filt_dev, f.data(), f.size() * 4, cudaMemcpyHostToDevice, st));
        CM_TRY(launch_build_skip(ix.ids, nullptr, ix.n, filt_dev, (int64_t)f.size(), doc_skip, st));
        CM_CUDA(cudaStreamSynchronize(st));
    }
    const long long vis_words = (ix.n + 31) / 32;
    const int cand_cap = (int)std::min<int64_t>(ix.n + 1, (int64_t)16 * ef + 4096);
    // Row-staging depth per query warp: few queries per SM -> deep (the traversal is a chain of HBM round trips, only
    // more bytes in flight per warp hide them), many -> shallow (resident warps hide them, shared memory is the limit)
    // (measured on 1M x 768, ef 128: 512 queries 129K / 167K / 154K q/s with 2 / 4 / 8 steps; 8192 queries 385K / 316K with 2 / 4)
    int stages = nq <= (int64_t)sm_count() * 12 ? 4 : 2;
    if (const char *e = getenv("COMET_B200_HNSW_STAGES")) { int v = atoi(e); stages = v >= 8 ? 8 : (v >= 4 ? 4 : 2); }
    while (stages > 2 && warp_scratch_bytes(ld, ef, stages) * HNSW_WARPS > max_smem_optin()) stages /= 2;
    size_t smem = warp_scratch_bytes(ld, ef, stages) * HNSW_WARPS;
    if (smem > max_smem_optin()) return fail(CM_ERR_UNSUPPORTED, "efSearch %d with dim %d does not fit shared memory", ef, ix.dim);
    // queries in groups so that the visited bitmaps stay bounded (<= 1 GiB)
    int64_t qgroup = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)(1ull << 30) / (vis_words * 4 + (int64_t)cand_cap * 8)));
    uint32_t *visited = nullptr;
    HCand *heaps = nullptr;
    CM_TRY(ws.get(&visited, (size_t)qgroup * vis_words * 4));
    CM_TRY(ws.get(&heaps, (size_t)qgroup * cand_cap * sizeof(HCand)));
    GraphView G = ix.view();
    for (int64_t q0 = 0; q0 < nq; q0 += qgroup) {
        int64_t m = std::min(qgroup, nq - q0);
        CM_CUDA(cudaMemsetAsync(visited, 0, (size_t)m * vis_words * 4, st));
        unsigned blocks = (unsigned)((m + HNSW_WARPS - 1) / HNSW_WARPS);
        ProfScope prof(CM_PROF_HNSW, st);
#define CM_HNSW_LAUNCH_NS(MM, F, NS)                                                                                  \
    do {                                                                                                              \
        CM_TRY(set_dyn_smem((const void *)hnsw_search_kernel<MM, F, NS>, smem));                                      \
        hnsw_search_kernel<MM, F, NS><<<blocks, HNSW_WARPS * 32, smem, st>>>(                                         \
            G, ix.entry_slot, ix.max_level, qp + (size_t)q0 * ld, (int)m, ef, (long long)p->k, p->threshold, doc_skip, visited, \
            vis_words, heaps, cand_cap, (long long)out_stride, out_ids + (size_t)q0 * out_stride,                     \
            out_scores + (size_t)q0 * out_stride, out_pos ? (long long *)out_pos + (size_t)q0 * out_stride : nullptr,  \
            (long long *)out_counts + q0, work ? (long long *)work + (size_t)q0 * 2 : nullptr);                       \
    } while (0)
#define CM_HNSW_LAUNCH(MM, F)                                                                                         \
    do {                                                                                                              \
        if (stages == 8) CM_HNSW_LAUNCH_NS(MM, F, 8);                                                                 \
        else if (stages == 4) CM_HNSW_LAUNCH_NS(MM, F, 4);                                                            \
        else CM_HNSW_LAUNCH_NS(MM, F, 2);                                                                             \
    } while (0)
        switch (ix.metric) {
        case CM_L2: if (fma) CM_HNSW_LAUNCH(CM_L2, true); else CM_HNSW_LAUNCH(CM_L2, false); break;
        case CM_L2SQ: if (fma) CM_HNSW_LAUNCH(CM_L2SQ, true); else CM_HNSW_LAUNCH(CM_L2SQ, false); break;
        default: if (fma) CM_HNSW_LAUNCH(CM_COSINE, true); else CM_HNSW_LAUNCH(CM_COSINE, false); break;
        }
#undef CM_HNSW_LAUNCH_NS
#undef CM_HNSW_LAUNCH
        count_launch();
        CM_CUDA(cudaGetLastError());
    }
    return CM_OK;
}

}  // namespace cm

struct cm_hnsw {
    cm::HNSWIndex ix;
};

extern "C" {

int cm_hnsw_create(int dim, int metric, int m, int ef_construction, int ef_search, cm_hnsw **out) {
    if (!out) return cm::fail(CM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return cm::fail(CM_ERR_INVALID_ARG, "dimension must be positive");                 // hnsw_index.go:173
    if (metric < 0 || metric > 2) return cm::fail(CM_ERR_INVALID_ARG, "unknown distance kind");
    CM_TRY(cm::ensure_device());
    cm_hnsw *h = new cm_hnsw();
    // hnsw_index.go:178-190 defaults: M 16, efConstruction 200, efSearch = efConstruction
    h->ix.dim = dim; h->ix.metric = metric;
    h->ix.ld = (dim + cm::SCAN_CHUNK - 1) / cm::SCAN_CHUNK * cm::SCAN_CHUNK;
    h->ix.m = m > 0 ? m : 16;
    h->ix.efc = ef_construction > 0 ? ef_construction : 200;
    h->ix.efs = ef_search > 0 ? ef_search : h->ix.efc;
    h->ix.E0 = 2 * h->ix.m + 1;
    h->ix.EU = h->ix.m + 1;
    cudaGetDevice(&h->ix.device);
    *out = h;
    return CM_OK;
}
int cm_hnsw_destroy(cm_hnsw *h) {
    delete h;
    return CM_OK;
}
int64_t cm_hnsw_size(const cm_hnsw *h) { return h ? h->ix.n : 0; }
int cm_hnsw_ef_search(const cm_hnsw *h) { return h ? h->ix.efs : 0; }
int cm_hnsw_max_level(const cm_hnsw *h) { return h ? h->ix.max_level : -1; }

// Upload a graph built elsewhere by HNSWIndex.Add / insertNode (hnsw_index.go:228-288, 493-552), e.g. the state
// decoded by HNSWIndex.ReadFrom.  rows are the STORED vectors (already preprocessed by Add); slots are insertion
// order; edge_off has sum(levels[i] + 1) + 1 entries, pairs ordered by (slot, layer); edge_ids are neighbour IDs.
int cm_hnsw_load_graph(cm_hnsw *h, int64_t n, const uint32_t *ids, const float *rows, const int32_t *levels,
                       const int64_t *edge_off, const uint32_t *edge_ids, uint32_t entry_id, int max_level) {
    if (!h || n < 0 || (n > 0 && (!ids || !rows || !levels || !edge_off || !edge_ids)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::HNSWIndex &ix = h->ix;
    ix.free_dev();
    ix.slot_of.clear(); ix.deleted_ids.clear(); ix.levels_host.clear();
    ix.n = 0; ix.n_up = 0; ix.entry_slot = -1; ix.max_level = -1;
    ix.E0 = 2 * ix.m + 1; ix.EU = ix.m + 1;
    if (n == 0) return CM_OK;
    // slot map, per-layer degrees -> slot widths
    std::vector<long long> base((size_t)n + 1, 0);
    int64_t n_up = 0;
    for (int64_t i = 0; i < n; i++) {
        if (levels[i] < 0 || levels[i] >= cm::HNSW_MAX_LEVELS) return cm::fail(CM_ERR_INVALID_ARG, "level %d at slot %lld out of range", levels[i], (long long)i);
        base[(size_t)i + 1] = base[(size_t)i] + levels[i] + 1;
        ix.slot_of[ids[i]] = i;
        if (levels[i] > 0) n_up++;
    }
    for (int64_t i = 0; i < n; i++)
        for (int l = 0; l <= levels[i]; l++) {
            long long deg = edge_off[base[(size_t)i] + l + 1] - edge_off[base[(size_t)i] + l];
            if (l == 0) ix.E0 = std::max<int>(ix.E0, (int)deg + 1); else ix.EU = std::max<int>(ix.EU, (int)deg + 1);
        }
    auto ent = ix.slot_of.find(entry_id);
    if (ent == ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "entry point ID %u not in the graph", entry_id);
    CM_TRY(ix.reserve(n, n_up));
    std::vector<uint32_t> a0((size_t)n * ix.E0, 0), aU((size_t)std::max<int64_t>(n_up, 1) * cm::HNSW_MAX_LEVELS * ix.EU, 0);
    std::vector<int> d0((size_t)n, 0), dU((size_t)std::max<int64_t>(n_up, 1) * cm::HNSW_MAX_LEVELS, 0), upof((size_t)n, -1);
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        int my_u = -1;
        if (levels[i] > 0) { my_u = (int)u++; upof[(size_t)i] = my_u; }
        for (int l = 0; l <= levels[i]; l++) {
            long long e0 = edge_off[base[(size_t)i] + l], e1 = edge_off[base[(size_t)i] + l + 1];
            for (long long e = e0; e < e1; e++) {
                auto it = ix.slot_of.find(edge_ids[e]);
                if (it == ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "edge %lld points to unknown node ID %u", e, edge_ids[e]);
                if (l == 0) a0[(size_t)i * ix.E0 + (e - e0)] = (uint32_t)it->second;
                else aU[((size_t)my_u * cm::HNSW_MAX_LEVELS + l) * ix.EU + (e - e0)] = (uint32_t)it->second;
            }
            if (l == 0) d0[(size_t)i] = (int)(e1 - e0); else dU[(size_t)my_u * cm::HNSW_MAX_LEVELS + l] = (int)(e1 - e0);
        }
    }
    CM_CUDA(cudaMemcpy2D(ix.rows, (size_t)ix.ld * 4, rows, (size_t)ix.dim * 4, (size_t)ix.dim * 4, (size_t)n, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(ix.ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(ix.levels, levels, (size_t)n * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(ix.adj0, a0.data(), a0.size() * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(ix.deg0, d0.data(), d0.size() * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(ix.up_of, upof.data(), upof.size() * 4, cudaMemcpyHostToDevice));
    if (n_up > 0) {
        CM_CUDA(cudaMemcpy(ix.adjU, aU.data(), (size_t)n_up * cm::HNSW_MAX_LEVELS * ix.EU * 4, cudaMemcpyHostToDevice));
        CM_CUDA(cudaMemcpy(ix.degU, dU.data(), (size_t)n_up * cm::HNSW_MAX_LEVELS * 4, cudaMemcpyHostToDevice));
    }
    ix.levels_host.assign(levels, levels + n);
    ix.n = n; ix.n_up = n_up;
    ix.entry_slot = ent->second;
    ix.max_level = max_level;
    return CM_OK;
}

// n successive HNSWIndex.Add calls (hnsw_index.go:228-288) on the device: PreprocessInPlace (rows written back unless
// writeback == 0), then insertNode with the caller's level draws (randomLevel, hnsw_index.go:474-484, stays with the
// caller: the reference uses an unseeded global RNG).  IDs must be non-zero and new.
int cm_hnsw_add(cm_hnsw *h, const uint32_t *ids, float *rows, const int32_t *levels, int64_t n, int writeback) {
    if (!h || (n > 0 && (!ids || !rows || !levels))) return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (n <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cm::HNSWIndex &ix = h->ix;
    if (ix.E0 != 2 * ix.m + 1 || ix.EU != ix.m + 1)
        return cm::fail(CM_ERR_UNSUPPORTED, "this graph was loaded with wider adjacency lists than M allows: insertion is not defined on it");
    int64_t add_up = 0;
    for (int64_t i = 0; i < n; i++) {
        if (ids[i] == 0) return cm::fail(CM_ERR_INVALID_ARG, "node ID 0 is reserved (auto-assignment, hnsw_index.go:247-263)");
        if (ix.slot_of.count(ids[i])) return cm::fail(CM_ERR_INVALID_ARG, "node ID %u already present", ids[i]);
        if (levels[i] < 0 || levels[i] >= cm::HNSW_MAX_LEVELS) return cm::fail(CM_ERR_INVALID_ARG, "level %d out of range", levels[i]);
        if (levels[i] > 0) add_up++;
    }
    bool fma = cm::rounding_mode() == CM_ROUND_FMA;
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    int rc = ix.reserve(ix.n + n, ix.n_up + add_up);
    float *raw = nullptr;
    int *flags = nullptr, *status = nullptr;
    uint32_t *vis = nullptr, *touched = nullptr;
    cm::HCand *cands = nullptr, *tmp = nullptr;
    const int cand_cap = (int)std::min<int64_t>(ix.n + n + 1, (int64_t)16 * ix.efc + 4096);
    const int64_t vis_words = (ix.n + n + 31) / 32;
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&raw, (size_t)n * ix.dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&flags, (size_t)n * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&status, 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&vis, (size_t)vis_words * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&touched, (size_t)(ix.n + n + 1) * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&cands, (size_t)cand_cap * sizeof(cm::HCand), st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&tmp, (size_t)4 * (2 * ix.m + 2) * sizeof(cm::HCand), st);
    int64_t good = n;
    if (rc == CM_OK) {
        cudaMemcpyAsync(raw, rows, (size_t)n * ix.dim * 4, cudaMemcpyHostToDevice, st);
        rc = cm::launch_preprocess_rows(ix.metric, fma, raw, n, ix.dim, ix.dim, ix.rows + (size_t)ix.n * ix.ld, ix.ld, flags, st);
    }
    if (rc == CM_OK && ix.metric == CM_COSINE) {
        std::vector<int> hf((size_t)n);
        cudaMemcpyAsync(hf.data(), flags, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        for (int64_t i = 0; i < n; i++)
            if (hf[(size_t)i]) { good = i; break; }
        if (writeback && good > 0)
            cudaMemcpy2DAsync(rows, (size_t)ix.dim * 4, ix.rows + (size_t)ix.n * ix.ld, (size_t)ix.ld * 4, (size_t)ix.dim * 4,
                              (size_t)good, cudaMemcpyDeviceToHost, st);
    }
    if (rc == CM_OK && good > 0) {
        // node records: id, level, empty adjacency, upper block for nodes above layer 0
        std::vector<int> upof((size_t)good, -1);
        int64_t u = ix.n_up;
        for (int64_t i = 0; i < good; i++)
            if (levels[i] > 0) upof[(size_t)i] = (int)u++;
        cudaMemcpyAsync(ix.ids + ix.n, ids, (size_t)good * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(ix.levels + ix.n, levels, (size_t)good * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(ix.up_of + ix.n, upof.data(), (size_t)good * 4, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(ix.deg0 + ix.n, 0, (size_t)good * 4, st);
        cudaMemsetAsync(ix.deleted + ix.n, 0, (size_t)good, st);
        if (u > ix.n_up) cudaMemsetAsync(ix.degU + (size_t)ix.n_up * cm::HNSW_MAX_LEVELS, 0, (size_t)(u - ix.n_up) * cm::HNSW_MAX_LEVELS * 4, st);
        cudaMemsetAsync(vis, 0, (size_t)vis_words * 4, st);
        cudaMemsetAsync(status, 0, 4, st);
        size_t smem = cm::warp_scratch_bytes(ix.ld, ix.efc) + (size_t)ix.ld * 4;
        cm::GraphView G = ix.view();
#define CM_HNSW_INS(MM, F)                                                                                            \
    do {                                                                                                              \
        cm::set_dyn_smem((const void *)cm::hnsw_insert_kernel<MM, F>, smem);                                               \
        cm::hnsw_insert_kernel<MM, F><<<1, 32, smem, st>>>(G, (long long)ix.n, (int)good, ix.entry_slot, ix.max_level, ix.m, ix.efc, \
                                                           vis, cands, cand_cap, touched, tmp, status);               \
    } while (0)
        switch (ix.metric) {
        case CM_L2: if (fma) CM_HNSW_INS(CM_L2, true); else CM_HNSW_INS(CM_L2, false); break;
        case CM_L2SQ: if (fma) CM_HNSW_INS(CM_L2SQ, true); else CM_HNSW_INS(CM_L2SQ, false); break;
        default: if (fma) CM_HNSW_INS(CM_COSINE, true); else CM_HNSW_INS(CM_COSINE, false); break;
        }
#undef CM_HNSW_INS
        cm::count_launch();
        int hstatus = 0;
        cudaMemcpyAsync(&hstatus, status, 4, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cm::fail(CM_ERR_CUDA, "hnsw_add: %s", cudaGetErrorString(e));
        else if (hstatus) rc = cm::fail(CM_ERR_UNSUPPORTED, "hnsw_add: candidate heap overflow (efConstruction %d)", ix.efc);
        if (rc == CM_OK) {
            for (int64_t i = 0; i < good; i++) {
                ix.slot_of[ids[i]] = ix.n + i;
                ix.levels_host.push_back(levels[i]);
                if (levels[i] > ix.max_level) ix.max_level = levels[i];
            }
            if (ix.entry_slot < 0) ix.entry_slot = 0;
            ix.n += good;
            ix.n_up = u;
        }
    }
    cm::ws_free(raw, st); cm::ws_free(flags, st); cm::ws_free(status, st); cm::ws_free(vis, st); cm::ws_free(touched, st);
    cm::ws_free(cands, st); cm::ws_free(tmp, st);
    cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && good < n) return cm::fail(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector (row %lld of this Add batch)", (long long)good);
    return rc;
}

// the graph back to the host (HNSWIndex.WriteTo, tests): levels, and per (slot, layer) the neighbour IDs.
// edge_off must hold sum(levels + 1) + 1 entries, edge_ids at least cm_hnsw_edge_count() entries.
int64_t cm_hnsw_edge_count(const cm_hnsw *h) {
    if (!h || h->ix.n == 0) return 0;
    const cm::HNSWIndex &ix = h->ix;
    cudaSetDevice(ix.device);
    std::vector<int> d0((size_t)ix.n), dU((size_t)std::max<int64_t>(ix.n_up, 1) * cm::HNSW_MAX_LEVELS, 0);
    if (cudaMemcpy(d0.data(), ix.deg0, (size_t)ix.n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (ix.n_up > 0 && cudaMemcpy(dU.data(), ix.degU, (size_t)ix.n_up * cm::HNSW_MAX_LEVELS * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int64_t total = 0;
    for (int v : d0) total += v;
    int64_t u = 0;
    for (int64_t i = 0; i < ix.n; i++)
        if (ix.levels_host[(size_t)i] > 0) {
            for (int l = 1; l <= ix.levels_host[(size_t)i]; l++) total += dU[(size_t)u * cm::HNSW_MAX_LEVELS + l];
            u++;
        }
    return total;
}
int cm_hnsw_export_graph(const cm_hnsw *h, int32_t *levels, int64_t *edge_off, uint32_t *edge_ids, uint32_t *entry_id,
                         int *max_level) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    const cm::HNSWIndex &ix = h->ix;
    if (max_level) *max_level = ix.max_level;
    if (ix.n == 0) { if (entry_id) *entry_id = 0; if (edge_off) edge_off[0] = 0; return CM_OK; }
    CM_CUDA(cudaSetDevice(ix.device));
    std::vector<uint32_t> a0((size_t)ix.n * ix.E0), aU((size_t)std::max<int64_t>(ix.n_up, 1) * cm::HNSW_MAX_LEVELS * ix.EU), hid((size_t)ix.n);
    std::vector<int> d0((size_t)ix.n), dU((size_t)std::max<int64_t>(ix.n_up, 1) * cm::HNSW_MAX_LEVELS, 0);
    CM_CUDA(cudaMemcpy(a0.data(), ix.adj0, a0.size() * 4, cudaMemcpyDeviceToHost));
    CM_CUDA(cudaMemcpy(d0.data(), ix.deg0, d0.size() * 4, cudaMemcpyDeviceToHost));
    CM_CUDA(cudaMemcpy(hid.data(), ix.ids, hid.size() * 4, cudaMemcpyDeviceToHost));
    if (ix.n_up > 0) {
        CM_CUDA(cudaMemcpy(aU.data(), ix.adjU, (size_t)ix.n_up * cm::HNSW_MAX_LEVELS * ix.EU * 4, cudaMemcpyDeviceToHost));
        CM_CUDA(cudaMemcpy(dU.data(), ix.degU, (size_t)ix.n_up * cm::HNSW_MAX_LEVELS * 4, cudaMemcpyDeviceToHost));
    }
    int64_t pair = 0, e = 0, u = 0;
    if (edge_off) edge_off[0] = 0;
    for (int64_t i = 0; i < ix.n; i++) {
        int L = ix.levels_host[(size_t)i];
        if (levels) levels[i] = L;
        for (int l = 0; l <= L; l++) {
            int deg = l == 0 ? d0[(size_t)i] : dU[(size_t)u * cm::HNSW_MAX_LEVELS + l];
            const uint32_t *src = l == 0 ? &a0[(size_t)i * ix.E0] : &aU[((size_t)u * cm::HNSW_MAX_LEVELS + l) * ix.EU];
            for (int j = 0; j < deg; j++) edge_ids[e++] = hid[src[j]];
            edge_off[++pair] = e;
        }
        if (L > 0) u++;
    }
    if (entry_id) *entry_id = hid[(size_t)ix.entry_slot];
    return CM_OK;
}

// node IDs and STORED vectors by slot (insertion order): what a host mirror needs after cm_hnsw_load
int cm_hnsw_get_nodes(const cm_hnsw *h, int64_t first, int64_t n, uint32_t *ids_out, float *rows_out) {
    if (!h || first < 0 || n < 0 || first + n > h->ix.n) return cm::fail(CM_ERR_INVALID_ARG, "bad range");
    if (n == 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    if (ids_out) CM_CUDA(cudaMemcpy(ids_out, h->ix.ids + first, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (rows_out)
        CM_CUDA(cudaMemcpy2D(rows_out, (size_t)h->ix.dim * 4, h->ix.rows + (size_t)first * h->ix.ld, (size_t)h->ix.ld * 4,
                             (size_t)h->ix.dim * 4, (size_t)n, cudaMemcpyDeviceToHost));
    return CM_OK;
}

int cm_hnsw_remove(cm_hnsw *h, uint32_t id) {     // hnsw_index.go:300-330 soft delete
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    auto it = h->ix.slot_of.find(id);
    if (it == h->ix.slot_of.end()) return cm::fail(CM_ERR_NOT_FOUND, "node %u not found", id);
    if (h->ix.deleted_ids.count(id)) return cm::fail(CM_ERR_NOT_FOUND, "node %u already deleted", id);
    h->ix.deleted_ids.insert(id);
    uint8_t one = 1;
    CM_CUDA(cudaMemcpy(h->ix.deleted + it->second, &one, 1, cudaMemcpyHostToDevice));
    return CM_OK;
}

int cm_hnsw_search_device(cm_hnsw *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                          int64_t *out_counts_dev, int64_t *work_dev, void *stream) {
    if (!h || !p || (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::hnsw_search_device(h->ix, queries_dev, nq, p, out_stride, out_ids_dev, out_scores_dev, out_pos_dev,
                                  out_counts_dev, work_dev, (cudaStream_t)stream, false);
}

int cm_hnsw_search(cm_hnsw *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                   uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, int64_t *out_work) {
    if (!h || !p || (nq > 0 && (!queries || !out_ids || !out_scores || !out_counts)))
        return cm::fail(CM_ERR_INVALID_ARG, "null argument");
    if (dim != h->ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected %d, got %d", h->ix.dim, dim);
    if (nq <= 0) return CM_OK;
    CM_CUDA(cudaSetDevice(h->ix.device));
    cudaStream_t st;
    CM_TRY(cm::acquire_stream(&st));
    float *dq = nullptr, *dsc = nullptr;
    uint32_t *dids = nullptr;
    int64_t *dpos = nullptr, *dcnt = nullptr, *dwork = nullptr;
    size_t no = (size_t)nq * (size_t)(out_stride > 0 ? out_stride : 1);
    int rc = cm::ws_alloc((void **)&dq, (size_t)nq * dim * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dids, no * 4, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dsc, no * 4, st);
    if (rc == CM_OK && out_pos) rc = cm::ws_alloc((void **)&dpos, no * 8, st);
    if (rc == CM_OK) rc = cm::ws_alloc((void **)&dcnt, (size_t)nq * 8, st);
    if (rc == CM_OK && out_work) rc = cm::ws_alloc((void **)&dwork, (size_t)nq * 16, st);
    if (rc == CM_OK) {
        cudaMemcpyAsync(dq, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st);
        rc = cm::hnsw_search_device(h->ix, dq, nq, p, out_stride, dids, dsc, dpos, dcnt, dwork, st, true);
    }
    if (rc == CM_OK) {
        cudaMemcpyAsync(out_ids, dids, no * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_scores, dsc, no * 4, cudaMemcpyDeviceToHost, st);
        if (out_pos) cudaMemcpyAsync(out_pos, dpos, no * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, dcnt, (size_t)nq * 8, cudaMemcpyDeviceToHost, st);
        if (out_work) {
            if (h->ix.n == 0) memset(out_work, 0, (size_t)nq * 16);
            else cudaMemcpyAsync(out_work, dwork, (size_t)nq * 16, cudaMemcpyDeviceToHost, st);
        }
    }
    cm::ws_free(dq, st); cm::ws_free(dids, st); cm::ws_free(dsc, st); cm::ws_free(dpos, st); cm::ws_free(dcnt, st); cm::ws_free(dwork, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cm::release_stream(st);
    if (rc == CM_OK && e != cudaSuccess) return cm::fail(CM_ERR_CUDA, "hnsw_search: %s", cudaGetErrorString(e));
    if (rc == CM_OK)
        for (int64_t q = 0; q < nq; q++)
            if (out_counts[q] < 0) return cm::fail(CM_ERR_UNSUPPORTED, "candidate heap overflow on query %lld (efSearch too small a bound)", (long long)q);
    return rc;
}

// ---- HNSWIndex.Flush (hnsw_index.go:348-430), WriteTo / ReadFrom (hnsw_index.go:734-1096) ----------------------
// the whole graph on the host, in slot (insertion) order
struct HostGraph {
    std::vector<uint32_t> ids;
    std::vector<float> rows;
    std::vector<int32_t> levels;
    std::vector<int64_t> edge_off;
    std::vector<uint32_t> edge_ids;
    uint32_t entry = 0;
    int max_level = -1;
};
static int hnsw_to_host(cm_hnsw *h, HostGraph *g) {
    cm::HNSWIndex &ix = h->ix;
    const int64_t n = ix.n;
    g->ids.resize((size_t)n);
    g->rows.resize((size_t)n * ix.dim);
    g->levels.resize((size_t)n);
    int64_t pairs = 0;
    for (int64_t i = 0; i < n; i++) pairs += ix.levels_host[(size_t)i] + 1;
    g->edge_off.assign((size_t)pairs + 1, 0);
    g->edge_ids.resize((size_t)std::max<int64_t>(cm_hnsw_edge_count(h), 1));
    g->entry = 0;
    g->max_level = ix.max_level;
    if (n == 0) return CM_OK;
    CM_CUDA(cudaMemcpy(g->ids.data(), ix.ids, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CM_CUDA(cudaMemcpy2D(g->rows.data(), (size_t)ix.dim * 4, ix.rows, (size_t)ix.ld * 4, (size_t)ix.dim * 4, (size_t)n, cudaMemcpyDeviceToHost));
    return cm_hnsw_export_graph(h, g->levels.data(), g->edge_off.data(), g->edge_ids.data(), &g->entry, &g->max_level);
}

// Flush: (1) every live node drops its edges to deleted nodes; (2) a deleted entry point is replaced by a live node at
// maxLevel, else by a node of the highest level left (maxLevel follows), else the index is empty; (3) deleted nodes go;
// (4) the deleted set is cleared.  The reference walks a Go map in phase 2, so WHICH of several eligible nodes becomes
// the entry point is random there; here it is the earliest inserted one (lowest slot) -- one of the outcomes the
// reference can produce, chosen deterministically.
int cm_hnsw_flush(cm_hnsw *h) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    cm::HNSWIndex &ix = h->ix;
    if (ix.deleted_ids.empty()) return CM_OK;
    CM_CUDA(cudaSetDevice(ix.device));
    HostGraph g;
    CM_TRY(hnsw_to_host(h, &g));
    const std::unordered_set<uint32_t> dead = ix.deleted_ids;
    const int64_t n = (int64_t)g.ids.size();
    uint32_t entry = g.entry;
    int max_level = g.max_level;
    if (dead.count(entry)) {
        bool found = false;
        for (int64_t i = 0; i < n && !found; i++)
            if (!dead.count(g.ids[(size_t)i]) && g.levels[(size_t)i] == max_level) { entry = g.ids[(size_t)i]; found = true; }
        if (!found) {
            int best = -1;
            for (int64_t i = 0; i < n; i++)
                if (!dead.count(g.ids[(size_t)i]) && g.levels[(size_t)i] > best) { best = g.levels[(size_t)i]; entry = g.ids[(size_t)i]; }
            if (best >= 0) max_level = best;
            else { entry = 0; max_level = -1; }
        }
    }
    HostGraph out;
    out.edge_off.push_back(0);
    int64_t pair = 0;
    for (int64_t i = 0; i < n; i++) {
        const int L = g.levels[(size_t)i];
        if (!dead.count(g.ids[(size_t)i])) {
            out.ids.push_back(g.ids[(size_t)i]);
            out.levels.push_back(L);
            out.rows.insert(out.rows.end(), g.rows.begin() + (size_t)i * ix.dim, g.rows.begin() + (size_t)(i + 1) * ix.dim);
            for (int l = 0; l <= L; l++) {
                for (int64_t e = g.edge_off[(size_t)(pair + l)]; e < g.edge_off[(size_t)(pair + l + 1)]; e++)
                    if (!dead.count(g.edge_ids[(size_t)e])) out.edge_ids.push_back(g.edge_ids[(size_t)e]);
                out.edge_off.push_back((int64_t)out.edge_ids.size());
            }
        }
        pair += L + 1;
    }
    if (out.edge_ids.empty()) out.edge_ids.push_back(0);
    return cm_hnsw_load_graph(h, (int64_t)out.ids.size(), out.ids.data(), out.rows.data(), out.levels.data(), out.edge_off.data(),
                              out.edge_ids.data(), entry, max_level);
}

// WriteTo: Flush; header; M, efConstruction, efSearch, levelMult (float64, 1 / ln M); maxLevel (int32), entryPoint;
// node count; per node (ID, level, vector dimension, vector, edge layer count, per layer edge count + IDs); roaring
// blob.  The reference writes nodes in Go map order (random); here in insertion order -- ReadFrom accepts any order.
static int hnsw_save(cm_hnsw *h, cm::wire::Sink &s) {
    cm::HNSWIndex &ix = h->ix;
    CM_TRY(cm_hnsw_flush(h));
    HostGraph g;
    CM_TRY(hnsw_to_host(h, &g));
    CM_TRY(cm::wire::write_header(s, "HNSW", ix.dim, ix.metric));
    CM_WIRE_PUT(s.u32((uint32_t)ix.m), "M");
    CM_WIRE_PUT(s.u32((uint32_t)ix.efc), "efConstruction");
    CM_WIRE_PUT(s.u32((uint32_t)ix.efs), "efSearch");
    CM_WIRE_PUT(s.f64(1.0 / std::log((double)ix.m)), "levelMult");          // hnsw_index.go:206
    CM_WIRE_PUT(s.i32((int32_t)g.max_level), "maxLevel");
    CM_WIRE_PUT(s.u32(g.ids.empty() ? 0u : g.entry), "entryPoint");
    CM_WIRE_PUT(s.u32((uint32_t)g.ids.size()), "node count");
    int64_t pair = 0;
    std::vector<uint8_t> rec;
    for (size_t i = 0; i < g.ids.size(); i++) {
        const int L = g.levels[i];
        rec.clear();
        auto app = [&](const void *p, size_t nbytes) { rec.insert(rec.end(), (const uint8_t *)p, (const uint8_t *)p + nbytes); };
        const uint32_t id = g.ids[i], d = (uint32_t)ix.dim, layers = (uint32_t)(L + 1);
        const int32_t lvl = L;
        app(&id, 4); app(&lvl, 4); app(&d, 4);
        app(&g.rows[i * (size_t)ix.dim], (size_t)ix.dim * 4);
        app(&layers, 4);
        for (int l = 0; l <= L; l++) {
            const int64_t e0 = g.edge_off[(size_t)(pair + l)], e1 = g.edge_off[(size_t)(pair + l + 1)];
            const uint32_t cnt = (uint32_t)(e1 - e0);
            app(&cnt, 4);
            if (cnt) app(&g.edge_ids[(size_t)e0], (size_t)cnt * 4);
        }
        pair += L + 1;
        CM_WIRE_PUT(s.put(rec.data(), rec.size()), "node");
    }
    CM_WIRE_PUT(cm::wire::write_empty_bitmap(s), "bitmap");
    return CM_OK;
}

static int hnsw_load(cm_hnsw *h, cm::wire::Source &s) {
    cm::HNSWIndex &ix = h->ix;
    CM_TRY(cm::wire::read_header(s, "HNSW", ix.dim, ix.metric));
    uint32_t M = 0, efc = 0, efs = 0, entry = 0, count = 0;
    int32_t max_level = 0;
    double level_mult = 0.0;
    CM_WIRE_GET(s.u32(&M), "M");
    CM_WIRE_GET(s.u32(&efc), "efConstruction");
    CM_WIRE_GET(s.u32(&efs), "efSearch");
    if ((int64_t)M != ix.m) return cm::fail(CM_ERR_INVALID_ARG, "m parameter mismatch: index has m=%d, serialized data has m=%u", ix.m, M);
    if ((int64_t)efc != ix.efc) return cm::fail(CM_ERR_INVALID_ARG, "efConstruction mismatch: index has %d, serialized data has %u", ix.efc, efc);
    if ((int64_t)efs != ix.efs) return cm::fail(CM_ERR_INVALID_ARG, "efSearch mismatch: index has %d, serialized data has %u", ix.efs, efs);
    CM_WIRE_GET(s.f64(&level_mult), "levelMult");      // randomLevel stays with the caller (cm_hnsw_add takes the level draws)
    CM_WIRE_GET(s.i32(&max_level), "maxLevel");
    CM_WIRE_GET(s.u32(&entry), "entryPoint");
    CM_WIRE_GET(s.u32(&count), "node count");
    HostGraph g;
    g.ids.resize(count);
    g.levels.resize(count);
    g.rows.resize((size_t)count * ix.dim);
    g.edge_off.push_back(0);
    for (uint32_t i = 0; i < count; i++) {
        uint32_t vd = 0, layers = 0;
        CM_WIRE_GET(s.u32(&g.ids[i]), "node ID");
        CM_WIRE_GET(s.i32(&g.levels[i]), "node level");
        CM_WIRE_GET(s.u32(&vd), "vector dimension");
        if ((int64_t)vd != ix.dim) return cm::fail(CM_ERR_DIM_MISMATCH, "node %u has dimension %u, expected %d", g.ids[i], vd, ix.dim);
        CM_WIRE_GET(s.get(&g.rows[(size_t)i * ix.dim], (size_t)ix.dim * 4), "vector data");
        CM_WIRE_GET(s.u32(&layers), "edge layer count");
        if ((int64_t)layers != (int64_t)g.levels[i] + 1)
            return cm::fail(CM_ERR_INVALID_ARG, "node %u has %u edge layers at level %d", g.ids[i], layers, g.levels[i]);
        for (uint32_t l = 0; l < layers; l++) {
            uint32_t cnt = 0;
            CM_WIRE_GET(s.u32(&cnt), "edge count");
            const size_t at = g.edge_ids.size();
            g.edge_ids.resize(at + cnt);
            CM_WIRE_GET(cnt == 0 || s.get(&g.edge_ids[at], (size_t)cnt * 4), "edge IDs");
            g.edge_off.push_back((int64_t)g.edge_ids.size());
        }
    }
    std::vector<uint32_t> dead;
    CM_TRY(cm::wire::read_bitmap(s, &dead));
    if (g.edge_ids.empty()) g.edge_ids.push_back(0);
    CM_TRY(cm_hnsw_load_graph(h, (int64_t)count, g.ids.data(), g.rows.data(), g.levels.data(), g.edge_off.data(), g.edge_ids.data(),
                              entry, (int)max_level));
    for (uint32_t id : dead)
        if (cm_hnsw_remove(h, id) != CM_OK) ix.deleted_ids.insert(id);
    return CM_OK;
}

int cm_hnsw_save(cm_hnsw *h, uint8_t *buf, int64_t cap, int64_t *bytes) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_buffer([&](cm::wire::Sink &s) { return hnsw_save(h, s); }, buf, cap, bytes);
}
int cm_hnsw_load(cm_hnsw *h, const uint8_t *buf, int64_t len, int64_t *consumed) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_buffer([&](cm::wire::Source &s) { return hnsw_load(h, s); }, buf, len, consumed);
}
int cm_hnsw_save_file(cm_hnsw *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::save_to_file([&](cm::wire::Sink &s) { return hnsw_save(h, s); }, path);
}
int cm_hnsw_load_file(cm_hnsw *h, const char *path) {
    if (!h) return cm::fail(CM_ERR_INVALID_ARG, "null handle");
    CM_CUDA(cudaSetDevice(h->ix.device));
    return cm::wire::load_from_file([&](cm::wire::Source &s) { return hnsw_load(h, s); }, path);
}

}  // extern "C"
