// select.cuh -- CTA-level "K smallest 64-bit keys" on a shared-memory buffer.
//
// Every top-K on the search path is a selection of the K smallest keys make_key(score, position)
// (common.cuh), so the result is the reference's full sort + truncate (flat_index_search.go:277-291)
// with ties resolved by scan position, independent of how rows were distributed over threads/CTAs.
//
// The buffer protocol: producers append keys below the current bound `tau` at `count++`; when fewer
// than `headroom` slots are free the CTA calls compact(): pad with KEY_INF, bitonic-sort the C slots,
// keep the first K, and tighten tau to the K-th key.  Expected appends for a random stream of n keys
// are K*ln(n/C) after the first fill, so compactions are rare.
#pragma once

#include "common.cuh"

namespace cm {

// Barrier policy: whole CTA (__syncthreads) or a named barrier over a subset of warps.
struct CtaBarrier {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct NamedBarrier {
    int id, nthreads;
    __device__ __forceinline__ void sync() const { named_bar_sync(id, nthreads); }
};

// In-place ascending bitonic sort of C (power of two) keys by `nthr` threads (tid in [0,nthr)).
template <typename Barrier>
__device__ __forceinline__ void bitonic_sort_smem(uint64_t *keys, int C, int tid, int nthr, const Barrier &bar) {
    for (int k = 2; k <= C; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (C >> 1); i += nthr) {
                int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                int hi = lo | j;
                bool up = (lo & k) == 0;
                uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            bar.sync();
        }
    }
}

// Keep the K smallest of keys[0..*count) (count <= C).  After return keys[0..min(count,K)) are sorted
// ascending, *count = min(count, K), *tau = K-th key (or KEY_INF while fewer than K are held).
// Must be called by all `nthr` threads; starts and ends with a barrier.
template <typename Barrier>
__device__ __forceinline__ void compact_topk(uint64_t *keys, int C, int K, int *count, uint64_t *tau, int tid,
                                             int nthr, const Barrier &bar) {
    bar.sync();
    int n = *count;
    if (n > C) n = C;
    int Cs = 64;                      // sort only the smallest power of two that holds the n keys
    while (Cs < n) Cs <<= 1;
    if (Cs > C) Cs = C;
    for (int i = n + tid; i < Cs; i += nthr) keys[i] = KEY_INF;
    bar.sync();
    bitonic_sort_smem(keys, Cs, tid, nthr, bar);
    if (tid == 0) {
        int m = n < K ? n : K;
        *count = m;
        *tau = (m >= K) ? keys[K - 1] : KEY_INF;
    }
    bar.sync();
}

// Cheaper compaction for streams that compact often (short lists: the first round of every CTA fills
// the buffer before any bound exists).  Instead of sorting, find a score bound v by a radix select on
// the ordered score bits (windows of 8 bits from the highest bit in which the staged keys differ; a
// window's bin is accepted whole as soon as that keeps no more than `keep_max` keys) and drop the keys
// above it in place.  Afterwards keys[0..*count) is an UNSORTED superset of the K smallest
// (K <= *count <= keep_max), *tau = the first key above the bound.  Returns false -- buffer untouched
// -- when ties make even a fully resolved bound keep more than keep_max keys (the caller sorts then).
// All `nthr` threads (a whole CTA, nthr <= 1024) must call it; hist: 256 ints, red: 4 words of scratch.
__device__ __forceinline__ bool compact_select(uint64_t *keys, int C, int K, int keep_max, int *count, uint64_t *tau,
                                               int tid, int nthr, int *hist, uint32_t *red) {
    __syncthreads();
    int n = *count;
    if (n > C) n = C;
    if (n <= keep_max) return true;                 // nothing has to go (uniform: *count is stable here)
    const int lane = tid & 31;
    // range of the score words
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < n; i += nthr) { uint32_t h = (uint32_t)(keys[i] >> 32); lo = min(lo, h); hi = max(hi, h); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (tid == 0) { red[0] = 0xFFFFFFFFu; red[1] = 0u; }
    __syncthreads();
    if (lane == 0) { atomicMin(&red[0], lo); atomicMax(&red[1], hi); }
    __syncthreads();
    const uint32_t base = red[0], range = red[1] - red[0];
    int bits_left = range == 0 ? 0 : 32 - __clz(range);
    uint32_t prefix = 0;            // decided high bits of (score - base)
    int rank = K, below = 0;        // rank of the wanted key among the undecided keys; keys under the prefix
    bool done = false;
    while (bits_left > 0 && !done) {
        const int width = min(8, bits_left), shift = bits_left - width;
        for (int i = tid; i < 256; i += nthr) hist[i] = 0;
        __syncthreads();
        const uint32_t mask = (shift + width) >= 32 ? 0u : (0xFFFFFFFFu << (shift + width));
        for (int i = tid; i < n; i += nthr) {
            uint32_t v = (uint32_t)(keys[i] >> 32) - base;
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & ((1u << width) - 1u)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            int h[8], sum = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) { h[u] = hist[lane * 8 + u]; sum += h[u]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int before = inc - sum;
            if (rank > before && rank <= inc) {
                int rr = rank - before, bb = 0, acc = before;
                for (; bb < 7; bb++) {
                    if (rr <= h[bb]) break;
                    rr -= h[bb];
                    acc += h[bb];
                }
                red[2] = (uint32_t)(lane * 8 + bb);     // selected bin
                red[3] = (uint32_t)acc;                 // undecided keys in lower bins
                hist[0] = h[bb];                        // keys in the selected bin (hist is rebuilt next pass)
            }
        }
        __syncthreads();
        const int bin = (int)red[2], under = (int)red[3], inbin = hist[0];
        __syncthreads();
        prefix |= (uint32_t)bin << shift;
        if (below + under + inbin <= keep_max) {
            // the whole bin fits: bound = its largest value
            prefix |= shift > 0 ? ((1u << shift) - 1u) : 0u;
            below += under + inbin;
            done = true;
        } else {
            below += under;
            rank -= under;
            bits_left = shift;
        }
    }
    if (!done) {
        // fully resolved (or all scores equal): every key with score == the K-th score stays
        if (bits_left == 0 && range != 0) {
            // below = keys strictly under the K-th score; ties at it decide
        }
        // count the keys at or under the bound
        if (tid == 0) red[2] = 0u;
        __syncthreads();
        int c = 0;
        for (int i = tid; i < n; i += nthr) c += ((uint32_t)(keys[i] >> 32) - base) <= prefix ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(&red[2], (uint32_t)c);
        __syncthreads();
        const int kept = (int)red[2];
        __syncthreads();
        if (kept > keep_max) return false;
    }
    // ---- drop the keys above the bound, in place: chunk by chunk, reads of a chunk before its writes ----
    const uint32_t bound = base + prefix;
    if (tid == 0) red[2] = 0u;
    __syncthreads();
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i0 = 0; i0 < n; i0 += nthr) {
        const int i = i0 + tid;
        const uint64_t key = i < n ? keys[i] : 0ull;
        const bool keep = i < n && (uint32_t)(key >> 32) <= bound;
        __syncthreads();
        const uint32_t b = __ballot_sync(0xffffffffu, keep);
        if (b != 0u) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(&red[2], (uint32_t)__popc(b));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (keep) keys[slot0 + __popc(b & lt_mask)] = key;
        }
        __syncthreads();
    }
    if (tid == 0) {
        *count = (int)red[2];
        *tau = bound == 0xFFFFFFFFu ? KEY_INF : ((uint64_t)(bound + 1u) << 32);
    }
    __syncthreads();
    return true;
}

static inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace cm
