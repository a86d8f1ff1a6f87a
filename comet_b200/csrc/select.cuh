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

static inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace cm
