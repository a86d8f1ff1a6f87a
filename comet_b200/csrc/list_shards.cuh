// list_shards.cuh -- what the list-sharded IVF and IVFPQ indexes share (cm_ivf_sharded_*, cm_ivfpq_sharded_*): one host
// process, W shards on the GPUs of one box, every shard a complete index with ALL centroids but only the inverted lists it
// owns.  A search runs on every shard concurrently (one stream + one enqueueing thread per shard); every shard returns
// its top-K by (score, candidate number) together with each winner's number in the reference's append loop over ALL
// probed lists (ivf_index_search.go:252-268, ivfpq_index_search.go:263-322) -- computed from the replicated global list
// lengths -- and devices[0] merges by (score, that number).  Queries are read and result lists written straight through
// NVLink peer mappings; without peer access the same bytes move by cudaMemcpyPeerAsync.
#pragma once

#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "shard_worker.cuh"

namespace cm {

struct ListShards {
    int dim = 0, nlist = 0, metric = 0;
    std::vector<int> dev;
    std::vector<int> owner;                 // [nlist] shard that holds the list (l mod W until a rebalance)
    std::vector<long long> glob_len;        // [nlist] vectors in the list, soft-deleted ones included (len(idx.lists[l]))
    std::vector<long long *> glob_len_dev;  // per shard: device copy of glob_len
    bool len_dirty = true;
    std::vector<cudaStream_t> st;
    std::vector<cudaEvent_t> done;
    std::vector<char> direct;               // shard r works on devices[0]'s query / gather buffers through peer mappings
    cudaEvent_t start = nullptr;
    struct Buf { float *q = nullptr; uint32_t *ids = nullptr, *gno = nullptr; float *sc = nullptr; int64_t *cnt = nullptr; int64_t cap_q = 0, cap_o = 0, cap_n = 0; };
    std::vector<Buf> buf;
    uint32_t *g_ids = nullptr, *g_gno = nullptr, *m_ids = nullptr;
    float *g_sc = nullptr, *m_sc = nullptr, *q_lead = nullptr;
    int64_t *g_cnt = nullptr, *m_cnt = nullptr;
    int64_t cap_g = 0, cap_m = 0, cap_ql = 0, cap_nq = 0;
    std::mutex search_mu;
    std::vector<std::unique_ptr<ShardWorker>> workers;
    std::vector<int64_t> last_scanned;      // per shard, last host search

    // search of shard r: (r, queries, nq, params with k = K, out_stride K, ids, scores, counts, stream, global lengths on
    // r's device, global candidate numbers out)
    using ShardSearch = std::function<int(int, const float *, int64_t, const cm_search_params *, int64_t, uint32_t *, float *,
                                          int64_t *, cudaStream_t, const long long *, uint32_t *)>;

    int W() const { return (int)dev.size(); }
    int init(int dim_, int nlist_, int metric_, const int *devices, int n_devices);
    void destroy();                         // streams, events, buffers, workers (the shards belong to the caller)
    int64_t effective_k(const cm_search_params *p) const;      // sanitizeK against the most candidates nprobes lists hold
    std::vector<int> greedy_plan() const;   // longest list first onto the lightest shard
    int search_impl(const ShardSearch &fn, const float *q_lead_dev, int64_t nq, const cm_search_params *p, int64_t K, cudaStream_t lead);
    int search_device(const ShardSearch &fn, const float *queries_dev, int64_t nq, const cm_search_params *p, int64_t out_stride,
                      uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev, cudaStream_t lead);
    int search_host(const ShardSearch &fn, const float *queries, int64_t nq, const cm_search_params *p, int64_t out_stride,
                    uint32_t *out_ids, float *out_scores, int64_t *out_counts);
};

}  // namespace cm
