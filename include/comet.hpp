// comet.hpp -- C++17 host mirror of comet's Go Index / Search builder API over the C ABI of
// libcomet_b200.so (include/comet_b200.h).  Same names, argument meaning and error texts as the
// reference so that code (and tests) written against
//     idx, _ := comet.NewFlatIndex(384, comet.Cosine)
//     idx.Add(*comet.NewVectorNodeWithID(7, vec))
//     results, err := idx.NewSearch().WithQuery(q).WithK(10).Execute()
// read the same here:
//     auto idx = comet::NewFlatIndex(384, comet::Cosine);
//     idx->Add(comet::NewVectorNodeWithID(7, vec));
//     auto results = idx->NewSearch()->WithQuery({q}).WithK(10).Execute();
// Go's (value, error) returns become exceptions of type comet::Error carrying the reference's message.
//
// What lives here is exactly what stays host code in the Go package: validation, the host
// mirror of the nodes (Node.Vector(), WithNode lookups), aggregation (aggregation.go:107-255), limit /
// autocut (limiter.go:12-118), rerank.  Every distance, scan, selection and graph traversal happens
// in the CUDA library: one cm_*_search call per Execute() carrying all of the builder's queries.
// There is no CPU fallback: without the library + a B200 every index operation throws.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <functional>
#include <istream>
#include <iterator>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "comet_b200.h"

namespace comet {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct ZeroVectorError : Error {                 // distance.go:9-12 ErrZeroVector
    ZeroVectorError() : Error(CM_ERR_ZERO_VECTOR, "cannot normalize zero vector") {}
};
inline void check(int rc) {
    if (rc == CM_OK) return;
    if (rc == CM_ERR_ZERO_VECTOR) throw ZeroVectorError();
    throw Error(rc, cm_last_error());
}

// distance.go:21-38
enum DistanceKind { Euclidean = CM_L2, L2Squared = CM_L2SQ, Cosine = CM_COSINE };
inline const char *DistanceKindName(DistanceKind k) { return k == Euclidean ? "l2" : (k == L2Squared ? "l2_squared" : "cosine"); }
// index.go:7-30
using VectorIndexKind = std::string;
// aggregation.go:20-40
enum ScoreAggregationKind { DefaultAggregation = 0, SumAggregation, MaxAggregation, MeanAggregation };

// node.go:30-108.  Copies share the vector storage like Go slices do: Add normalises the caller's data
// in place for cosine (flat_index.go:182).
class VectorNode {
public:
    VectorNode() : id_(0), vec_(std::make_shared<std::vector<float>>()) {}
    VectorNode(uint32_t id, std::shared_ptr<std::vector<float>> v) : id_(id), vec_(std::move(v)) {}
    uint32_t ID() const { return id_; }
    std::vector<float> &Vector() { return *vec_; }
    const std::vector<float> &Vector() const { return *vec_; }

private:
    uint32_t id_;
    std::shared_ptr<std::vector<float>> vec_;
};
inline std::atomic<uint32_t> &nodeIDCounter() { static std::atomic<uint32_t> c{0}; return c; }
inline VectorNode NewVectorNode(std::vector<float> v) {                       // node.go:55
    return VectorNode(nodeIDCounter().fetch_add(1) + 1, std::make_shared<std::vector<float>>(std::move(v)));
}
inline VectorNode NewVectorNodeWithID(uint32_t id, std::vector<float> v) {   // node.go:87
    return VectorNode(id, std::make_shared<std::vector<float>>(std::move(v)));
}

// index_search.go:84-100
struct VectorResult {
    VectorNode Node;
    float Score = 0.0f;
    uint32_t GetId() const { return Node.ID(); }
    float GetScore() const { return Score; }
};
// index_search.go:50-60
using Reranker = std::function<std::vector<VectorResult>(std::vector<VectorResult>)>;

// ---- limiter.go -----------------------------------------------------------------------------
inline long sanitizeK(long k, long maxResults) { return (k <= 0 || k > maxResults) ? maxResults : k; }   // :12-17
template <class T>
std::vector<T> LimitResults(std::vector<T> results, long k) {                                             // :28
    results.resize((size_t)sanitizeK(k, (long)results.size()));
    return results;
}
inline long Autocut(const std::vector<float> &y, int cutOff) {                                            // :70-118
    const long n = (long)y.size();
    if (n <= 1) return n;
    std::vector<float> diff((size_t)n);
    const float step = 1.0f / ((float)n - 1.0f);
    for (long i = 0; i < n; i++) {
        float x = 0.0f + (float)i * step;
        float yn = (y[(size_t)i] - y[0]) / (y[(size_t)n - 1] - y[0]);
        diff[(size_t)i] = yn - x;
    }
    int extrema = 0;
    for (long i = 1; i < n; i++) {
        bool peak;
        if (i == n - 1) peak = diff[(size_t)i] > diff[(size_t)i - 1] && (i < 2 || diff[(size_t)i] > diff[(size_t)i - 2]);
        else peak = diff[(size_t)i] > diff[(size_t)i - 1] && diff[(size_t)i] > diff[(size_t)i + 1];
        if (peak && ++extrema >= cutOff) return i;
    }
    return n;
}
template <class T>
std::vector<T> AutocutResults(std::vector<T> results, int cutoff) {                                       // :52
    if (cutoff == -1 || results.empty()) return results;
    std::vector<float> s(results.size());
    for (size_t i = 0; i < results.size(); i++) s[i] = results[i].GetScore();
    results.resize((size_t)Autocut(s, cutoff));
    return results;
}

// ---- aggregation.go:107-255: combine by node ID, sort ascending.  Go iterates a map and sorts with an
// unstable sort, so equal scores have no defined order there; here: first occurrence, stable sort. ----
inline std::vector<VectorResult> Aggregate(ScoreAggregationKind kind, const std::vector<VectorResult> &results) {
    if (results.empty()) return results;
    std::vector<uint32_t> order;
    std::unordered_map<uint32_t, std::vector<float>> scores;
    std::unordered_map<uint32_t, VectorNode> nodes;
    for (const auto &r : results) {
        auto it = scores.find(r.Node.ID());
        if (it == scores.end()) { order.push_back(r.Node.ID()); scores[r.Node.ID()] = {r.Score}; }
        else it->second.push_back(r.Score);
        nodes[r.Node.ID()] = r.Node;
    }
    std::vector<VectorResult> out;
    out.reserve(order.size());
    for (uint32_t id : order) {
        const auto &s = scores[id];
        float v;
        if (kind == MaxAggregation) {
            v = s[0];
            for (size_t i = 1; i < s.size(); i++) if (s[i] > v) v = s[i];
        } else {
            float sum = 0.0f;
            for (float x : s) sum += x;
            v = kind == MeanAggregation ? sum / (float)s.size() : sum;
        }
        out.push_back(VectorResult{nodes[id], v});
    }
    std::stable_sort(out.begin(), out.end(), [](const VectorResult &a, const VectorResult &b) { return a.Score < b.Score; });
    return out;
}

class VectorIndex;

// index_search.go:141-279 -- single use, not thread safe, like the reference's builders
class VectorSearch {
public:
    explicit VectorSearch(VectorIndex *ix) : index_(ix) {}
    VectorSearch &WithQuery(std::vector<std::vector<float>> queries) { queries_ = std::move(queries); return *this; }
    VectorSearch &WithNode(std::vector<uint32_t> nodeIDs) { nodeIDs_ = std::move(nodeIDs); return *this; }
    VectorSearch &WithK(int k) { k_ = k; return *this; }
    VectorSearch &WithNProbes(int n) { nprobes_ = n; return *this; }
    VectorSearch &WithEfSearch(int ef) { efSearch_ = ef; return *this; }
    VectorSearch &WithThreshold(float t) { threshold_ = t; return *this; }
    VectorSearch &WithScoreAggregation(ScoreAggregationKind k) { aggregation_ = k; return *this; }
    VectorSearch &WithCutoff(int c) { cutoff_ = c; return *this; }
    VectorSearch &WithDocumentIDs(std::vector<uint32_t> ids) { documentIDs_ = std::move(ids); return *this; }
    VectorSearch &WithReranker(Reranker r) { reranker_ = std::move(r); return *this; }
    inline std::vector<VectorResult> Execute();                       // flat_index_search.go:109-165 and siblings
    inline std::vector<std::vector<VectorResult>> ExecuteBatch();     // additive: per-query lists, no aggregation

private:
    inline std::vector<std::vector<float>> allQueries();
    VectorIndex *index_;
    std::vector<std::vector<float>> queries_;
    std::vector<uint32_t> nodeIDs_, documentIDs_;
    int k_ = 10, nprobes_ = -1, efSearch_ = 0, cutoff_ = -1;          // flat_index.go:305: k=10, cutoff=-1
    float threshold_ = 0.0f;
    ScoreAggregationKind aggregation_ = DefaultAggregation;
    Reranker reranker_;
    friend class VectorIndex;
};

// index.go:32-63, io.WriterTo / io.ReaderFrom included: WriteTo / ReadFrom move the reference's byte formats
// (FLAT / IVFX / PQIX / IVPQ / HNSW) through cm_*_save / cm_*_load
class VectorIndex {
public:
    virtual ~VectorIndex() = default;
    virtual void Train(const std::vector<VectorNode> &vectors) = 0;
    virtual void Add(VectorNode vector) = 0;
    virtual void Remove(const VectorNode &vector) = 0;
    virtual void Flush() = 0;
    std::unique_ptr<VectorSearch> NewSearch() { auto s = std::make_unique<VectorSearch>(this); s->nprobes_ = defaultNProbes(); return s; }
    int Dimensions() const { return dim_; }
    comet::DistanceKind DistanceKind() const { return kind_; }
    virtual VectorIndexKind Kind() const = 0;
    virtual bool Trained() const = 0;
    size_t Len() const { return nodes_.size(); }
    // WriteTo (e.g. flat_index.go:366-470): flushes, then writes the reference's stream; returns the bytes written
    int64_t WriteTo(std::ostream &w) {
        Flush();
        int64_t need = 0;
        saveBytes(nullptr, 0, &need);
        std::vector<uint8_t> buf((size_t)std::max<int64_t>(need, 1));
        check(saveBytes(buf.data(), need, &need));
        w.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)need);
        if (!w) throw Error(CM_ERR_INVALID_ARG, "failed to write index data");
        return need;
    }
    // ReadFrom (e.g. flat_index.go:488-614) on a pre-constructed index with matching parameters: replaces its state
    int64_t ReadFrom(std::istream &r) {
        std::vector<uint8_t> buf((std::istreambuf_iterator<char>(r)), std::istreambuf_iterator<char>());
        int64_t used = 0;
        check(loadBytes(buf.data(), (int64_t)buf.size(), &used));
        nodes_.clear(); by_id_.clear(); deleted_.clear();
        rebuildMirror();
        return used;
    }

protected:
    friend class VectorSearch;
    virtual int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) = 0;
    virtual int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) = 0;
    virtual void rebuildMirror() = 0;             // host nodes (ID + stored vector where the index keeps one) from the device state
    void mirrorFrom(const std::vector<uint32_t> &ids, const std::vector<float> *rows) {
        for (size_t i = 0; i < ids.size(); i++)
            remember(NewVectorNodeWithID(ids[i], rows ? std::vector<float>(rows->begin() + (long)(i * (size_t)dim_), rows->begin() + (long)((i + 1) * (size_t)dim_))
                                                      : std::vector<float>()));
    }
    VectorIndex(int dim, comet::DistanceKind kind) : dim_(dim), kind_(kind) {}
    // nq searchSingleQuery calls in ONE device batch: flat row-major queries -> ids/scores/counts [nq][stride]
    virtual void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride,
                             std::vector<uint32_t> &ids, std::vector<float> &scores, std::vector<int64_t> &counts) = 0;
    virtual int64_t resultBound(const VectorSearch &s) const { return (int64_t)nodes_.size(); }
    virtual int defaultNProbes() const { return -1; }
    void checkDim(const VectorNode &v) const {
        if ((int)v.Vector().size() != dim_)
            throw Error(CM_ERR_DIM_MISMATCH, "vector dimension mismatch: expected " + std::to_string(dim_) + ", got " +
                                                 std::to_string(v.Vector().size()));
    }
    void remember(const VectorNode &v) { by_id_[v.ID()] = nodes_.size(); nodes_.push_back(v); }
    const VectorNode *nodeByID(uint32_t id) const {
        auto it = by_id_.find(id);
        return it == by_id_.end() ? nullptr : &nodes_[it->second];
    }
    void forget(const std::unordered_set<uint32_t> &ids) {          // Flush: drop soft-deleted nodes, keep order
        std::vector<VectorNode> keep;
        for (auto &n : nodes_) if (!ids.count(n.ID())) keep.push_back(n);
        nodes_.swap(keep);
        by_id_.clear();
        for (size_t i = 0; i < nodes_.size(); i++) by_id_[nodes_[i].ID()] = i;
    }
    int dim_;
    comet::DistanceKind kind_;
    std::vector<VectorNode> nodes_;                                  // host mirror, insertion order
    std::unordered_map<uint32_t, size_t> by_id_;
    std::unordered_set<uint32_t> deleted_;
};

inline std::vector<std::vector<float>> VectorSearch::allQueries() {
    if (queries_.empty() && nodeIDs_.empty()) throw Error(CM_ERR_INVALID_ARG, "must specify either queries or node IDs");   // :112
    std::vector<std::vector<float>> all = queries_;
    for (uint32_t id : nodeIDs_) {                                   // lookupNodeVectors, flat_index_search.go:171-196
        const VectorNode *n = index_->nodeByID(id);
        if (!n || index_->deleted_.count(id)) throw Error(CM_ERR_NOT_FOUND, "node ID " + std::to_string(id) + " not found in index");
        all.push_back(n->Vector());
    }
    for (const auto &q : all)
        if ((int)q.size() != index_->dim_)
            throw Error(CM_ERR_DIM_MISMATCH, "query dimension mismatch: expected " + std::to_string(index_->dim_) + ", got " +
                                                 std::to_string(q.size()));                                                  // :227
    return all;
}

inline std::vector<std::vector<VectorResult>> VectorSearch::ExecuteBatch() {
    auto all = allQueries();
    const int64_t nq = (int64_t)all.size();
    std::vector<float> flat;
    flat.reserve((size_t)nq * index_->dim_);
    for (const auto &q : all) flat.insert(flat.end(), q.begin(), q.end());
    cm_search_params p{};
    p.k = k_; p.threshold = threshold_; p.nprobes = nprobes_; p.ef_search = efSearch_; p.path = CM_PATH_AUTO;
    p.filter_ids = documentIDs_.empty() ? nullptr : documentIDs_.data();
    p.nfilter = (int64_t)documentIDs_.size();
    int64_t bound = index_->resultBound(*this);
    int64_t stride = std::max<int64_t>(1, sanitizeK(k_, bound));
    std::vector<uint32_t> ids((size_t)(nq * stride));
    std::vector<float> scores((size_t)(nq * stride));
    std::vector<int64_t> counts((size_t)nq, 0);
    index_->searchBatch(flat, nq, p, stride, ids, scores, counts);
    std::vector<std::vector<VectorResult>> out((size_t)nq);
    for (int64_t q = 0; q < nq; q++)
        for (int64_t j = 0; j < counts[(size_t)q]; j++) {
            uint32_t id = ids[(size_t)(q * stride + j)];
            const VectorNode *n = index_->nodeByID(id);
            out[(size_t)q].push_back(VectorResult{n ? *n : NewVectorNodeWithID(id, {}), scores[(size_t)(q * stride + j)]});
        }
    return out;
}

inline std::vector<VectorResult> VectorSearch::Execute() {
    auto per_query = ExecuteBatch();
    std::vector<VectorResult> all;
    for (auto &l : per_query) all.insert(all.end(), l.begin(), l.end());
    ScoreAggregationKind kind = aggregation_ == DefaultAggregation ? SumAggregation : aggregation_;    // :116-119
    auto results = Aggregate(kind, all);
    results = LimitResults(std::move(results), k_);
    results = AutocutResults(std::move(results), cutoff_);
    if (reranker_) results = reranker_(std::move(results));
    return results;
}

// ---- FlatIndex (flat_index.go) -----------------------------------------------------------------
class FlatIndex : public VectorIndex {
public:
    FlatIndex(int dim, comet::DistanceKind kind) : VectorIndex(dim, kind) { check(cm_flat_create(dim, (int)kind, &h_)); }
    ~FlatIndex() override { cm_flat_destroy(h_); }
    void Train(const std::vector<VectorNode> &) override {}                                       // flat_index.go:145: no-op
    void Add(VectorNode v) override {
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_flat_add(h_, &id, v.Vector().data(), 1, 1));                                     // normalises v in place (cosine)
        remember(v);
    }
    void AddBatch(std::vector<VectorNode> &vs) {                                                  // n successive Add()s, one upload
        std::vector<uint32_t> ids;
        std::vector<float> rows;
        for (auto &v : vs) { checkDim(v); ids.push_back(v.ID()); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        int rc = cm_flat_add(h_, ids.data(), rows.data(), (int64_t)vs.size(), 1);
        int64_t added = cm_flat_size(h_) - (int64_t)nodes_.size();
        for (int64_t i = 0; i < added; i++) {
            std::copy(rows.begin() + i * dim_, rows.begin() + (i + 1) * dim_, vs[(size_t)i].Vector().begin());
            remember(vs[(size_t)i]);
        }
        check(rc);
    }
    void Remove(const VectorNode &v) override { check(cm_flat_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_flat_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "flat"; }
    bool Trained() const override { return true; }

protected:
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_flat_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), nullptr, counts.data()));
    }
    int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) override { return cm_flat_save(h_, buf, cap, bytes); }
    int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) override { return cm_flat_load(h_, buf, len, used); }
    void rebuildMirror() override {
        const int64_t n = cm_flat_size(h_);
        std::vector<uint32_t> ids((size_t)n);
        std::vector<float> rows((size_t)n * dim_);
        std::vector<int64_t> pos((size_t)n);
        for (int64_t i = 0; i < n; i++) pos[(size_t)i] = i;
        if (n > 0) { check(cm_flat_get_ids(h_, 0, n, ids.data())); check(cm_flat_get_rows(h_, pos.data(), n, rows.data())); }
        mirrorFrom(ids, &rows);
    }
    cm_flat *h_ = nullptr;
};
inline std::unique_ptr<FlatIndex> NewFlatIndex(int dim, DistanceKind kind) { return std::make_unique<FlatIndex>(dim, kind); }   // flat_index.go:118

// The same index row-sharded over the GPUs of the box (cm_flat_sharded_*: one process drives every device; NVLink peer
// mappings carry queries and per-shard top-K lists; one merge kernel).  Same API, same answers; WriteTo / ReadFrom are
// not offered on this variant (save each segment from a single-GPU index).
class ShardedFlatIndex : public VectorIndex {
public:
    ShardedFlatIndex(int dim, comet::DistanceKind kind, const std::vector<int> &devices, int64_t rowsPerShard) : VectorIndex(dim, kind) {
        check(cm_flat_sharded_create(dim, (int)kind, devices.data(), (int)devices.size(), rowsPerShard, &h_));
    }
    ~ShardedFlatIndex() override { cm_flat_sharded_destroy(h_); }
    void Train(const std::vector<VectorNode> &) override {}
    void Add(VectorNode v) override {
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_flat_sharded_add(h_, &id, v.Vector().data(), 1, 1));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_flat_sharded_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_flat_sharded_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "flat"; }
    bool Trained() const override { return true; }
    int Shards() const { return cm_flat_sharded_shards(h_); }

protected:
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_flat_sharded_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), counts.data()));
    }
    int saveBytes(uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    int loadBytes(const uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    void rebuildMirror() override {}
    cm_flat_sharded *h_ = nullptr;
};
inline std::unique_ptr<ShardedFlatIndex> NewShardedFlatIndex(int dim, DistanceKind kind, const std::vector<int> &devices, int64_t rowsPerShard) {
    return std::make_unique<ShardedFlatIndex>(dim, kind, devices, rowsPerShard);
}

// ---- IVFIndex (ivf_index.go) -------------------------------------------------------------------
class IVFIndex : public VectorIndex {
public:
    IVFIndex(int dim, int nlist, comet::DistanceKind kind) : VectorIndex(dim, kind), nlist_(nlist) { check(cm_ivf_create(dim, nlist, (int)kind, &h_)); }
    ~IVFIndex() override { cm_ivf_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {                                 // ivf_index.go:205-246
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_ivf_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetCentroids(const std::vector<float> &c) { check(cm_ivf_set_centroids(h_, c.data())); }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_ivf_add(h_, &id, v.Vector().data(), 1, 1, nullptr));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_ivf_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_ivf_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "ivf"; }
    bool Trained() const override { return cm_ivf_trained(h_) != 0; }

protected:
    int defaultNProbes() const override { return cm_ivf_default_nprobes(h_); }
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_ivf_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), nullptr, counts.data()));
    }
    int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) override { return cm_ivf_save(h_, buf, cap, bytes); }
    int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) override { return cm_ivf_load(h_, buf, len, used); }
    void rebuildMirror() override {
        const int64_t n = cm_ivf_size(h_);
        std::vector<uint32_t> ids((size_t)n);
        std::vector<float> rows((size_t)n * dim_);
        std::vector<int64_t> pos((size_t)n);
        for (int64_t i = 0; i < n; i++) pos[(size_t)i] = i;
        if (n > 0) { check(cm_ivf_get_ids(h_, 0, n, ids.data())); check(cm_ivf_get_rows(h_, pos.data(), n, rows.data())); }
        mirrorFrom(ids, &rows);
    }
    int nlist_;
    cm_ivf *h_ = nullptr;
};
inline std::unique_ptr<IVFIndex> NewIVFIndex(int dim, int nlist, DistanceKind kind) { return std::make_unique<IVFIndex>(dim, nlist, kind); }   // ivf_index.go:147

// ---- PQIndex (pq_index.go) ---------------------------------------------------------------------
class PQIndex : public VectorIndex {
public:
    PQIndex(int dim, comet::DistanceKind kind, int M, int Nbits) : VectorIndex(dim, kind) { check(cm_pq_create(dim, (int)kind, M, Nbits, &h_)); }
    ~PQIndex() override { cm_pq_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {                                 // pq_index.go:193-247
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_pq_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetCodebooks(const std::vector<float> &cb) { check(cm_pq_set_codebooks(h_, cb.data())); }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_pq_add(h_, &id, v.Vector().data(), 1, 1));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_pq_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_pq_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "pq"; }
    bool Trained() const override { return cm_pq_trained(h_) != 0; }

protected:
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_pq_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), nullptr, counts.data()));
    }
    int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) override { return cm_pq_save(h_, buf, cap, bytes); }
    int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) override { return cm_pq_load(h_, buf, len, used); }
    void rebuildMirror() override {                      // PQIndex.ReadFrom keeps IDs only: NewVectorNodeWithID(id, nil), pq_index.go:815
        const int64_t n = cm_pq_size(h_);
        std::vector<uint32_t> ids((size_t)n);
        if (n > 0) check(cm_pq_get_ids(h_, 0, n, ids.data()));
        mirrorFrom(ids, nullptr);
    }
    cm_pq *h_ = nullptr;
};
inline std::unique_ptr<PQIndex> NewPQIndex(int dim, DistanceKind kind, int M, int Nbits) { return std::make_unique<PQIndex>(dim, kind, M, Nbits); }   // pq_index.go:135

// ---- IVFPQIndex (ivfpq_index.go) ---------------------------------------------------------------
class IVFPQIndex : public VectorIndex {
public:
    IVFPQIndex(int dim, comet::DistanceKind kind, int nlist, int m, int nbits) : VectorIndex(dim, kind) {
        check(cm_ivfpq_create(dim, (int)kind, nlist, m, nbits, &h_));
    }
    ~IVFPQIndex() override { cm_ivfpq_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {                                 // ivfpq_index.go:180-259
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_ivfpq_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetTrained(const std::vector<float> &centroids, const std::vector<float> &codebooks) {
        check(cm_ivfpq_set_trained(h_, centroids.data(), codebooks.data()));
    }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_ivfpq_add(h_, &id, v.Vector().data(), 1, 1, nullptr));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_ivfpq_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_ivfpq_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "ivfpq"; }
    bool Trained() const override { return cm_ivfpq_trained(h_) != 0; }

protected:
    int defaultNProbes() const override { return cm_ivfpq_default_nprobes(h_); }
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_ivfpq_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), nullptr, counts.data()));
    }
    int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) override { return cm_ivfpq_save(h_, buf, cap, bytes); }
    int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) override { return cm_ivfpq_load(h_, buf, len, used); }
    void rebuildMirror() override {
        const int64_t n = cm_ivfpq_size(h_);
        std::vector<uint32_t> ids((size_t)n);
        if (n > 0) check(cm_ivfpq_get_ids(h_, 0, n, ids.data()));
        mirrorFrom(ids, nullptr);
    }
    cm_ivfpq *h_ = nullptr;
};
inline std::unique_ptr<IVFPQIndex> NewIVFPQIndex(int dim, DistanceKind kind, int nlist, int m, int nbits) {   // ivfpq_index.go:114
    return std::make_unique<IVFPQIndex>(dim, kind, nlist, m, nbits);
}

// ---- multi-GPU layouts of the trained indexes (one process drives every device) ---------------
// Same VectorIndex surface as the single-device classes; WriteTo / ReadFrom are not offered on these (segments per shard
// are the host's business, storage.go:545-626).
class ShardedIVFIndex : public VectorIndex {             // lists spread over the devices, global candidate numbering
public:
    ShardedIVFIndex(int dim, int nlist, comet::DistanceKind kind, const std::vector<int> &devices) : VectorIndex(dim, kind) {
        check(cm_ivf_sharded_create(dim, nlist, (int)kind, devices.data(), (int)devices.size(), &h_));
    }
    ~ShardedIVFIndex() override { cm_ivf_sharded_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_ivf_sharded_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetCentroids(const std::vector<float> &c) { check(cm_ivf_sharded_set_centroids(h_, c.data())); }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_ivf_sharded_add(h_, &id, v.Vector().data(), 1, 1, nullptr));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_ivf_sharded_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_ivf_sharded_flush(h_)); forget(deleted_); deleted_.clear(); }
    void Rebalance() { check(cm_ivf_sharded_rebalance(h_)); }
    VectorIndexKind Kind() const override { return "ivf"; }
    bool Trained() const override { return cm_ivf_sharded_trained(h_) != 0; }
    int Shards() const { return cm_ivf_sharded_shards(h_); }

protected:
    int defaultNProbes() const override { return cm_ivf_sharded_default_nprobes(h_); }
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_ivf_sharded_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), counts.data()));
    }
    int saveBytes(uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    int loadBytes(const uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    void rebuildMirror() override {}
    cm_ivf_sharded *h_ = nullptr;
};
inline std::unique_ptr<ShardedIVFIndex> NewShardedIVFIndex(int dim, int nlist, DistanceKind kind, const std::vector<int> &devices) {
    return std::make_unique<ShardedIVFIndex>(dim, nlist, kind, devices);
}

class ShardedPQIndex : public VectorIndex {              // rows spread over the devices like the flat row shards
public:
    ShardedPQIndex(int dim, comet::DistanceKind kind, int M, int Nbits, const std::vector<int> &devices, int64_t rowsPerShard)
        : VectorIndex(dim, kind) {
        check(cm_pq_sharded_create(dim, (int)kind, M, Nbits, devices.data(), (int)devices.size(), rowsPerShard, &h_));
    }
    ~ShardedPQIndex() override { cm_pq_sharded_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_pq_sharded_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetCodebooks(const std::vector<float> &cb) { check(cm_pq_sharded_set_codebooks(h_, cb.data())); }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding vectors");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_pq_sharded_add(h_, &id, v.Vector().data(), 1, 1));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_pq_sharded_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_pq_sharded_flush(h_)); forget(deleted_); deleted_.clear(); }
    VectorIndexKind Kind() const override { return "pq"; }
    bool Trained() const override { return cm_pq_sharded_trained(h_) != 0; }
    int Shards() const { return cm_pq_sharded_shards(h_); }

protected:
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_pq_sharded_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), counts.data()));
    }
    int saveBytes(uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    int loadBytes(const uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    void rebuildMirror() override {}
    cm_pq_sharded *h_ = nullptr;
};
inline std::unique_ptr<ShardedPQIndex> NewShardedPQIndex(int dim, DistanceKind kind, int M, int Nbits, const std::vector<int> &devices,
                                                         int64_t rowsPerShard) {
    return std::make_unique<ShardedPQIndex>(dim, kind, M, Nbits, devices, rowsPerShard);
}

class ShardedIVFPQIndex : public VectorIndex {           // lists spread over the devices, codes travel with their list
public:
    ShardedIVFPQIndex(int dim, comet::DistanceKind kind, int nlist, int m, int nbits, const std::vector<int> &devices)
        : VectorIndex(dim, kind) {
        check(cm_ivfpq_sharded_create(dim, (int)kind, nlist, m, nbits, devices.data(), (int)devices.size(), &h_));
    }
    ~ShardedIVFPQIndex() override { cm_ivfpq_sharded_destroy(h_); }
    void Train(const std::vector<VectorNode> &vectors) override {
        std::vector<float> rows;
        for (const auto &v : vectors) { checkDim(v); rows.insert(rows.end(), v.Vector().begin(), v.Vector().end()); }
        check(cm_ivfpq_sharded_train(h_, rows.data(), (int64_t)vectors.size()));
    }
    void SetTrained(const std::vector<float> &centroids, const std::vector<float> &codebooks) {
        check(cm_ivfpq_sharded_set_trained(h_, centroids.data(), codebooks.data()));
    }
    void Add(VectorNode v) override {
        if (!Trained()) throw Error(CM_ERR_NOT_TRAINED, "index must be trained before adding");
        checkDim(v);
        uint32_t id = v.ID();
        check(cm_ivfpq_sharded_add(h_, &id, v.Vector().data(), 1, 1, nullptr));
        remember(v);
    }
    void Remove(const VectorNode &v) override { check(cm_ivfpq_sharded_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    void Flush() override { check(cm_ivfpq_sharded_flush(h_)); forget(deleted_); deleted_.clear(); }
    void Rebalance() { check(cm_ivfpq_sharded_rebalance(h_)); }
    VectorIndexKind Kind() const override { return "ivfpq"; }
    bool Trained() const override { return cm_ivfpq_sharded_trained(h_) != 0; }
    int Shards() const { return cm_ivfpq_sharded_shards(h_); }

protected:
    int defaultNProbes() const override { return cm_ivfpq_sharded_default_nprobes(h_); }
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_ivfpq_sharded_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), counts.data()));
    }
    int saveBytes(uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    int loadBytes(const uint8_t *, int64_t, int64_t *) override { return CM_ERR_UNSUPPORTED; }
    void rebuildMirror() override {}
    cm_ivfpq_sharded *h_ = nullptr;
};
inline std::unique_ptr<ShardedIVFPQIndex> NewShardedIVFPQIndex(int dim, DistanceKind kind, int nlist, int m, int nbits,
                                                               const std::vector<int> &devices) {
    return std::make_unique<ShardedIVFPQIndex>(dim, kind, nlist, m, nbits, devices);
}

// ---- HNSWIndex (hnsw_index.go): insertion and search on the device; LoadGraph restores a serialised graph ----
class HNSWIndex : public VectorIndex {
public:
    HNSWIndex(int dim, comet::DistanceKind kind, int m, int efConstruction, int efSearch) : VectorIndex(dim, kind) {
        check(cm_hnsw_create(dim, (int)kind, m, efConstruction, efSearch, &h_));
        m_ = m > 0 ? m : 16;
    }
    ~HNSWIndex() override { cm_hnsw_destroy(h_); }
    void Train(const std::vector<VectorNode> &) override {}                                       // hnsw_index.go:215: no-op
    // hnsw_index.go:228-288.  The level is drawn like randomLevel (:474-484: geometric, p = 1/M, capped at 16) from this
    // index's own generator (the reference uses the unseeded global math/rand/v2; seed with SetLevelSeed for replays).
    void Add(VectorNode v) override { AddWithLevel(v, randomLevel()); }
    void AddWithLevel(VectorNode v, int level) {
        checkDim(v);
        uint32_t id = v.ID();
        int32_t lv = level;
        check(cm_hnsw_add(h_, &id, v.Vector().data(), &lv, 1, 1));
        remember(v);
    }
    void SetLevelSeed(uint64_t seed) { rng_state_ = seed ? seed : 0x9E3779B97F4A7C15ull; }
    // nodes in insertion order with their STORED vectors; edges per (node, layer) as neighbour IDs
    void LoadGraph(const std::vector<VectorNode> &nodes, const std::vector<int32_t> &levels, const std::vector<int64_t> &edge_off,
                   const std::vector<uint32_t> &edge_ids, uint32_t entry_id, int max_level) {
        std::vector<uint32_t> ids;
        std::vector<float> rows;
        for (const auto &n : nodes) { checkDim(n); ids.push_back(n.ID()); rows.insert(rows.end(), n.Vector().begin(), n.Vector().end()); }
        check(cm_hnsw_load_graph(h_, (int64_t)nodes.size(), ids.data(), rows.data(), levels.data(), edge_off.data(), edge_ids.data(),
                                 entry_id, max_level));
        nodes_.clear(); by_id_.clear(); deleted_.clear();
        for (const auto &n : nodes) remember(n);
    }
    void Remove(const VectorNode &v) override { check(cm_hnsw_remove(h_, v.ID())); deleted_.insert(v.ID()); }
    // hnsw_index.go:348-430: edges to deleted nodes go, a deleted entry point is replaced, deleted nodes are freed
    void Flush() override { check(cm_hnsw_flush(h_)); forget(deleted_); deleted_.clear(); }
    int MaxLevel() const { return cm_hnsw_max_level(h_); }
    VectorIndexKind Kind() const override { return "hnsw"; }
    bool Trained() const override { return true; }

protected:
    int64_t resultBound(const VectorSearch &) const override { return std::max<int64_t>(1, (int64_t)nodes_.size()); }
    void searchBatch(const std::vector<float> &flat, int64_t nq, const cm_search_params &p, int64_t stride, std::vector<uint32_t> &ids,
                     std::vector<float> &scores, std::vector<int64_t> &counts) override {
        check(cm_hnsw_search(h_, flat.data(), nq, dim_, &p, stride, ids.data(), scores.data(), nullptr, counts.data(), nullptr));
    }
    int saveBytes(uint8_t *buf, int64_t cap, int64_t *bytes) override { return cm_hnsw_save(h_, buf, cap, bytes); }
    int loadBytes(const uint8_t *buf, int64_t len, int64_t *used) override { return cm_hnsw_load(h_, buf, len, used); }
    void rebuildMirror() override {
        const int64_t n = cm_hnsw_size(h_);
        std::vector<uint32_t> ids((size_t)n);
        std::vector<float> rows((size_t)n * dim_);
        if (n > 0) check(cm_hnsw_get_nodes(h_, 0, n, ids.data(), rows.data()));
        mirrorFrom(ids, &rows);
    }
    int randomLevel() {
        int m = m_ > 0 ? m_ : 16, level = 0;
        const double p = 1.0 / (double)m;
        for (;;) {
            rng_state_ ^= rng_state_ << 13; rng_state_ ^= rng_state_ >> 7; rng_state_ ^= rng_state_ << 17;   // xorshift64
            double u = (double)(rng_state_ >> 11) * (1.0 / 9007199254740992.0);
            if (!(u < p) || level >= 16) break;
            level++;
        }
        return level;
    }
    int m_ = 16;
    uint64_t rng_state_ = 0x9E3779B97F4A7C15ull;
    cm_hnsw *h_ = nullptr;
};
inline std::unique_ptr<HNSWIndex> NewHNSWIndex(int dim, DistanceKind kind, int m, int efConstruction, int efSearch) {   // hnsw_index.go:172
    return std::make_unique<HNSWIndex>(dim, kind, m, efConstruction, efSearch);
}

}  // namespace comet
