// host_api_test.cpp -- the reference's own search tests replayed against the C++ host mirror
// (include/comet.hpp).  `--cpu` runs only the host-side logic (aggregation, limiter, autocut, builder
// validation); without it every index test runs on the GPU through libcomet_b200.so.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <iostream>

#include "comet.hpp"

using namespace comet;
static int failures = 0;
#define EXPECT(cond)                                                              \
    do {                                                                          \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)
template <class F>
static bool throws(F f, const char *needle = nullptr) {
    try { f(); } catch (const Error &e) { return !needle || std::strstr(e.what(), needle) != nullptr; }
    return false;
}

static void test_limiter_and_aggregation() {
    // limiter_test.go:7-73 sanitizeK table
    EXPECT(sanitizeK(0, 5) == 5); EXPECT(sanitizeK(-1, 5) == 5); EXPECT(sanitizeK(3, 5) == 3); EXPECT(sanitizeK(10, 5) == 5);
    // aggregation_test.go:7-115: sum 0.1+0.15+0.05 == float32(0.3), max 0.5, mean == 0.2
    auto n1 = NewVectorNodeWithID(1, {1, 0}), n2 = NewVectorNodeWithID(2, {0, 1});
    std::vector<VectorResult> r = {{n1, 0.1f}, {n2, 0.5f}, {n1, 0.15f}, {n1, 0.05f}};
    auto s = Aggregate(SumAggregation, r);
    EXPECT(s.size() == 2 && s[0].GetId() == 1 && s[0].Score == 0.1f + 0.15f + 0.05f && s[1].Score == 0.5f);
    auto m = Aggregate(MaxAggregation, r);
    EXPECT(m[0].GetId() == 1 && m[0].Score == 0.15f);
    auto a = Aggregate(MeanAggregation, r);
    EXPECT(a[0].Score == (0.1f + 0.15f + 0.05f) / 3.0f);
    // limiter_test.go:185-256 Autocut: one clear jump -> cut before it
    EXPECT(Autocut({0.1f, 0.11f, 0.12f, 0.9f, 0.91f}, 1) == 3);
    EXPECT(Autocut({0.5f}, 1) == 1);
    EXPECT(LimitResults(s, 1).size() == 1 && LimitResults(s, 0).size() == 2);
    // node.go:55: auto IDs count up from 1
    uint32_t a0 = NewVectorNode({1}).ID(), a1 = NewVectorNode({1}).ID();
    EXPECT(a1 == a0 + 1 && a0 >= 1);
}

static void test_flat() {
    // flat_index_search_test.go:10-48
    auto idx = NewFlatIndex(3, Euclidean);
    idx->Add(NewVectorNodeWithID(1, {1, 0, 0}));
    idx->Add(NewVectorNodeWithID(2, {0, 1, 0}));
    idx->Add(NewVectorNodeWithID(3, {0, 0, 1}));
    idx->Add(NewVectorNodeWithID(4, {1, 1, 0}));
    auto res = idx->NewSearch()->WithQuery({{1, 0, 0}}).WithK(2).Execute();
    EXPECT(res.size() == 2 && res[0].GetId() == 1 && res[0].Score == 0.0f && res[0].Node.Vector()[0] == 1.0f);
    // flat_index_search_test.go:229-278: multi-query union, default aggregation = sum, dedup by id
    auto mq = idx->NewSearch()->WithQuery({{1, 0, 0}, {0, 1, 0}}).WithK(3).Execute();
    // per-query top 3: {1:0, 4:1, 2:sqrt2} and {2:0, 4:1, 1:sqrt2} -> sums 1: sqrt2, 2: sqrt2, 4: 2
    EXPECT(mq.size() == 3 && mq[2].GetId() == 4 && mq[2].Score == 2.0f && mq[0].Score == mq[1].Score);
    // :51-86 threshold, :348-389 k bounds
    EXPECT(idx->NewSearch()->WithQuery({{1, 0, 0}}).WithK(10).WithThreshold(1.2f).Execute().size() == 2);
    EXPECT(idx->NewSearch()->WithQuery({{1, 0, 0}}).WithK(0).Execute().size() == 4);
    EXPECT(idx->NewSearch()->WithQuery({{1, 0, 0}}).WithK(-5).Execute().size() == 4);
    // WithNode, WithDocumentIDs, errors
    auto nn = idx->NewSearch()->WithNode({2}).WithK(1).Execute();
    EXPECT(nn.size() == 1 && nn[0].GetId() == 2);
    auto df = idx->NewSearch()->WithQuery({{1, 0, 0}}).WithDocumentIDs({2, 3}).WithK(10).Execute();
    EXPECT(df.size() == 2 && (df[0].GetId() == 2 || df[0].GetId() == 3));
    EXPECT(throws([&] { idx->NewSearch()->Execute(); }, "must specify either queries or node IDs"));
    EXPECT(throws([&] { idx->NewSearch()->WithQuery({{1, 0}}).Execute(); }, "query dimension mismatch: expected 3, got 2"));
    EXPECT(throws([&] { idx->NewSearch()->WithNode({99}).Execute(); }, "node ID 99 not found in index"));
    EXPECT(throws([&] { idx->Add(NewVectorNodeWithID(9, {1, 2})); }, "vector dimension mismatch: expected 3, got 2"));
    EXPECT(throws([&] { NewFlatIndex(0, Euclidean); }, "dimension must be positive"));
    // soft delete + flush (flat_index_test.go:343-434)
    idx->Remove(NewVectorNodeWithID(1, {}));
    EXPECT(idx->NewSearch()->WithQuery({{1, 0, 0}}).WithK(10).Execute().size() == 3);
    idx->Flush();
    EXPECT(idx->Len() == 3);
    // cosine: Add normalises the caller's storage in place; zero vector -> ErrZeroVector
    auto cidx = NewFlatIndex(2, Cosine);
    auto v = NewVectorNodeWithID(1, {3, 4});
    cidx->Add(v);
    EXPECT(std::fabs(v.Vector()[0] - 0.6f) < 1e-6f && std::fabs(v.Vector()[1] - 0.8f) < 1e-6f);
    bool zero = false;
    try { cidx->Add(NewVectorNodeWithID(2, {0, 0})); } catch (const ZeroVectorError &) { zero = true; }
    EXPECT(zero);
    EXPECT(idx->Kind() == "flat" && idx->Dimensions() == 3 && idx->DistanceKind() == Euclidean && idx->Trained());
}

static void test_trained_indexes() {
    // two obvious clusters: IVF / PQ / IVFPQ must be trainable through the builder API and find exact matches first
    std::vector<VectorNode> nodes;
    for (int i = 0; i < 200; i++) {
        float b = i % 2 ? 10.0f : 0.0f;
        nodes.push_back(NewVectorNodeWithID((uint32_t)(i + 1), {b + (i % 7) * 0.1f, b + (i % 5) * 0.1f, b + (i % 3) * 0.1f, b + (i % 11) * 0.1f}));
    }
    auto ivf = NewIVFIndex(4, 2, Euclidean);
    EXPECT(!ivf->Trained());
    EXPECT(throws([&] { ivf->Add(nodes[0]); }, "index must be trained before adding vectors"));
    ivf->Train(nodes);
    EXPECT(ivf->Trained());
    for (auto &n : nodes) ivf->Add(n);
    auto r = ivf->NewSearch()->WithQuery({nodes[10].Vector()}).WithK(1).WithNProbes(1).Execute();
    EXPECT(r.size() == 1 && r[0].Score == 0.0f);
    auto pq = NewPQIndex(4, Euclidean, 2, 4);
    pq->Train(nodes);
    for (auto &n : nodes) pq->Add(n);
    auto rp = pq->NewSearch()->WithQuery({nodes[11].Vector()}).WithK(5).Execute();
    EXPECT(rp.size() == 5 && rp[0].Node.Vector().size() == 4);
    auto ivfpq = NewIVFPQIndex(4, Euclidean, 2, 2, 4);
    ivfpq->Train(nodes);
    for (auto &n : nodes) ivfpq->Add(n);
    auto ri = ivfpq->NewSearch()->WithQuery({nodes[12].Vector()}).WithK(5).WithNProbes(2).Execute();
    EXPECT(ri.size() == 5);
    EXPECT(throws([&] { NewPQIndex(4, Euclidean, 2, 256); }));          // pq_index.go:152 (the reference's broken benchmark argument)
    EXPECT(throws([&] { NewIVFIndex(4, 50, Euclidean)->Train(nodes); NewIVFIndex(4, 500, Euclidean)->Train(nodes); }, "need at least 500"));
}

static void test_hnsw() {
    // hnsw_index_search_test.go:123-203 style: after Add, the exact match ranks first
    auto idx = NewHNSWIndex(4, Euclidean, 4, 20, 20);
    idx->SetLevelSeed(7);
    std::vector<VectorNode> nodes;
    for (int i = 0; i < 20; i++) nodes.push_back(NewVectorNodeWithID((uint32_t)(i + 1), {(float)i, (float)(i % 3), (float)(i % 5), 1.0f}));
    for (auto &n : nodes) idx->Add(n);
    EXPECT(idx->Len() == 20 && idx->Kind() == "hnsw");
    auto r = idx->NewSearch()->WithQuery({nodes[2].Vector()}).WithK(3).Execute();
    EXPECT(r.size() == 3 && r[0].GetId() == 3 && r[0].Score == 0.0f);
    EXPECT(throws([&] { idx->NewSearch()->WithQuery({{1, 2}}).Execute(); }, "query dimension mismatch: expected 4, got 2"));
}


static void test_serialisation_flush_and_shards() {
    // flat_index_test.go / ivf_index_test.go ... "WriteTo and ReadFrom" round trips: a restored index answers like the saved one
    std::vector<VectorNode> nodes;
    for (int i = 0; i < 300; i++)
        nodes.push_back(NewVectorNodeWithID((uint32_t)(i + 1), {(float)(i % 17), (float)(i % 5) * 0.5f, (float)(i / 30), 1.0f + (float)(i % 3)}));
    auto same = [](const std::vector<VectorResult> &a, const std::vector<VectorResult> &b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); i++)
            if (a[i].GetId() != b[i].GetId() || a[i].Score != b[i].Score) return false;
        return true;
    };
    const std::vector<float> q = {3.0f, 1.0f, 4.0f, 2.0f};
    {
        auto a = NewFlatIndex(4, Cosine);
        for (auto &n : nodes) a->Add(n);
        a->Remove(nodes[5]);
        std::stringstream ss;
        int64_t wrote = a->WriteTo(ss);
        EXPECT(wrote == (int64_t)ss.str().size() && a->Len() == 299);          // WriteTo flushes (flat_index.go:367-370)
        EXPECT(ss.str().compare(0, 4, "FLAT") == 0);
        auto b = NewFlatIndex(4, Cosine);
        EXPECT(b->ReadFrom(ss) == wrote && b->Len() == 299);
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(7).Execute(), b->NewSearch()->WithQuery({q}).WithK(7).Execute()));
        EXPECT(b->NewSearch()->WithQuery({q}).WithK(1).Execute()[0].Node.Vector().size() == 4);    // the mirror holds the stored vectors
        std::stringstream again(ss.str());
        EXPECT(throws([&] { NewFlatIndex(5, Cosine)->ReadFrom(again); }, "dimension mismatch: index has dim=5, serialized data has dim=4"));
    }
    {
        auto a = NewIVFPQIndex(4, Euclidean, 3, 2, 4);
        a->Train(nodes);
        for (auto &n : nodes) a->Add(n);
        std::stringstream ss;
        a->WriteTo(ss);
        EXPECT(ss.str().compare(0, 4, "IVPQ") == 0);
        auto b = NewIVFPQIndex(4, Euclidean, 3, 2, 4);
        b->ReadFrom(ss);
        EXPECT(b->Trained() && b->Len() == 300);
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(9).WithNProbes(2).Execute(), b->NewSearch()->WithQuery({q}).WithK(9).WithNProbes(2).Execute()));
    }
    {
        // HNSWIndex.Flush (hnsw_index.go:348-430): removed nodes disappear for good, the entry point is replaced
        auto a = NewHNSWIndex(4, Euclidean, 4, 20, 20);
        a->SetLevelSeed(11);
        for (auto &n : nodes) a->Add(n);
        a->Remove(nodes[0]);                                    // the first node inserted is the entry point
        a->Remove(nodes[1]);
        a->Flush();
        EXPECT(a->Len() == 298);
        auto r = a->NewSearch()->WithQuery({nodes[40].Vector()}).WithK(3).Execute();
        EXPECT(!r.empty() && r[0].GetId() != 1 && r[0].GetId() != 2);
        EXPECT(throws([&] { a->Remove(nodes[0]); }));
        std::stringstream ss;
        a->WriteTo(ss);
        auto b = NewHNSWIndex(4, Euclidean, 4, 20, 20);
        b->ReadFrom(ss);
        EXPECT(b->Len() == 298 && b->MaxLevel() == a->MaxLevel());
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(5).Execute(), b->NewSearch()->WithQuery({q}).WithK(5).Execute()));
    }
    {
        // the row-sharded flat index answers like the single one (three shards on the visible device)
        auto a = NewFlatIndex(4, Euclidean);
        auto s = NewShardedFlatIndex(4, Euclidean, {0, 0, 0}, 120);
        for (auto &n : nodes) { a->Add(n); s->Add(n); }
        EXPECT(s->Shards() == 3 && s->Len() == 300);
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(25).Execute(), s->NewSearch()->WithQuery({q}).WithK(25).Execute()));
        a->Remove(nodes[7]); s->Remove(nodes[7]);
        a->Flush(); s->Flush();
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(25).Execute(), s->NewSearch()->WithQuery({q}).WithK(25).Execute()));
        EXPECT(throws([&] { s->Add(NewVectorNodeWithID(999, {1, 2, 3})); }, "vector dimension mismatch: expected 4, got 3"));
    }
    {
        // list shards (IVF, IVFPQ) and PQ row shards: same trained state, same answers as the single index, ties included
        auto a = NewIVFIndex(4, 5, Euclidean);
        a->Train(nodes);
        auto s = NewShardedIVFIndex(4, 5, Euclidean, {0, 0});
        EXPECT(throws([&] { s->Add(nodes[0]); }, "index must be trained before adding vectors"));
        s->Train(nodes);
        for (auto &n : nodes) { a->Add(n); s->Add(n); }
        EXPECT(s->Shards() == 2 && s->Len() == 300 && s->Trained());
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(40).WithNProbes(3).Execute(), s->NewSearch()->WithQuery({q}).WithK(40).WithNProbes(3).Execute()));
        s->Rebalance();
        a->Remove(nodes[9]); s->Remove(nodes[9]);
        a->Flush(); s->Flush();
        EXPECT(same(a->NewSearch()->WithQuery({q}).WithK(40).Execute(), s->NewSearch()->WithQuery({q}).WithK(40).Execute()));

        auto pa = NewPQIndex(4, Euclidean, 2, 4);
        pa->Train(nodes);
        auto ps = NewShardedPQIndex(4, Euclidean, 2, 4, {0, 0, 0}, 110);
        ps->Train(nodes);
        for (auto &n : nodes) { pa->Add(n); ps->Add(n); }
        EXPECT(ps->Shards() == 3 && ps->Len() == 300);
        EXPECT(same(pa->NewSearch()->WithQuery({q}).WithK(30).Execute(), ps->NewSearch()->WithQuery({q}).WithK(30).Execute()));

        auto ia = NewIVFPQIndex(4, Euclidean, 3, 2, 4);
        ia->Train(nodes);
        auto is = NewShardedIVFPQIndex(4, Euclidean, 3, 2, 4, {0, 0});
        is->Train(nodes);
        for (auto &n : nodes) { ia->Add(n); is->Add(n); }
        EXPECT(is->Shards() == 2 && is->Len() == 300);
        EXPECT(same(ia->NewSearch()->WithQuery({q}).WithK(30).WithNProbes(2).Execute(), is->NewSearch()->WithQuery({q}).WithK(30).WithNProbes(2).Execute()));
    }
}

int main(int argc, char **argv) {
    bool cpu_only = argc > 1 && std::strcmp(argv[1], "--cpu") == 0;
    test_limiter_and_aggregation();
    if (cpu_only) {
        // without a GPU the product must refuse loudly, never fall back
        EXPECT(throws([] { NewFlatIndex(3, Euclidean); }));
    } else {
        test_flat();
        test_trained_indexes();
        test_hnsw();
        test_serialisation_flush_and_shards();
    }
    std::printf(failures ? "FAILED (%d)\n" : "ok\n", failures);
    return failures ? 1 : 0;
}
