// Replays the reference's own unit tests (src/index/flat.rs:187-252, src/index/hnsw.rs:530-662,
// src/lib.rs:681-694) through the C++ host mirror.  Prints "CPP_MIRROR PASS" on success.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "vectorlite.hpp"
using namespace vectorlite;

#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
    {   // flat.rs:187-201
        FlatIndex idx(3, {{1, {1, 0, 0}, "test", {}}, {2, {0, 1, 0}, "test", {}}, {3, {0, 0, 1}, "test", {}}});
        auto r = idx.search({1.0, 0.0, 0.0}, 2, SimilarityMetric::Cosine);
        CHECK(r.size() == 2 && r[0].id == 1 && std::fabs(r[0].score - 1.0) < 1e-10 && r[1].id == 2 && r[0].text == "test");
        CHECK(idx.len() == 3 && idx.dimension() == 3 && *idx.max_id() == 3);
        bool threw = false;
        try { idx.add({4, {1, 2}, "", {}}); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "dimension") != nullptr; }
        CHECK(threw);
        threw = false;
        try { idx.add({2, {1, 2, 3}, "", {}}); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "already exists") != nullptr; }
        CHECK(threw);
        threw = false;
        try { idx.search({1.0, 0.0}, 1, SimilarityMetric::Cosine); } catch (const DimensionMismatch& e) { threw = e.expected == 3 && e.actual == 2; }
        CHECK(threw);
        idx.remove(99);  // flat.rs:93-96
        idx.remove(2);
        CHECK(idx.len() == 2 && !idx.get_vector(2) && idx.get_vector(3)->values[2] == 1.0);
    }
    {   // flat.rs:204-252
        FlatIndex e(2, {{1, {0, 0}, "", {}}, {2, {3, 4}, "", {}}, {3, {6, 8}, "", {}}});
        auto r = e.search({0, 0}, 2, SimilarityMetric::Euclidean);
        CHECK(r[0].id == 1 && r[0].score == 1.0 && r[1].id == 2 && r[1].score == 1.0 / 6.0);
        r = e.search({0, 0}, 2, SimilarityMetric::Manhattan);
        CHECK(r[0].score == 1.0 && r[1].score == 0.125);
        FlatIndex d(2, {{1, {1, 2}, "", {}}, {2, {2, 1}, "", {}}, {3, {0, 0}, "", {}}});
        r = d.search({1, 2}, 2, SimilarityMetric::DotProduct);
        CHECK(r[0].id == 1 && r[0].score == 5.0 && r[1].score == 4.0);
    }
    {   // the same trait over a store row-sharded across devices of one process (three shards on device 0 here)
        ShardedFlatIndex sh(3, {0, 0, 0}, 2);
        const double rows[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {0, 1, 0}};
        for (uint64_t i = 0; i < 7; ++i) sh.add({i + 10, {rows[i][0], rows[i][1], rows[i][2]}, "t" + std::to_string(i), {}});
        auto sizes = sh.shard_sizes();
        CHECK(sizes[0] == 2 && sizes[1] == 2 && sizes[2] == 3 && sh.len() == 7 && *sh.max_id() == 16);
        auto r = sh.search({0.0, 1.0, 0.0}, 4, SimilarityMetric::Cosine);   // ids 11, 13, 16 tie at 1.0 across all three shards
        CHECK(r.size() == 4 && r[0].id == 11 && r[1].id == 13 && r[2].id == 16 && r[3].id == 14 && r[0].score == 1.0 && r[1].text == "t3");
        bool threw = false;
        try { sh.add({13, {1, 2, 3}, "", {}}); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "already exists") != nullptr; }
        CHECK(threw);
        sh.remove(13); sh.remove(999);
        r = sh.search({0.0, 1.0, 0.0}, 2, SimilarityMetric::Cosine);
        CHECK(r.size() == 2 && r[0].id == 11 && r[1].id == 16 && !sh.get_vector(13) && sh.get_vector(16)->values[1] == 1.0);
        threw = false;
        try { sh.search({1.0, 0.0}, 1, SimilarityMetric::Cosine); } catch (const DimensionMismatch& e) { threw = e.expected == 3 && e.actual == 2; }
        CHECK(threw);
    }
    {   // hnsw.rs:605-662
        HNSWIndex h(3, SimilarityMetric::Euclidean);
        CHECK(h.is_empty() && h.dimension() == 3);
        h.add({100, {1, 0, 0}, "a", {}}); h.add({200, {0, 1, 0}, "", {}}); h.add({300, {0, 0, 1}, "", {}}); h.add({400, {1, 1, 0}, "", {}});
        auto r = h.search({1.1, 0.1, 0.1}, 2, SimilarityMetric::Euclidean);
        CHECK(!r.empty() && r.size() <= 2 && r[0].id == 100 && r[0].text == "a");
        bool threw = false;
        try { h.search({1.1, 0.1, 0.1}, 2, SimilarityMetric::Cosine); } catch (const MetricMismatch&) { threw = true; }
        CHECK(threw);
        threw = false;
        try { h.remove(999); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "does not exist") != nullptr; }
        CHECK(threw);
        h.remove(100);
        CHECK(h.len() == 3 && !h.get_vector(100));
        VectorIndexWrapper w(std::make_unique<HNSWIndex>(3, SimilarityMetric::Cosine));
        CHECK(w.index_type() == IndexType::HNSW && *w.metric() == SimilarityMetric::Cosine);
    }
    std::printf("CPP_MIRROR PASS\n");
    return 0;
}
