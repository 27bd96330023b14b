// vectorlite.hpp — header-only C++ host mirror of VectorLite's index interface over the C ABI
// (include/vectorlite_cuda.h).  Same names, argument meaning and error behaviour as the reference:
//   SimilarityMetric             src/lib.rs:363-378
//   Vector / SearchResult        src/lib.rs:163-174, 193-203
//   VectorIndex (abstract)       src/lib.rs:224-245
//   FlatIndex                    src/index/flat.rs:59-136
//   HNSWIndex                    src/index/hnsw.rs:197-518
//   VectorIndexWrapper           src/lib.rs:270-346
// `Result<(), String>` errors become std::runtime_error carrying the reference's message (the callers
// at src/client.rs:334-345,366-377 substring-match "dimension" / "already exists" / "does not exist");
// `VectorLiteError::{DimensionMismatch, MetricMismatch}` become the exception types below.
// Text / metadata stay on the host and are attached to the <= k hits only.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "vectorlite_cuda.h"

namespace vectorlite {

enum class SimilarityMetric : int { Cosine = 0, Euclidean = 1, Manhattan = 2, DotProduct = 3 };
enum class IndexType : int { Flat = 0, HNSW = 1 };

struct Vector {
    uint64_t id = 0;
    std::vector<double> values;
    std::string text;
    std::optional<std::string> metadata;  // serialized JSON (serde_json::Value in the reference)
};
struct SearchResult {
    uint64_t id = 0;
    double score = 0.0;
    std::string text;
    std::optional<std::string> metadata;
};

struct VectorLiteError : std::runtime_error {
    int code;
    VectorLiteError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct DimensionMismatch : VectorLiteError {
    size_t expected, actual;
    DimensionMismatch(size_t e, size_t a)
        : VectorLiteError(VL_ERR_DIM, "Dimension mismatch: expected " + std::to_string(e) + ", got " + std::to_string(a)),
          expected(e), actual(a) {}
};
struct MetricMismatch : VectorLiteError {
    SimilarityMetric requested, index;
    MetricMismatch(SimilarityMetric r, SimilarityMetric i) : VectorLiteError(VL_ERR_METRIC_MISMATCH, "Metric mismatch"), requested(r), index(i) {}
};

class VectorIndex {  // src/lib.rs:224-245
public:
    virtual ~VectorIndex() = default;
    virtual void add(const Vector& v) = 0;
    virtual void remove(uint64_t id) = 0;  // `delete` is a C++ keyword
    virtual std::vector<SearchResult> search(const std::vector<double>& query, size_t k, SimilarityMetric m) const = 0;
    virtual size_t len() const = 0;
    virtual bool is_empty() const { return len() == 0; }
    virtual std::optional<Vector> get_vector(uint64_t id) const = 0;
    virtual size_t dimension() const = 0;
};

class CudaIndex : public VectorIndex {
protected:
    vl_index* h_ = nullptr;
    struct Meta { std::string text; std::optional<std::string> metadata; };
    std::unordered_map<uint64_t, Meta> meta_;
    static std::string err() { return vl_last_error(); }

public:
    CudaIndex() = default;
    CudaIndex(const CudaIndex&) = delete;
    CudaIndex& operator=(const CudaIndex&) = delete;
    ~CudaIndex() override { vl_index_destroy(h_); }

    void add(const Vector& v) override {
        const int st = vl_index_add_f64(h_, v.id, v.values.data(), static_cast<uint32_t>(v.values.size()));
        if (st != VL_OK) throw std::runtime_error(err());
        meta_[v.id] = Meta{v.text, v.metadata};
    }
    void remove(uint64_t id) override {
        if (vl_index_delete(h_, id) != VL_OK) throw std::runtime_error(err());
        meta_.erase(id);
    }
    std::vector<SearchResult> search(const std::vector<double>& q, size_t k, SimilarityMetric m) const override {
        return search_ef(q, k, m, 0);
    }
    std::vector<SearchResult> search_ef(const std::vector<double>& q, size_t k, SimilarityMetric m, uint32_t ef) const {
        std::vector<uint64_t> ids(k);
        std::vector<double> scores(k);
        uint32_t count = 0;
        const int st = vl_index_search_f64(h_, q.data(), 1, static_cast<uint32_t>(q.size()), static_cast<uint32_t>(k),
                                           static_cast<int>(m), ef, ids.data(), scores.data(), &count);
        if (st == VL_ERR_DIM) throw DimensionMismatch(dimension(), q.size());
        if (st == VL_ERR_METRIC_MISMATCH) throw MetricMismatch(m, static_cast<SimilarityMetric>(vl_index_metric(h_)));
        if (st != VL_OK) throw VectorLiteError(st, err());
        std::vector<SearchResult> out(count);
        for (uint32_t i = 0; i < count; ++i) {
            out[i].id = ids[i];
            out[i].score = scores[i];
            auto it = meta_.find(ids[i]);
            if (it != meta_.end()) { out[i].text = it->second.text; out[i].metadata = it->second.metadata; }
        }
        return out;
    }
    size_t len() const override { return static_cast<size_t>(vl_index_len(h_)); }
    std::optional<Vector> get_vector(uint64_t id) const override {
        std::vector<float> f(dimension());
        if (vl_index_get_vector(h_, id, f.data()) != VL_OK) return std::nullopt;
        Vector v;
        v.id = id;
        v.values.assign(f.begin(), f.end());
        auto it = meta_.find(id);
        if (it != meta_.end()) { v.text = it->second.text; v.metadata = it->second.metadata; }
        return v;
    }
    size_t dimension() const override { return vl_index_dim(h_); }
    std::optional<uint64_t> max_id() const {  // flat.rs:76-78, hnsw.rs:267-269
        uint64_t id = 0;
        if (vl_index_max_id(h_, &id) != VL_OK) return std::nullopt;
        return id;
    }
    vl_index* handle() const { return h_; }
};

class FlatIndex : public CudaIndex {  // src/index/flat.rs:59-136
public:
    explicit FlatIndex(size_t dim, const std::vector<Vector>& data = {}, int device = 0) {
        if (vl_flat_create(static_cast<uint32_t>(dim), device, &h_) != VL_OK) throw VectorLiteError(VL_ERR_CUDA, err());
        for (const Vector& v : data) add(v);  // FlatIndex::new takes the vectors as given (flat.rs:68)
    }
    std::optional<SimilarityMetric> metric() const { return std::nullopt; }
    IndexType index_type() const { return IndexType::Flat; }
};

class HNSWIndex : public CudaIndex {  // src/index/hnsw.rs:197-518
public:
    // profiles of hnsw.rs:95-109 are runtime parameters: 16/32 default, 8/16 memory-optimized, 32/64 high-accuracy
    HNSWIndex(size_t dim, SimilarityMetric metric, uint32_t M = 16, uint32_t M0 = 32, uint32_t ef_construction = 400,
              int device = 0) {
        if (dim == 0) throw std::invalid_argument("HNSW index dimension cannot be 0");  // hnsw.rs:217-219 panics
        if (vl_hnsw_create(static_cast<uint32_t>(dim), static_cast<int>(metric), M, M0, ef_construction, device, &h_) != VL_OK)
            throw VectorLiteError(VL_ERR_CUDA, err());
    }
    SimilarityMetric metric() const { return static_cast<SimilarityMetric>(vl_index_metric(h_)); }
    IndexType index_type() const { return IndexType::HNSW; }
    // device beam = factor x ef, ef = the reference's min(k, len) (hnsw.rs:437); 1 = equal ef, default 8
    void set_beam_factor(uint32_t factor) {
        if (vl_hnsw_set_beam_factor(h_, factor) != VL_OK) throw VectorLiteError(VL_ERR_INVALID, err());
    }
    // Graph persistence: what `#[serde(skip)] index_internal` (hnsw.rs:199-200) leaves out.  export_graph() throws
    // (VL_ERR_UNSUPPORTED) when the graph holds soft-deleted nodes; import_graph() needs an EMPTY index and the rows
    // in the order they were exported (vl_index_export).
    std::vector<unsigned char> export_graph() const {
        uint64_t nb = 0, w = 0;
        if (vl_hnsw_graph_bytes(h_, &nb) != VL_OK) throw VectorLiteError(VL_ERR_INVALID, err());
        std::vector<unsigned char> blob(nb);
        const int st = vl_hnsw_export_graph(h_, blob.data(), nb, &w);
        if (st != VL_OK) throw VectorLiteError(st, err());
        blob.resize(w);
        return blob;
    }
    void import_graph(const std::vector<uint64_t>& ids, const std::vector<float>& rows, const std::vector<unsigned char>& blob) {
        const int st = vl_hnsw_import_graph(h_, ids.data(), rows.data(), ids.size(), blob.data(), blob.size());
        if (st != VL_OK) throw VectorLiteError(st, err());
    }
};

// A flat store row-sharded over several devices of ONE process behind the same trait: shard g holds
// the contiguous storage-order range [base_g, base_g + n_g); appends go to the shard that owns the tail and move
// on when it holds `shard_rows` rows (the last shard keeps growing; vectorlite_b200/multi_gpu.py adds the even
// re-split); searches run on every shard at once through vl_group_search, whose stable merge in shard order is
// the stable sort of flat.rs:116 over the whole store.
class ShardedFlatIndex : public VectorIndex {
    std::vector<std::unique_ptr<FlatIndex>> shards_;
    std::unordered_map<uint64_t, size_t> where_;   // id -> shard
    vl_group* group_ = nullptr;
    size_t dim_, shard_rows_, tail_ = 0;

public:
    ShardedFlatIndex(size_t dim, const std::vector<int>& devices, size_t shard_rows = size_t(1) << 20)
        : dim_(dim), shard_rows_(shard_rows ? shard_rows : 1) {
        if (devices.empty()) throw std::invalid_argument("at least one device is required");
        std::vector<vl_index*> hs;
        for (int d : devices) {
            shards_.push_back(std::make_unique<FlatIndex>(dim, std::vector<Vector>{}, d));
            hs.push_back(shards_.back()->handle());
        }
        if (vl_group_create(hs.data(), static_cast<uint32_t>(hs.size()), &group_) != VL_OK)
            throw VectorLiteError(VL_ERR_INVALID, vl_last_error());
    }
    ShardedFlatIndex(const ShardedFlatIndex&) = delete;
    ShardedFlatIndex& operator=(const ShardedFlatIndex&) = delete;
    ~ShardedFlatIndex() override { vl_group_destroy(group_); }   // before the shards it borrows

    void add(const Vector& v) override {
        if (v.values.size() != dim_) throw std::runtime_error("Vector dimension mismatch");          // flat.rs:84
        if (where_.count(v.id)) throw std::runtime_error("Vector ID " + std::to_string(v.id) + " already exists");  // flat.rs:87
        while (tail_ + 1 < shards_.size() && shards_[tail_]->len() >= shard_rows_) ++tail_;
        shards_[tail_]->add(v);
        where_[v.id] = tail_;
    }
    void remove(uint64_t id) override {                          // flat.rs:93-96: missing id is Ok
        auto it = where_.find(id);
        if (it == where_.end()) return;
        shards_[it->second]->remove(id);
        where_.erase(it);
    }
    std::vector<SearchResult> search(const std::vector<double>& q, size_t k, SimilarityMetric m) const override {
        std::vector<float> qf(q.begin(), q.end());
        std::vector<uint64_t> ids(k);
        std::vector<double> scores(k);
        uint32_t count = 0;
        const int st = vl_group_search(group_, qf.data(), 1, static_cast<uint32_t>(qf.size()), static_cast<uint32_t>(k),
                                       static_cast<int>(m), ids.data(), scores.data(), &count);
        if (st == VL_ERR_DIM) throw DimensionMismatch(dim_, q.size());
        if (st != VL_OK) throw VectorLiteError(st, vl_last_error());
        std::vector<SearchResult> out(count);
        for (uint32_t i = 0; i < count; ++i) {
            auto v = shards_[where_.at(ids[i])]->get_vector(ids[i]);   // text / metadata live with the owning shard
            out[i].id = ids[i];
            out[i].score = scores[i];
            if (v) { out[i].text = v->text; out[i].metadata = v->metadata; }
        }
        return out;
    }
    size_t len() const override { return where_.size(); }
    std::optional<Vector> get_vector(uint64_t id) const override {
        auto it = where_.find(id);
        if (it == where_.end()) return std::nullopt;
        return shards_[it->second]->get_vector(id);
    }
    size_t dimension() const override { return dim_; }
    std::optional<uint64_t> max_id() const {
        std::optional<uint64_t> best;
        for (const auto& s : shards_) {
            auto m = s->max_id();
            if (m && (!best || *m > *best)) best = m;
        }
        return best;
    }
    std::vector<size_t> shard_sizes() const {
        std::vector<size_t> out;
        for (const auto& s : shards_) out.push_back(s->len());
        return out;
    }
    std::optional<SimilarityMetric> metric() const { return std::nullopt; }
    IndexType index_type() const { return IndexType::Flat; }
};

class VectorIndexWrapper {  // src/lib.rs:270-346
    std::unique_ptr<CudaIndex> idx_;
    IndexType type_;

public:
    explicit VectorIndexWrapper(std::unique_ptr<FlatIndex> f) : idx_(std::move(f)), type_(IndexType::Flat) {}
    explicit VectorIndexWrapper(std::unique_ptr<HNSWIndex> h) : idx_(std::move(h)), type_(IndexType::HNSW) {}
    VectorIndex& operator*() { return *idx_; }
    CudaIndex* operator->() { return idx_.get(); }
    IndexType index_type() const { return type_; }
    std::optional<SimilarityMetric> metric() const {
        if (type_ == IndexType::Flat) return std::nullopt;
        return static_cast<SimilarityMetric>(vl_index_metric(idx_->handle()));
    }
};

}  // namespace vectorlite
