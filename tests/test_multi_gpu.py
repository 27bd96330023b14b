"""Shard-aware indexes behind the VectorIndex interface (vectorlite_b200/multi_gpu.py): routing
of inserts into contiguous storage-order ranges, order-preserving deletes, re-splits, and bit-exact parity of the
merged top-k with the oracle on the whole store — exercised on ONE device by placing several shards on it (the
routing and the merge do not care where a shard lives)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vl():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import vectorlite_b200
    vectorlite_b200.lib()
    return vectorlite_b200


def _hex(a):
    return [float(x).hex() for x in a]


def _check(oracle_mod, idx, rows, ids, queries, k, metric):
    gi, gs, gc = idx.search_batch(queries, k, metric)
    for qi in range(queries.shape[0]):
        st, oi, os_ = oracle_mod.flat_search(rows, ids, queries[qi], k, int(metric))
        assert st == 0
        c = int(gc[qi])
        assert c == len(oi), (metric, k, c, len(oi))
        assert list(map(int, gi[qi, :c])) == list(map(int, oi)), (metric, k, qi)
        assert _hex(gs[qi, :c]) == _hex(os_), (metric, k, qi)
        assert all(int(x) == 2**64 - 1 for x in gi[qi, c:])


def test_insert_routing_and_resplit(vl, oracle_mod):
    from vectorlite_b200.multi_gpu import MultiGpuFlatIndex
    dim = 8
    rows = oracle_mod.synth_rows(42, 0, 400, dim)
    idx = MultiGpuFlatIndex(dim, [0, 0, 0], shard_rows=100)
    assert idx.is_empty() and idx.max_id() is None and idx.search(rows[0], 3, vl.SimilarityMetric.Cosine) == []
    for i in range(250):                                   # appends fill the tail shard, then move on
        idx.add(vl.Vector(i, rows[i], f"t{i}", {"i": i}))
    assert idx.shard_sizes() == [100, 100, 50] and idx.len() == 250 and idx.max_id() == 249
    assert idx.shard_of(0) == 0 and idx.shard_of(150) == 1 and idx.shard_of(249) == 2
    with pytest.raises(ValueError, match="already exists"):  # flat.rs:87 — over the whole store, not one shard
        idx.add(vl.Vector(5, rows[5]))
    with pytest.raises(ValueError, match="dimension"):       # flat.rs:84
        idx.add(vl.Vector(999, [1.0, 2.0]))
    idx.delete(12345)                                        # flat.rs:93-96: deleting a missing id is Ok
    idx.delete(10)                                           # frees a slot in shard 0; the tail does not move back
    idx.add(vl.Vector(250, rows[250]))
    assert idx.shard_sizes() == [99, 100, 51]
    for i in range(251, 300):
        idx.add(vl.Vector(i, rows[i]))
    assert idx.shard_sizes() == [99, 100, 100]
    idx.add(vl.Vector(300, rows[300]))                       # last shard full → even re-split with headroom
    sizes = idx.shard_sizes()
    assert sum(sizes) == 300 and max(sizes) - min(sizes) <= 1, sizes
    live = np.array([i for i in range(301) if i != 10], dtype=np.uint64)
    eids, erows = idx.export()                               # global storage order survives the re-split
    assert np.array_equal(eids, live) and np.array_equal(erows, rows[live.astype(np.int64)])
    for metric in vl.SimilarityMetric:
        _check(oracle_mod, idx, rows[live.astype(np.int64)], live, rows[300:304], 7, metric)
    v = idx.get_vector(42)
    assert v.text == "t42" and v.metadata == {"i": 42} and np.array_equal(np.asarray(v.values, dtype=np.float32), rows[42])
    assert idx.get_vector(10) is None
    res = idx.search(rows[42], 1, vl.SimilarityMetric.Euclidean)
    assert res[0].id == 42 and res[0].text == "t42" and res[0].score == 1.0
    with pytest.raises(vl.DimensionMismatch):                # flat.rs:99-104
        idx.search([1.0, 2.0], 1, vl.SimilarityMetric.Cosine)
    idx.close()


def test_sharded_parity_with_ties_across_shards(vl, oracle_mod):
    """Bulk load split evenly over 3 shards; duplicates of one row live in different shards, so the merged order
    must fall back to global insertion order exactly like the reference's stable sort."""
    from vectorlite_b200.multi_gpu import MultiGpuFlatIndex
    n, dim, k = 21000, 384, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    rows[9000] = rows[5]
    rows[20000] = rows[5]
    rows[15000] = rows[8000]
    q = oracle_mod.synth_rows(43, 0, 5, dim)
    q[1] = rows[5]
    q[2] = rows[8000]
    ids = np.arange(1000, 1000 + n, dtype=np.uint64)         # ids != positions
    idx = MultiGpuFlatIndex(dim, [0, 0, 0])
    idx.add_batch(ids, rows)
    assert idx.shard_sizes() == [7000, 7000, 7000]
    for metric in vl.SimilarityMetric:
        _check(oracle_mod, idx, rows, ids, q, k, metric)
    _check(oracle_mod, idx, rows, ids, q[:2], 100, vl.SimilarityMetric.Cosine)
    # k larger than one shard's share of the hits, and larger than the store
    small = MultiGpuFlatIndex(dim, [0, 0], shard_rows=8)
    for i in range(13):
        small.add(vl.Vector(i, rows[i]))
    _check(oracle_mod, small, rows[:13], np.arange(13, dtype=np.uint64), q[:2], 50, vl.SimilarityMetric.DotProduct)
    small.close()
    # order-preserving deletes in two shards (one of the tied rows among them)
    for d in (1005, 9500, 10000):
        idx.delete(d)
    keep = np.array([i for i in range(n) if i + 1000 not in (1005, 9500, 10000)], dtype=np.int64)
    assert idx.shard_sizes() == [6999, 6998, 7000] and idx.len() == n - 3
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.Manhattan):
        _check(oracle_mod, idx, rows[keep], ids[keep], q, k, metric)
    idx.close()


def test_collection_over_sharded_index_and_persistence(vl, oracle_mod, tmp_path):
    """The Collection / client layer is unchanged above a shard-aware index (client.rs:243-431), and a saved
    collection loads back into a sharded one with a single bulk upload per shard."""
    from vectorlite_b200 import collection as col

    class Emb:
        def generate_embedding(self, text):
            rng = np.random.default_rng(abs(hash(text)) % (2**32))
            v = rng.standard_normal(16)
            return list(v / np.linalg.norm(v))

        def dimension(self):
            return 16

    client = col.VectorLiteClient(Emb(), devices=[0, 0], shard_rows=5)
    client.create_collection("docs", vl.IndexType.Flat)
    ids = [client.add_text_to_collection("docs", f"text {i}", {"n": i}) for i in range(12)]
    assert ids == list(range(12))
    c = client.get_collection("docs")
    assert c.index_read().num_shards() == 2 and sum(c.index_read().shard_sizes()) == 12
    hit = client.search_text_in_collection("docs", "text 7", 3, vl.SimilarityMetric.Cosine)
    assert hit[0].id == 7 and hit[0].text == "text 7" and hit[0].metadata == {"n": 7}
    client.delete_from_collection("docs", 7)
    assert client.get_vector_from_collection("docs", 7) is None and client.get_collection_info("docs").count == 11
    path = str(tmp_path / "docs.vlc")
    c.save_to_file(path)
    back = col.Collection.load_from_file(path, devices=[0, 0, 0])
    assert back.index_read().num_shards() == 3 and back.get_info().count == 11 and back.next_id() == 12
    a = c.search_vector(Emb().generate_embedding("text 3"), 5, vl.SimilarityMetric.Euclidean)
    b = back.search_vector(Emb().generate_embedding("text 3"), 5, vl.SimilarityMetric.Euclidean)
    assert [(r.id, r.score.hex(), r.text) for r in a] == [(r.id, r.score.hex(), r.text) for r in b]


def test_hnsw_replicas_behind_the_index_interface(vl, oracle_mod):
    from vectorlite_b200.multi_gpu import MultiGpuHnswIndex
    n, dim, k = 6000, 64, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    q = oracle_mod.synth_rows(43, 0, 64, dim)
    M = vl.SimilarityMetric
    idx = MultiGpuHnswIndex(dim, M.Cosine, [0, 0], ef_construction=100)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.build()
    assert idx.num_replicas() == 2 and idx.len() == n and idx.metric() == M.Cosine and idx.max_id() == n - 1
    flat = vl.FlatIndex(dim)
    flat.add_batch(np.arange(n, dtype=np.uint64), rows)
    truth, _, _ = flat.search_batch(q, k, M.Cosine)
    gi, gs, gc = idx.search_batch(q, k, M.Cosine, 64)        # 64 queries: split over both replicas
    assert gi.shape == (64, k) and all(int(c) == k for c in gc)
    recall = sum(len(set(map(int, gi[i])) & set(map(int, truth[i]))) for i in range(64)) / (64 * k)
    assert recall >= 0.9, recall
    one = [idx.search(q[0], k, M.Cosine, 64) for _ in range(2)]     # single queries rotate over the replicas
    assert len(one[0]) == k and len(one[1]) == k
    with pytest.raises(vl.MetricMismatch):                   # hnsw.rs:425-430
        idx.search(q[0], k, M.Euclidean)
    idx.delete(int(gi[0, 0]))                                # soft delete reaches every replica
    for _ in range(2):
        assert int(gi[0, 0]) not in [r.id for r in idx.search(q[0], k, M.Cosine, 64)]
    with pytest.raises(ValueError, match="does not exist"):  # hnsw.rs:402
        idx.delete(10**9)
    idx.close()
