"""Collection / client / persistence layer replaying the reference's own client and
persistence tests (src/client.rs:499-850, src/persistence.rs:178-352, tests/persistence_api_test.rs)
against CUDA-backed indexes, plus the additive vector / batch search and the micro-batcher."""
import json
import threading

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


class MockEmbedding:                       # client.rs:504-523: vec![1.0; dim]
    def __init__(self, dim=3):
        self.dim = dim

    def generate_embedding(self, text):
        return [1.0] * self.dim

    def dimension(self):
        return self.dim


class HashEmbedding:                       # embeddings.rs:295-342 flavour: text-dependent, unit norm
    def __init__(self, dim=16):
        self.dim = dim

    def generate_embedding(self, text):
        rng = np.random.default_rng(abs(hash(text)) % (2**32))
        v = rng.standard_normal(self.dim)
        return list(v / np.linalg.norm(v))

    def dimension(self):
        return self.dim


def test_client_collection_semantics(vl):
    from vectorlite_b200 import collection as col
    M, T = vl.SimilarityMetric, vl.IndexType
    client = col.VectorLiteClient(MockEmbedding(3))
    client.create_collection("docs", T.Flat)
    with pytest.raises(col.CollectionAlreadyExists):                 # client.rs:84-86
        client.create_collection("docs", T.Flat)
    with pytest.raises(col.InvalidRequest):                          # client.rs:709-722 HNSW needs a metric
        client.create_collection("h", T.HNSW)
    client.create_collection("h", T.HNSW, M.Euclidean)
    assert sorted(client.list_collections()) == ["docs", "h"] and client.has_collection("docs")
    ids = [client.add_text_to_collection("docs", f"text {i}") for i in range(3)]
    assert ids == [0, 1, 2]                                          # client.rs:600-622 ids start at 0
    info = client.get_collection_info("docs")
    assert (info.name, info.count, info.is_empty, info.dimension) == ("docs", 3, False, 3)
    r = client.search_text_in_collection("docs", "anything", 1)      # default metric Cosine (client.rs:150-153)
    assert len(r) == 1 and r[0].id == 0 and r[0].text == "text 0"     # all-equal embeddings → first inserted (client.rs:665-667)
    client.add_text_to_collection("h", "hello", {"a": 1})
    r = client.search_text_in_collection("h", "q", 5)                # metric None → the HNSW's own metric
    assert len(r) == 1 and r[0].metadata == {"a": 1}
    with pytest.raises(col.VectorNotFound):                          # client.rs:384-390
        client.delete_from_collection("h", 99)
    client.delete_from_collection("docs", 99)                        # flat: deleting a missing id is Ok
    client.delete_from_collection("docs", 1)
    assert client.get_vector_from_collection("docs", 1) is None
    assert client.get_collection("docs").next_id() == 3
    with pytest.raises(col.CollectionNotFound):
        client.add_text_to_collection("nope", "x")
    # a failed add still consumes an id (client.rs:350)
    c = client.get_collection("docs")
    with pytest.raises(vl.DimensionMismatch):
        c.add_text("bad", MockEmbedding(2))
    assert c.next_id() == 4
    client.delete_collection("h")
    assert not client.has_collection("h")


@pytest.mark.parametrize("kind", ["flat", "hnsw"])
def test_persistence_roundtrip_reference_format(vl, tmp_path, kind):
    from vectorlite_b200 import collection as col
    M = vl.SimilarityMetric
    emb = HashEmbedding(16)
    index = vl.FlatIndex(16) if kind == "flat" else vl.HNSWIndex(16, M.Cosine)
    c = col.Collection("persist", index)
    for i in range(40):
        c.add_text_with_metadata(f"doc {i}", {"i": i} if i % 2 else None, emb)
    c.delete(7)
    path = str(tmp_path / "sub" / "persist.vlc")
    c.save_to_file(path)
    doc = json.load(open(path))
    assert doc["header"]["version"] == "1.0.0" and doc["header"]["format"] == "vectorlite-collection"   # persistence.rs:88-96
    assert doc["metadata"]["name"] == "persist" and doc["metadata"]["vector_count"] == 39
    assert doc["metadata"]["index_type"] == ("Flat" if kind == "flat" else "HNSW")
    if kind == "flat":
        d = doc["index"]["Flat"]                                        # flat.rs:59-65 JSON shape
        assert d["dim"] == 16 and [e["id"] for e in d["data"]] == [i for i in range(40) if i != 7]
        assert set(d["data"][0]) == {"id", "values", "text", "metadata"}
    else:
        d = doc["index"]["HNSW"]                                        # hnsw.rs:197-213 JSON shape
        assert d["metric"] == "Cosine" and set(d) >= {"dim", "metric", "id_to_index", "index_to_id", "metadata", "vector_values"}
        assert "7" not in d["vector_values"] and len(d["vector_values"]) == 39
    c2 = col.Collection.load_from_file(path)
    assert c2.name() == "persist" and c2.get_info().count == 39 and c2.next_id() == 40   # client.rs:297-308
    q = emb.generate_embedding("doc 11")
    a = c.search_vector(q, 5, M.Cosine, ef=64)
    b = c2.search_vector(q, 5, M.Cosine, ef=64)
    assert a[0].id == 11 and abs(a[0].score - 1.0) < 1e-6
    assert [(r.id, r.text, r.metadata) for r in a] == [(r.id, r.text, r.metadata) for r in b]
    if kind == "flat":
        assert [r.score for r in a] == [r.score for r in b]
    assert c2.get_vector(7) is None and c2.get_vector(8).text == "doc 8"
    # error paths: persistence.rs:149-176
    with pytest.raises(col.FileNotFound):
        col.load_collection_from_file(str(tmp_path / "missing.vlc"))
    doc["header"]["version"] = "2.0.0"
    bad = tmp_path / "bad.vlc"
    bad.write_text(json.dumps(doc))
    with pytest.raises(col.VersionMismatch):
        col.load_collection_from_file(str(bad))
    doc["header"]["version"] = "1.0.0"
    doc["header"]["format"] = "other"
    bad.write_text(json.dumps(doc))
    with pytest.raises(col.InvalidFormat):
        col.load_collection_from_file(str(bad))


def test_loads_a_file_written_in_the_reference_shape(vl, tmp_path):
    """A .vlc document exactly as serde_json::to_string_pretty writes it (f64 values, RFC3339 dates)."""
    from vectorlite_b200 import collection as col
    doc = {"header": {"version": "1.0.0", "format": "vectorlite-collection", "created_at": "2025-01-01T00:00:00.123456789Z"},
           "metadata": {"name": "ref", "created_at": "2025-01-01T00:00:00Z", "vector_count": 2, "dimension": 3, "index_type": "Flat"},
           "index": {"Flat": {"dim": 3, "data": [
               {"id": 0, "values": [1.0, 2.0, 3.0], "text": "First document", "metadata": None},
               {"id": 1, "values": [4.0, 5.0, 6.0], "text": "Second document", "metadata": {"k": "v"}}]}}}
    p = tmp_path / "ref.vlc"
    p.write_text(json.dumps(doc, indent=2))
    c = col.load_collection_from_file(str(p))
    r = c.search_vector([1.1, 2.1, 3.1], 1, vl.SimilarityMetric.Cosine)      # persistence.rs:247-249
    assert r[0].id == 0 and r[0].text == "First document" and abs(r[0].score - 0.9998592903536574) < 1e-6
    assert c.next_id() == 2


def test_batch_search_and_micro_batcher(vl):
    from vectorlite_b200 import collection as col
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((5000, 32)).astype(np.float32)
    c = col.Collection("mb", vl.FlatIndex(32))
    assert list(c.add_vectors(rows, texts=[f"t{i}" for i in range(5000)])) == list(range(5000))
    qs = rng.standard_normal((64, 32)).astype(np.float32)
    single = [c.search_vector(q, 5, vl.SimilarityMetric.Euclidean) for q in qs]
    batched = c.search_batch(qs, 5, vl.SimilarityMetric.Euclidean)
    assert [[(r.id, r.score, r.text) for r in a] for a in single] == [[(r.id, r.score, r.text) for r in b] for b in batched]
    c.enable_micro_batching(max_batch=64, max_wait_us=20000)
    out = [None] * 64

    def worker(i):
        out[i] = c.search_vector(qs[i], 5, vl.SimilarityMetric.Euclidean)
    th = [threading.Thread(target=worker, args=(i,)) for i in range(64)]
    [t.start() for t in th]
    [t.join() for t in th]
    served = list(c._batcher.batches_served)
    c.disable_micro_batching()
    assert [[(r.id, r.score) for r in a] for a in out] == [[(r.id, r.score) for r in a] for a in single]
    assert sum(served) == 64 and max(served) > 1          # concurrent callers really shared launches


def test_hnsw_graph_is_persisted_and_restored(vl, tmp_path):
    """the reference rebuilds its HNSW graph on every load, in HashMap order (hnsw.rs:199-200,
    322-348) — a different graph each time.  The optional `.graph` side file restores the SAVED graph: same
    adjacency (structure audit equal), same answers bit for bit, nothing rebuilt; a corrupt side file or a graph
    with soft-deleted nodes falls back to the reference behaviour (rebuild)."""
    import oracle
    from vectorlite_b200 import collection as col
    n, dim, k = 6000, 48, 10
    M = vl.SimilarityMetric
    rows = oracle.synth_rows(42, 0, n, dim, 32)
    q = oracle.synth_rows(43, 0, 32, dim, 32)
    ids = (np.arange(n, dtype=np.uint64) * 7 + 3)[::-1].copy()          # not sorted: the order must come from the side file
    h = vl.HNSWIndex(dim, M.Cosine, ef_construction=100)
    h.add_batch(ids, rows, [f"t{int(i)}" for i in ids], None)
    c = col.Collection("g", h)
    want = h.search_batch(q, k, M.Cosine, 16)
    path = str(tmp_path / "g.vlc")
    c.save_to_file(path)
    import os
    assert os.path.exists(path + col.GRAPH_SUFFIX)
    c2 = col.Collection.load_from_file(path)
    h2 = c2.index_read()
    assert h2.len() == n and h2.graph_check() == h.graph_check()
    assert h2.build_info()["builder"] == "none"                          # nothing was built
    got = h2.search_batch(q, k, M.Cosine, 16)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint64), want[1].view(np.uint64))
    assert h2.search(q[0], 1, M.Cosine)[0].text.startswith("t")
    # the restored index keeps working as an index: incremental add + search
    h2.add(vl.Vector(10**9, q[1]))
    assert h2.search(q[1], 1, M.Cosine)[0].id == 10**9
    # the .vlc alone is still a valid reference-format file: without the side file the graph is rebuilt
    os.remove(path + col.GRAPH_SUFFIX)
    c3 = col.Collection.load_from_file(path)
    assert c3.index_read().len() == n and c3.index_read().graph_check()["invalid"] == 0
    # corrupt side file → rebuild, never a broken graph
    c.save_to_file(path)
    with open(path + col.GRAPH_SUFFIX, "r+b") as f:
        f.seek(16 + 8 * n + 200)
        f.write(b"\xff" * 64)
    c4 = col.Collection.load_from_file(path)
    assert c4.index_read().len() == n and c4.index_read().graph_check()["invalid"] == 0
    # soft-deleted nodes: no side file is written (their rows are not exported)
    h.delete(int(ids[5]))
    c.save_to_file(path)
    assert not os.path.exists(path + col.GRAPH_SUFFIX)
    assert col.Collection.load_from_file(path).index_read().len() == n - 1
