"""GPU parity tests proper: the CUDA flat path, called through the C ABI, against the CPU oracle.

Bar: top-k ids identical to the oracle under the stable-sort tie-break
(score desc, insertion order asc); scores BIT-IDENTICAL in f64 (stronger than the 1e-5 the
north star asks for), all four metrics.
"""
import ctypes as C

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


def _check(vl, oracle_mod, idx, rows, ids, queries, k, metric):
    gi, gs, gc = idx.search_batch(queries, k, metric)
    for qi in range(queries.shape[0]):
        st, oi, os_ = oracle_mod.flat_search(rows, ids, queries[qi], k, int(metric))
        assert st == 0
        c = int(gc[qi])
        assert c == len(oi), (metric, k, c, len(oi))
        assert list(map(int, gi[qi, :c])) == list(map(int, oi)), (metric, k, qi)
        assert _hex(gs[qi, :c]) == _hex(os_), (metric, k, qi)
        assert all(int(x) == 2**64 - 1 for x in gi[qi, c:])


def test_reference_kats(vl, oracle_mod, kats):
    """Every flat known-answer test of the reference's own suite (SURVEY §8c ①-⑩)."""
    for case in kats["flat"]:
        dim = len(case["query"])
        idx = vl.FlatIndex(dim, [vl.Vector(r["id"], r["values"], "test") for r in case["rows"]])
        res = idx.search(case["query"], case["k"], vl.SimilarityMetric(case["metric"]))
        assert len(res) == case["expect_len"], case["name"]
        assert [r.id for r in res] == case["exact_ids"], case["name"]
        # the reference's value (f64 inputs) within the north star's 1e-5 relative: the ABI stores
        # f32, and 1.1 / 0.1 are not f32-representable
        for got, want in zip([r.score for r in res], case["exact_scores"]):
            assert abs(got - want) <= 1e-5 * max(abs(want), 1e-30) + 1e-12, case["name"]
        # and bit-identical to the oracle evaluated on the f32-narrowed inputs
        rows32 = np.array([r["values"] for r in case["rows"]], dtype=np.float32)
        ids64 = np.array([r["id"] for r in case["rows"]], dtype=np.uint64)
        st, oi, os_ = oracle_mod.flat_search(rows32, ids64, np.array(case["query"], dtype=np.float32),
                                             case["k"], case["metric"])
        assert [r.id for r in res] == list(map(int, oi)), case["name"]
        assert _hex([r.score for r in res]) == _hex(os_), case["name"]
        for a in case["asserts"]:
            if "score" in a:
                assert abs(res[a["index"]].score - a["score"]) < a["tol"]
            if "score_gt" in a:
                assert res[a["index"]].score > a["score_gt"]
        assert res[0].text == "test"


@pytest.mark.parametrize("dim", [1, 3, 17, 100, 384, 768, 1000])
def test_random_parity_dims(vl, oracle_mod, dim):
    rng = np.random.default_rng(dim)
    for n in (1, 5, 63, 64, 65, 700, 5000):
        rows = rng.standard_normal((n, dim)).astype(np.float32)
        if n > 10:
            rows[3] = 0.0          # zero row: cosine → 0.0 branch (lib.rs:439-440)
            rows[7] = rows[6]      # exact duplicate: tie → earlier position wins
        ids = (np.arange(n, dtype=np.uint64) * 3 + 11)
        queries = rng.standard_normal((2, dim)).astype(np.float32)
        idx = vl.FlatIndex(dim)
        idx.add_batch(ids, rows)
        assert idx.len() == n and idx.dimension() == dim and idx.max_id() == int(ids[-1])
        for metric in vl.SimilarityMetric:
            for k in (1, 10, 100):
                _check(vl, oracle_mod, idx, rows, ids, queries, k, metric)
        idx.close()


def test_unit_norm_384_all_metrics_k(vl, oracle_mod):
    n, dim = 20000, 384
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    queries = oracle_mod.synth_rows(43, 0, 4, dim)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    for metric in vl.SimilarityMetric:
        for k in (1, 10, 100, 256, 300):   # 300 > over-select capacity → exact path
            _check(vl, oracle_mod, idx, rows, None, queries, k, metric)
    st = idx.stats()
    assert st["fast_queries"] > 0 and st["exact_queries"] > 0


def test_exact_mode_equals_auto(vl, oracle_mod):
    n, dim = 3000, 96
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    queries = rng.standard_normal((3, dim)).astype(np.float32)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.set_mode(vl.Mode.Exact)
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx, rows, None, queries, 10, metric)
    assert idx.stats()["fast_queries"] == 0


def test_heavy_ties_fall_back_to_exact_path(vl, oracle_mod):
    """All-equal embeddings (the reference's mock embedder, client.rs:504-523): every score ties,
    the answer is the first k inserted.  The certificate cannot hold → exact path, same answer."""
    n, dim = 2000, 384
    rows = np.tile(np.linspace(0.1, 1.0, dim, dtype=np.float32), (n, 1))
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    q = rows[:1].copy()
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx, rows, None, q, 10, metric)
    assert idx.stats()["exact_queries"] >= 4
    # near-ties: scores differ by ~1 ulp of f32 — ids must still be the oracle's
    rng = np.random.default_rng(1)
    base = rng.standard_normal(dim).astype(np.float32)
    rows2 = np.tile(base, (500, 1))
    rows2[:, 0] += (rng.integers(0, 4, 500) * 1e-7).astype(np.float32)
    idx2 = vl.FlatIndex(dim)
    idx2.add_batch(np.arange(500, dtype=np.uint64), rows2)
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx2, rows2, None, base[None, :], 10, metric)


def test_zero_query_and_zero_rows(vl, oracle_mod):
    dim = 8
    rows = np.zeros((10, dim), dtype=np.float32)
    rows[4, 2] = 1.0
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(10, dtype=np.uint64), rows)
    qs = np.zeros((2, dim), dtype=np.float32)
    qs[1, 2] = 2.0
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx, rows, None, qs, 3, metric)


def test_flat_semantics(vl):
    idx = vl.FlatIndex(3)
    # flat.rs:99: no dimension check while empty
    assert idx.search([1.0, 2.0], 5, vl.SimilarityMetric.Cosine) == []
    idx.add(vl.Vector(1, [1.0, 0.0, 0.0], "a", {"k": 1}))
    idx.add(vl.Vector(2, [0.0, 1.0, 0.0], "b"))
    idx.add(vl.Vector(3, [0.0, 0.0, 1.0], "c"))
    with pytest.raises(ValueError, match="dimension"):       # flat.rs:84
        idx.add(vl.Vector(4, [1.0, 2.0]))
    with pytest.raises(ValueError, match="already exists"):  # flat.rs:87
        idx.add(vl.Vector(2, [1.0, 2.0, 3.0]))
    with pytest.raises(vl.DimensionMismatch) as e:           # flat.rs:100-103
        idx.search([1.0, 0.0], 1, vl.SimilarityMetric.Cosine)
    assert e.value.expected == 3 and e.value.actual == 2
    assert idx.len() == 3 and not idx.is_empty() and idx.max_id() == 3
    r = idx.search([1.0, 0.0, 0.0], 10, vl.SimilarityMetric.Cosine)      # k > n → n results
    assert [x.id for x in r] == [1, 2, 3] and r[0].text == "a" and r[0].metadata == {"k": 1}
    assert idx.search([1.0, 0.0, 0.0], 0, vl.SimilarityMetric.Cosine) == []
    v = idx.get_vector(2)
    assert v is not None and list(v.values) == [0.0, 1.0, 0.0] and v.text == "b"
    assert idx.get_vector(99) is None
    idx.delete(99)                                            # flat.rs:93-96: missing id is Ok
    idx.delete(2)                                             # order of the rest is preserved
    assert idx.len() == 2 and idx.get_vector(2) is None
    r = idx.search([0.0, 0.0, 0.0], 2, vl.SimilarityMetric.DotProduct)   # all tie → insertion order
    assert [x.id for x in r] == [1, 3]
    idx.add(vl.Vector(2, [0.0, 1.0, 0.0]))                    # re-add goes to the END
    r = idx.search([0.0, 0.0, 0.0], 3, vl.SimilarityMetric.DotProduct)
    assert [x.id for x in r] == [1, 3, 2]
    idx.delete(3)
    assert idx.max_id() == 2
    w = vl.VectorIndexWrapper(idx)
    assert w.metric() is None and w.index_type() == vl.IndexType.Flat and w.len() == 2


def test_delete_preserves_order_large(vl, oracle_mod):
    n, dim = 3000, 20
    rng = np.random.default_rng(9)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    rows[100:200] = rows[100]        # a block of duplicates: ties expose any reordering
    ids = np.arange(n, dtype=np.uint64) + 1000
    idx = vl.FlatIndex(dim)
    idx.add_batch(ids, rows)
    keep = np.ones(n, dtype=bool)
    for d in (1000, 1105, 1150, 3999, 2500):
        idx.delete(int(d))
        keep[d - 1000] = False
    rows2, ids2 = rows[keep], ids[keep]
    eids, erows = idx.export()
    assert np.array_equal(eids, ids2) and np.array_equal(erows, rows2)
    q = rows[100:101]
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx, rows2, ids2, q, 20, metric)


def test_nan_is_an_error_not_a_panic(vl):
    idx = vl.FlatIndex(2)
    idx.add(vl.Vector(1, [1.0, 0.0]))
    idx.add(vl.Vector(2, [float("nan"), 1.0]))
    with pytest.raises(vl.VectorLiteError) as e:
        idx.search([1.0, 1.0], 1, vl.SimilarityMetric.DotProduct)
    assert e.value.code == vl.VL_ERR_NAN


def test_device_generator_matches_oracle(vl, oracle_mod):
    for clusters in (0, 16):
        idx = vl.FlatIndex(384)
        idx.fill_synthetic(42, 3000, first_row=500, clusters=clusters)
        ids, rows = idx.export()
        assert np.array_equal(ids, np.arange(500, 3500, dtype=np.uint64))
        assert np.array_equal(rows, oracle_mod.synth_rows(42, 500, 3000, 384, clusters))
    idx = vl.FlatIndex(100)
    idx.fill_synthetic(7, 257)
    assert np.array_equal(idx.export()[1], oracle_mod.synth_rows(7, 0, 257, 100))


def test_search_device_and_sharded_merge(vl, oracle_mod):
    """Row-sharded flat index emulated on one GPU: two shards with position bases, device-side
    search into the all-gather layout [G][nq][k], merge kernel → equals the unsharded oracle."""
    import torch
    n, dim, k, nq, G = 6000, 384, 10, 5, 2
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    rows[3000:3010] = rows[10]       # cross-shard duplicates: tie-break must stay global
    queries = np.concatenate([oracle_mod.synth_rows(43, 0, nq - 1, dim), rows[10:11]])
    dev = torch.device("cuda:0")
    d_q = torch.from_numpy(queries).to(dev)
    g_ids = torch.zeros((G, nq, k), dtype=torch.int64, device=dev)
    g_sc = torch.zeros((G, nq, k), dtype=torch.float64, device=dev)
    g_pos = torch.zeros((G, nq, k), dtype=torch.int64, device=dev)
    g_cnt = torch.zeros((G, nq), dtype=torch.int32, device=dev)
    g_flg = torch.zeros((G, nq), dtype=torch.int32, device=dev)
    shards = []
    per = n // G
    stream = torch.cuda.current_stream().cuda_stream or 1   # 1 == cudaStreamLegacy
    for metric in vl.SimilarityMetric:
        for g in range(G):
            s = vl.FlatIndex(dim)
            s.add_batch(np.arange(g * per, (g + 1) * per, dtype=np.uint64), rows[g * per:(g + 1) * per])
            s.set_pos_base(g * per)
            s.search_device(d_q.data_ptr(), nq, k, metric, g_ids[g].data_ptr(), g_sc[g].data_ptr(),
                            g_pos[g].data_ptr(), g_cnt[g].data_ptr(), g_flg[g].data_ptr(), stream)
            shards.append(s)
        o_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        o_sc = torch.zeros((nq, k), dtype=torch.float64, device=dev)
        o_pos = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        o_cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
        st = vl.lib().vl_merge_topk_device(0, G, nq, k, g_ids.data_ptr(), g_sc.data_ptr(), g_pos.data_ptr(),
                                           g_cnt.data_ptr(), o_ids.data_ptr(), o_sc.data_ptr(),
                                           o_pos.data_ptr(), o_cnt.data_ptr(), stream)
        assert st == 0
        torch.cuda.synchronize()
        assert int(g_flg.max()) == 0, "certificate failed on a shard"
        for qi in range(nq):
            st, oi, os_ = oracle_mod.flat_search(rows, None, queries[qi], k, int(metric))
            assert o_ids[qi].tolist() == list(map(int, oi)), (metric, qi)
            assert _hex(o_sc[qi].tolist()) == _hex(os_)
            assert o_pos[qi].tolist() == list(map(int, oi))
        shards.clear()


def test_peer_exchange_three_shards_one_gpu(vl, oracle_mod):
    """The peer-memory exchange protocol (csrc/exchange.cu) with three shards living in ONE process on
    one GPU (vl_exchange_connect_local): each shard's final-top-k kernel stores its block into every
    peer's slot and stamps it, each shard's merge kernel waits for the stamps.  Shards run on separate
    streams (a merge kernel spins until its peers have pushed).  Result on EVERY shard == the unsharded
    oracle, cross-shard ties included; repeated past the slot ring depth (slot reuse + acknowledgements),
    single-query, chunked small-batch and batched (tensor-core) paths, plain and PDL-pipelined handles."""
    import torch
    from vectorlite_b200.sharded import PeerExchange
    n, dim, k, G = 9000, 384, 10, 3
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    rows[3000:3005] = rows[10]
    rows[6000:6003] = rows[10]
    queries = np.concatenate([oracle_mod.synth_rows(43, 0, 39, dim), rows[10:11]])
    dev = torch.device("cuda:0")
    d_q = torch.from_numpy(queries).to(dev)
    per = n // G
    shards, xs, streams = [], [], []
    for g in range(G):
        s = vl.FlatIndex(dim)
        s.add_batch(np.arange(g * per, (g + 1) * per, dtype=np.uint64), rows[g * per:(g + 1) * per])
        s.set_pos_base(g * per)
        shards.append(s)
        xs.append(PeerExchange(0, G, g, max_nq=64, max_k=16))
        streams.append(torch.cuda.Stream(device=dev))
    PeerExchange.connect_local(xs)
    torch.cuda.synchronize()

    def outs(nq):
        return dict(o_ids=torch.zeros((nq, k), dtype=torch.int64, device=dev),
                    o_sc=torch.zeros((nq, k), dtype=torch.float64, device=dev),
                    o_pos=torch.zeros((nq, k), dtype=torch.int64, device=dev),
                    o_cnt=torch.zeros((nq,), dtype=torch.int32, device=dev),
                    xflg=torch.zeros((1, nq), dtype=torch.int32, device=dev))

    expected = {}

    def oracle_of(metric, lo, hi):
        key = (int(metric), lo, hi)
        if key not in expected:
            st, oi, os_ = oracle_mod.flat_search_batch(rows, None, queries[lo:hi], k, int(metric))
            assert st == 0
            expected[key] = (oi, os_)
        return expected[key]

    rounds = [(vl.SimilarityMetric.Cosine, 39, 40, False), (vl.SimilarityMetric.Cosine, 0, 1, False),
              (vl.SimilarityMetric.Euclidean, 0, 5, False), (vl.SimilarityMetric.Manhattan, 0, 40, False),
              (vl.SimilarityMetric.DotProduct, 0, 40, False)]
    rounds += [(vl.SimilarityMetric.Cosine, i, i + 1, True) for i in range(10)]   # overlapped, > ring depth
    rounds += [(vl.SimilarityMetric.Cosine, 0, 40, True), (vl.SimilarityMetric.Cosine, 3, 4, False)]
    # Warm-up without the exchange: scratch (re)allocations synchronise the whole device, which in this
    # one-process emulation would stall a shard behind a peer's spinning merge kernel (separate processes
    # on separate GPUs, the real deployment, do not share that dependency).
    for metric, lo, hi, _ in rounds:
        for g in range(G):
            o = outs(hi - lo)
            shards[g].search_device(d_q[lo:hi].data_ptr(), hi - lo, k, metric, o["o_ids"].data_ptr(),
                                    o["o_sc"].data_ptr(), o["o_pos"].data_ptr(), o["o_cnt"].data_ptr(),
                                    o["xflg"].data_ptr(), streams[g].cuda_stream)
    torch.cuda.synchronize()
    # One unchecked round through the exchange itself: first use loads the merge kernel and sizes the
    # exchange scratch, which may synchronise the device and (in this one-GPU emulation only) leave a
    # shard's merge spinning until its time-out.  Stamps only grow, so the checked rounds below start clean.
    for g in range(G):
        shards[g].set_pipelined(False)
        xs[g].search(shards[g], d_q[0:1], k, vl.SimilarityMetric.Cosine, outs(1), streams[g].cuda_stream)
    torch.cuda.synchronize()
    pending = []
    for metric, lo, hi, overlap in rounds:
        res = []
        for g in range(G):
            o = outs(hi - lo)
            shards[g].set_pipelined(overlap)
            xs[g].search(shards[g], d_q[lo:hi], k, metric, o, streams[g].cuda_stream)
            res.append(o)
        pending.append((metric, lo, hi, res))
        if not overlap or len(pending) == 4:
            torch.cuda.synchronize()
            for metric_, lo_, hi_, res_ in pending:
                oi, os_ = oracle_of(metric_, lo_, hi_)
                for g in range(G):
                    assert int(res_[g]["xflg"].max()) == 0, (metric_, lo_, g, res_[g]["xflg"])
                    assert np.array_equal(res_[g]["o_ids"].cpu().numpy().astype(np.uint64), oi), (metric_, lo_, g)
                    assert np.array_equal(res_[g]["o_sc"].cpu().numpy().view(np.uint64), os_.view(np.uint64))
                    assert np.array_equal(res_[g]["o_pos"].cpu().numpy().astype(np.uint64), oi)
                    assert int(res_[g]["o_cnt"].min()) == k
            pending = []
    torch.cuda.synchronize()
    # a peer that never shows up: bounded wait, FLAG_EXCHANGE (bit 4), no hang
    o = outs(1)
    xs[0].search(shards[0], d_q[0:1], k, vl.SimilarityMetric.Cosine, o, streams[0].cuda_stream)
    torch.cuda.synchronize()
    assert int(o["xflg"][0, 0]) & 16
    for x in xs:
        x.close()


def test_full_size_1m_properties_and_sampled_oracle(vl, oracle_mod):
    """BASELINE config 2 size (1M × 384): size-independent properties on the device-generated
    store + full-oracle check on sampled queries."""
    n, dim, k = 1_000_000, 384, 10
    idx = vl.FlatIndex(dim)
    idx.fill_synthetic(42, n)
    assert idx.len() == n
    # property: a stored row queried back is its own top-1 with cosine == 1 (to 1e-6), L2/L1 sim == 1
    probe_ids = [0, 1, 63, 64, 65, 12345, 999_999]
    probes = np.stack([oracle_mod.synth_rows(42, r, 1, dim)[0] for r in probe_ids])
    for metric in vl.SimilarityMetric:
        gi, gs, gc = idx.search_batch(probes, k, metric)
        assert [int(x) for x in gi[:, 0]] == probe_ids, metric
        assert np.all(np.diff(gs, axis=1) <= 0)                      # non-increasing scores
        if metric in (vl.SimilarityMetric.Euclidean, vl.SimilarityMetric.Manhattan):
            assert np.all(gs[:, 0] == 1.0)
        else:
            assert np.allclose(gs[:, 0], 1.0, atol=1e-6)
    # idempotence: same query twice → identical bits
    q = oracle_mod.synth_rows(43, 0, 2, dim)
    a = idx.search_batch(q, k, vl.SimilarityMetric.Cosine)
    b = idx.search_batch(q, k, vl.SimilarityMetric.Cosine)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # sampled full oracle (regenerates the 1M rows on the host with the same counter-based recipe)
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    for metric in vl.SimilarityMetric:
        _check(vl, oracle_mod, idx, rows, None, q, k, metric)
    assert idx.stats()["exact_queries"] == 0, "certificate should hold on i.i.d. data"
    # BASELINE config 2, B = 1024 at full size: the tcgen05 pipeline (all three threshold stages, coalesced
    # stage-0 stores, convergent survivor flush, streaming rescore) against the oracle on sampled queries
    # and against the independent single-query fp32 path
    bq = oracle_mod.synth_rows(43, 1000, 1024, dim)
    sample = [0, 1, 127, 128, 511, 1023]
    # (manhattan has no tensor-core form: the same staged pipeline on the CUDA-core tile kernel)
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.Euclidean, vl.SimilarityMetric.DotProduct,
                   vl.SimilarityMetric.Manhattan):
        before = idx.stats()
        bi, bs, bc = idx.search_batch(bq, k, metric)
        after = idx.stats()
        assert np.all(bc == k) and np.all(np.diff(bs, axis=1) <= 0)
        assert after["exact_queries"] == before["exact_queries"], "every certificate must hold on the batched path"
        st, oi, os_ = oracle_mod.flat_search_batch(rows, None, bq[sample], k, int(metric), nthreads=8)
        assert st == 0
        assert np.array_equal(bi[sample], oi), metric
        assert np.array_equal(bs[sample].view(np.uint64), os_.view(np.uint64)), metric
        for j in (5, 700):
            si, ss, _ = idx.search_batch(bq[j:j + 1], k, metric)
            assert np.array_equal(si[0], bi[j]) and np.array_equal(ss[0].view(np.uint64), bs[j].view(np.uint64))


@pytest.mark.parametrize("n,dim", [(5, 8), (3000, 96), (20000, 384), (70000, 100)])
def test_batched_pipeline_parity(vl, oracle_mod, n, dim):
    """nq >= 8 goes through the staged-threshold tile pipeline (CUDA-core kernel): ids exact, f64
    scores bit-identical to the oracle, all four metrics, stage boundaries (4096, 65536) crossed."""
    rng = np.random.default_rng(n + dim)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    if n > 100:
        rows[50] = rows[7]            # duplicate across tile boundaries → positional tie-break
        rows[n - 1] = rows[7]
        rows[9] = 0.0
    queries = rng.standard_normal((40, dim)).astype(np.float32)
    queries[3] = rows[min(7, n - 1)]
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    st, _, _ = oracle_mod.flat_search_batch(rows, None, queries[:1], 1, 0)
    for metric in vl.SimilarityMetric:
        for k in (10, 100):
            gi, gs, gc = idx.search_batch(queries, k, metric)
            st, oi, os_ = oracle_mod.flat_search_batch(rows, None, queries, k, int(metric), nthreads=8)
            assert st == 0
            kk = min(k, n)
            assert np.all(gc == kk)
            assert np.array_equal(gi[:, :kk], oi[:, :kk]), (metric, k)
            assert np.array_equal(gs[:, :kk].view(np.uint64), os_[:, :kk].view(np.uint64)), (metric, k)


def test_batched_pipeline_ties_fall_back(vl, oracle_mod):
    n, dim = 6000, 64
    rows = np.tile(np.linspace(0.1, 1.0, dim, dtype=np.float32), (n, 1))
    queries = np.tile(rows[:1], (16, 1))
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    gi, gs, gc = idx.search_batch(queries, 10, vl.SimilarityMetric.Cosine)
    assert np.all(gi == np.arange(10, dtype=np.uint64)[None, :])
    assert idx.stats()["exact_queries"] >= 16


def test_tensor_core_batched_path(vl, oracle_mod):
    """B=256 queries on unit-norm 384-d data: cosine / L2 / dot go through the tcgen05 kernel (bf16
    inputs, fp32 TMEM accumulators), candidates are re-scored in f64 and certified with the bf16
    error bound.  Results must equal the oracle bit for bit, and most queries must be certified on
    the tensor path (not silently re-run on the exact path)."""
    n, dim, nq, k = 50000, 384, 256, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    queries = oracle_mod.synth_rows(43, 0, nq, dim)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.DotProduct, vl.SimilarityMetric.Euclidean):
        before = idx.stats()
        gi, gs, gc = idx.search_batch(queries, k, metric)
        after = idx.stats()
        st, oi, os_ = oracle_mod.flat_search_batch(rows, None, queries, k, int(metric), nthreads=8)
        assert st == 0 and np.all(gc == k)
        assert np.array_equal(gi, oi), metric
        assert np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), metric
        certified = after["fast_queries"] - before["fast_queries"]
        assert certified >= 0.9 * nq, (metric, certified)
        assert after["tensor_queries"] - before["tensor_queries"] >= nq, metric
    # k = 100 and a ragged query count (not a multiple of the 128-query MMA tile)
    gi, gs, gc = idx.search_batch(queries[:77], 100, vl.SimilarityMetric.Cosine)
    st, oi, os_ = oracle_mod.flat_search_batch(rows, None, queries[:77], 100, 0, nthreads=8)
    assert np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64))
    # FP32 mode keeps the batched pipeline on CUDA cores and must agree
    idx.set_mode(vl.Mode.Fp32)
    gi2, gs2, _ = idx.search_batch(queries[:77], 100, vl.SimilarityMetric.Cosine)
    assert np.array_equal(gi2, oi) and np.array_equal(gs2.view(np.uint64), os_.view(np.uint64))


@pytest.mark.parametrize("dim", [768, 1000, 1536])
def test_tensor_core_batched_path_wide_rows(vl, oracle_mod, dim):
    """Rows wider than 384 elements (768 / 1024 / 1536-d embedding models; 1000 exercises the K padding): the query
    block no longer fits in shared memory beside the row ring, its K-chunks are streamed through the ring with the
    row chunks (batch_tc.cu, STREAM_A).  Same bar as 384-d: ids and f64 scores bit-identical to the oracle, the batch
    served by the tcgen05 kernel and certified there.  256 queries run as CTA pairs, 77 as single CTAs."""
    n, k = 30000, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    queries = oracle_mod.synth_rows(43, 0, 256, dim)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.DotProduct, vl.SimilarityMetric.Euclidean):
        for nq in (256, 77):
            before = idx.stats()
            gi, gs, gc = idx.search_batch(queries[:nq], k, metric)
            after = idx.stats()
            st, oi, os_ = oracle_mod.flat_search_batch(rows, None, queries[:nq], k, int(metric), nthreads=8)
            assert st == 0 and np.all(gc == k)
            assert np.array_equal(gi, oi), (metric, nq)
            assert np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), (metric, nq)
            assert after["tensor_queries"] - before["tensor_queries"] >= nq, (metric, nq)
            assert after["fast_queries"] - before["fast_queries"] >= 0.9 * nq, (metric, nq)
            assert after["exact_queries"] == before["exact_queries"], (metric, nq)


def test_bf16_mirror_single_query_scan(vl, oracle_mod):
    """AUTO mode at 384-d: single-query scans of all four metrics read the bf16 mirror of the rows (half the HBM
    bytes); the f64 rescore + bf16-bound certificate keep ids and scores bit-identical to the oracle.  A
    certificate that cannot hold under the bf16 bound is retried on the fp32 scan, not on the exact path."""
    n, dim, k = 30000, 384, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    rows[777] = rows[5]                                   # an exact duplicate: positional tie-break
    q = oracle_mod.synth_rows(43, 0, 6, dim)
    q[2] = rows[5]
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    for metric in vl.SimilarityMetric:
        before = idx.stats()
        for j in range(q.shape[0]):
            _check(vl, oracle_mod, idx, rows, None, q[j:j + 1], k, metric)
        after = idx.stats()
        used = after["bf16_scans"] - before["bf16_scans"]
        assert used >= q.shape[0], (metric, used)
        assert after["bf16_retries"] == before["bf16_retries"], (metric, "the bf16 bound should certify random rows")
        assert after["exact_queries"] == before["exact_queries"], metric
    # FP32 mode: same answers, the mirror is not touched
    idx.set_mode(vl.Mode.Fp32)
    b = idx.stats()["bf16_scans"]
    _check(vl, oracle_mod, idx, rows, None, q[:1], k, vl.SimilarityMetric.Cosine)
    assert idx.stats()["bf16_scans"] == b
    idx.set_mode(vl.Mode.Auto)
    # delete shifts positions: the mirror is positional and must be rebuilt
    idx.delete(3)
    rows2 = np.delete(rows, 3, axis=0)
    ids2 = np.delete(np.arange(n, dtype=np.uint64), 3)
    _check(vl, oracle_mod, idx, rows2, ids2, q[:2], k, vl.SimilarityMetric.Cosine)
    # near-ties below the bf16 bound: the base certificate fails, the query is re-run with a larger over-selection
    # (and on the fp32 arena if that does not certify either) — never on the exact path
    base = oracle_mod.synth_rows(44, 0, 1, dim)[0]
    rng = np.random.default_rng(5)
    near = (base[None, :] + 2e-3 * rng.standard_normal((200, dim))).astype(np.float32)   # cosines within ~3e-4
    far = oracle_mod.synth_rows(45, 0, 4000, dim)
    rows3 = np.concatenate([far, near])
    t = vl.FlatIndex(dim)
    t.add_batch(np.arange(rows3.shape[0], dtype=np.uint64), rows3)
    before = t.stats()
    _check(vl, oracle_mod, t, rows3, None, base[None, :], k, vl.SimilarityMetric.Cosine)
    after = t.stats()
    assert after["bf16_scans"] > before["bf16_scans"]
    assert after["bf16_retries"] > before["bf16_retries"]
    assert after["exact_queries"] == before["exact_queries"], "a larger over-selection / the fp32 scan should certify these near-ties"
    # the same clustered rows under L1 (the mirror's bound there is 2^-9·sqrt(dim)·max‖row‖ ≈ 0.04, absolute)
    _check(vl, oracle_mod, t, rows3, None, base[None, :], k, vl.SimilarityMetric.Manhattan)
    assert t.stats()["exact_queries"] == before["exact_queries"]


@pytest.mark.parametrize("dim", [128, 256, 768, 1024, 1536])
def test_bf16_mirror_scan_other_widths(vl, oracle_mod, dim):
    """The bf16-mirror single-query scan also serves 128- and 256-element rows (NCH = 1, 2) and the widths of larger
    embedding models (768 / 1024 / 1536 elements: NCH = 6 / 8 / 12, query still in registers)."""
    n, k = 12000, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    q = oracle_mod.synth_rows(43, 0, 3, dim)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    for metric in vl.SimilarityMetric:
        b = idx.stats()["bf16_scans"]
        for j in range(q.shape[0]):
            _check(vl, oracle_mod, idx, rows, None, q[j:j + 1], k, metric)
        assert idx.stats()["bf16_scans"] - b >= q.shape[0], (dim, metric)


def test_concurrent_single_query_callers_are_combined(vl, oracle_mod):
    """Many host threads calling search() with ONE query each on the same handle (the reference's serving pattern,
    client.rs:398 under a read lock): the handle combines what queues up behind a running launch into one batched
    search.  Every caller must still get exactly its own oracle answer; errors stay with their caller."""
    import threading
    n, dim, k, T, per = 60000, 384, 10, 8, 12      # (stores below 2^24 elements are served per caller, uncombined)
    rows = oracle_mod.synth_rows(42, 0, n, dim)
    q = oracle_mod.synth_rows(43, 0, T * per, dim)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    idx.search_batch(q[:1], k, vl.SimilarityMetric.Cosine)            # build the mirrors outside the race
    idx.search_batch(q[:4], k, vl.SimilarityMetric.Cosine)
    st, oi, os_ = oracle_mod.flat_search_batch(rows, None, q, k, 0, nthreads=8)
    assert st == 0
    got = {}
    errors = []

    def worker(t):
        try:
            for j in range(per):
                i = t * per + j
                metric = vl.SimilarityMetric.Cosine
                gi, gs, gc = idx.search_batch(q[i:i + 1], k, metric)
                got[i] = (gi[0].copy(), gs[0].copy(), int(gc[0]))
            if t == 0:                                                 # a bad query fails alone
                bad = q[0].copy(); bad[3] = np.nan
                try:
                    idx.search_batch(bad[None, :], k, vl.SimilarityMetric.Cosine)
                    errors.append("NaN query did not raise")
                except Exception:
                    pass
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    before = idx.stats()
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for i in range(T * per):
        gi, gs, gc = got[i]
        assert gc == k and np.array_equal(gi, oi[i]) and np.array_equal(gs.view(np.uint64), os_[i].view(np.uint64)), i
    after = idx.stats()
    print("combined:", after["combined_queries"] - before["combined_queries"], "of", T * per)
    assert after["combined_queries"] > before["combined_queries"], "8 concurrent callers should have been combined"


def test_cpp_host_mirror(vl, tmp_path):
    """include/vectorlite.hpp (the C++ mirror of the reference interface) replays the reference's own
    flat / hnsw unit tests against the C ABI."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_mirror_test")
    libdir = os.path.join(root, "vectorlite_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "host_mirror_test.cpp"), "-o", exe, "-L" + libdir,
                    "-lvectorlite_cuda", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "CPP_MIRROR PASS" in out.stdout, out.stdout + out.stderr



def test_adversarial_bf16_rounding_is_refused(vl, oracle_mod):
    """experiments/adversarial_bf16_rounding.py: every element of the query and of the true best row sits on a bf16
    rounding boundary, 70 bf16-exact rows score just below it.  With both operands rounded (tensor-core batch) a
    constant bound of 0.0040 certifies a top-k WITHOUT the best row; the measured-norm bound (and the worst-case
    constant 0.0079 it is capped by) must refuse that certificate, and the retry levels must return the oracle's
    answer.  The single-query mirror scan (one rounded operand) sees the same rows."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vl_adv", os.path.join(root, "experiments", "adversarial_bf16_rounding.py"))
    adv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(adv)
    rows, q, pos_a = adv.dataset()
    metric = vl.SimilarityMetric.DotProduct
    idx = vl.FlatIndex(adv.D)
    idx.add_batch(np.arange(rows.shape[0], dtype=np.uint64), rows)
    st, oi, os_ = oracle_mod.flat_search(rows, None, q, 10, int(metric))
    assert int(oi[0]) == pos_a
    before = idx.stats()
    gi, gs, gc = idx.search_batch(np.stack([q, q, q]), 10, metric)        # tensor-core batch
    after = idx.stats()
    for j in range(3):
        assert list(map(int, gi[j])) == list(map(int, oi)), j
        assert _hex(gs[j]) == _hex(os_)
    assert after["bf16_retries"] - before["bf16_retries"] >= 3, "the base certificate must NOT hold on this input"
    gi1, gs1, _ = idx.search_batch(q[None, :], 10, metric)                # single-query mirror scan
    assert list(map(int, gi1[0])) == list(map(int, oi)) and _hex(gs1[0]) == _hex(os_)


def test_clustered_rows_never_reach_the_exact_path(vl, oracle_mod):
    """1024-centre-mixture-like data (~940 rows per centre, as the bench's 1M x 1024 set): top-10 / top-64 cosine gaps
    (~0.003) are below the bf16 bound, so base certificates fail.  Failing queries must be answered by the larger
    over-selection (or the fp32 arena) — as ONE compacted batch, never by per-query exact scans — the handle must
    switch to the larger over-selection by itself, and every answer must equal the oracle's."""
    n, dim, k, clusters = 120_000, 384, 10, 128
    rows = oracle_mod.synth_rows(42, 0, n, dim, clusters)
    bq = oracle_mod.synth_rows(43, 0, 320, dim, clusters)
    idx = vl.FlatIndex(dim)
    idx.fill_synthetic(42, n, clusters=clusters)
    sample = [0, 1, 2, 3, 100, 200, 319]
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.DotProduct, vl.SimilarityMetric.Euclidean):
        st, oi, os_ = oracle_mod.flat_search_batch(rows, None, bq[sample], k, int(metric), nthreads=8)
        assert st == 0
        before = idx.stats()
        for rep in range(3):
            gi, gs, gc = idx.search_batch(bq, k, metric)
            assert np.all(gc == k)
            assert np.array_equal(gi[sample], oi), (metric, rep)
            assert np.array_equal(gs[sample].view(np.uint64), os_.view(np.uint64)), (metric, rep)
        for j in sample[:4]:                                               # lone queries: mirror scan + retries
            si, ss, _ = idx.search_batch(bq[j:j + 1], k, metric)
            assert np.array_equal(si[0], gi[j]) and np.array_equal(ss[0].view(np.uint64), gs[j].view(np.uint64))
        after = idx.stats()
        assert after["exact_queries"] == before["exact_queries"], (metric, "exact path reached", after)
    st = idx.stats()
    assert st["bf16_retries"] > 0, "this data is expected to defeat the base bf16 certificate"
    assert st["boosted_queries"] > 0, "the handle should have switched to the larger over-selection"


def test_device_search_repacks_unaligned_queries(vl, oracle_mod):
    """vl_index_search_device takes dense [nq][dim] queries; the kernels read at the arena pitch (dim rounded to 4).
    dim % 4 != 0 must not read query i > 0 from the wrong offset."""
    import torch
    n, dim, nq, k = 3000, 10, 5, 10
    rng = np.random.default_rng(7)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    idx = vl.FlatIndex(dim)
    idx.add_batch(np.arange(n, dtype=np.uint64), rows)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    o_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    o_sc = torch.zeros((nq, k), dtype=torch.float64, device=dev)
    o_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    o_flg = torch.zeros(nq, dtype=torch.int32, device=dev)
    for metric in vl.SimilarityMetric:
        for m in (1, nq):                         # single-query scans and (nq >= 2, cosine/dot/L2) the tensor path
            idx.search_device(d_q.data_ptr(), m, k, metric, o_ids.data_ptr(), o_sc.data_ptr(), 0, o_cnt.data_ptr(),
                              o_flg.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
            torch.cuda.synchronize()
            assert int((o_flg[:m] & 1).max()) == 0
            for j in range(m):
                st, oi, os_ = oracle_mod.flat_search(rows, None, q[j], k, int(metric))
                assert list(map(int, o_ids[j].cpu().numpy())) == list(map(int, oi)), (metric, m, j)
                assert _hex(o_sc[j].cpu().numpy()) == _hex(os_), (metric, m, j)
