"""Pins the CPU oracle (C++ restatement + pure-Python restatement) against every known-answer
test the reference's own tests hold for the search hot path."""
import math

import numpy as np
import pytest

from oracle import py_oracle as po


def _check_asserts(case, ids, scores):
    assert len(ids) == case["expect_len"]
    assert list(ids[:len(case["expect_ids_prefix"])]) == case["expect_ids_prefix"]
    for a in case["asserts"]:
        if "score" in a:
            assert abs(scores[a["index"]] - a["score"]) < a["tol"]
        if "score_gt" in a:
            assert scores[a["index"]] > a["score_gt"]
        if a.get("non_increasing"):
            assert all(scores[i - 1] >= scores[i] for i in range(1, len(scores)))


def test_flat_kats_cpp(oracle_mod, kats):
    for case in kats["flat"]:
        rows = np.array([r["values"] for r in case["rows"]], dtype=np.float64)
        ids = np.array([r["id"] for r in case["rows"]], dtype=np.uint64)
        st, oi, os_ = oracle_mod.flat_search(rows, ids, case["query"], case["k"], case["metric"])
        assert st == oracle_mod.OK, case["name"]
        _check_asserts(case, [int(x) for x in oi], list(os_))
        # bit-exact against the IEEE evaluation of the reference formulae
        assert [int(x) for x in oi] == case["exact_ids"], case["name"]
        assert [float(s).hex() for s in os_] == case["exact_scores_hex"], case["name"]


def test_flat_kats_python(kats):
    for case in kats["flat"]:
        idx = po.FlatIndex(len(case["query"]), [(r["id"], r["values"]) for r in case["rows"]])
        res = idx.search(case["query"], case["k"], case["metric"])
        _check_asserts(case, [r[0] for r in res], [r[1] for r in res])


def test_metric_kats(oracle_mod, kats):
    for m in kats["metrics"]:
        v = oracle_mod.metric(m["metric"], m["a"], m["b"])
        assert abs(v - m["expect"]) < m["tol"], m["cite"]
        assert float(v).hex() == m["exact_hex"], m["cite"]
        assert abs(po.calculate(m["metric"], m["a"], m["b"]) - m["expect"]) < m["tol"]
    # lib.rs:255-274: cosine of identical non-unit vectors is not exactly 1, dot is 5
    assert oracle_mod.metric(0, [1.0, 2.0], [1.0, 2.0]) == 0.9999999999999998
    assert oracle_mod.metric(3, [1.0, 2.0], [1.0, 2.0]) == 5.0


def test_convert_kats(oracle_mod, kats):
    for c in kats["convert"]:
        for fn in (oracle_mod.convert_distance_to_similarity, po.convert_distance_to_similarity):
            v = fn(c["d"], c["metric"])
            if c["tol"] == 0.0:
                assert v == c["expect"]
            else:
                assert abs(v - c["expect"]) < c["tol"]
    # hnsw.rs:933-953 monotonicity, 955-975 edge cases
    for m in range(4):
        prev = float("inf")
        for d in [0.0, 0.1, 0.5, 1.0, 10.0, 100.0, 1000.0]:
            s = oracle_mod.convert_distance_to_similarity(d, m)
            assert s <= prev
            prev = s
    for m in (1, 2):
        assert 0.9 < oracle_mod.convert_distance_to_similarity(0.0001, m) <= 1.0
        assert 0.0 < oracle_mod.convert_distance_to_similarity(100000.0, m) < 0.01


def test_cpp_matches_python_random(oracle_mod):
    rng = np.random.default_rng(7)
    for dim in (1, 3, 17, 384, 768, 1536):   # the oracle at the widths the wide-row GPU tests use
        rows = rng.standard_normal((40, dim)).astype(np.float32)
        rows[3] = 0.0               # zero row → cosine 0.0 branch
        rows[5] = rows[4]           # exact duplicate → tie resolved by position
        q = rng.standard_normal(dim).astype(np.float32)
        for m in range(4):
            st, oi, os_ = oracle_mod.flat_search(rows, None, q, 10, m)
            assert st == 0
            idx = po.FlatIndex(dim, [(i, [float(x) for x in rows[i]]) for i in range(40)])
            ref = idx.search([float(x) for x in q], 10, m)
            assert [int(x) for x in oi] == [r[0] for r in ref]
            assert [float(s).hex() for s in os_] == [float(r[1]).hex() for r in ref]
            for a, b in zip(rows, [q] * 3):
                assert oracle_mod.hnsw_distance(m, a, b) == po.hnsw_distance(
                    m, [float(x) for x in a], [float(x) for x in b])


def test_flat_edge_semantics(oracle_mod):
    rows = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.float64)
    # flat.rs:99-104: dimension mismatch only when non-empty
    st, _, _ = oracle_mod.flat_search(rows, None, [1.0, 0.0, 0.0], 1, 0)
    assert st == oracle_mod.ERR_DIM
    st, oi, _ = oracle_mod.flat_search(np.zeros((0, 2)), None, [1.0, 0.0, 0.0], 1, 0)
    assert st == oracle_mod.OK and len(oi) == 0
    # k > n → n results; k == 0 → empty
    st, oi, _ = oracle_mod.flat_search(rows, None, [1.0, 0.0], 10, 0)
    assert st == 0 and len(oi) == 2
    st, oi, _ = oracle_mod.flat_search(rows, None, [1.0, 0.0], 0, 0)
    assert st == 0 and len(oi) == 0
    # NaN → the reference panics (flat.rs:116) → ERR_NAN
    bad = rows.copy()
    bad[1, 0] = float("nan")
    st, _, _ = oracle_mod.flat_search(bad, None, [1.0, 0.0], 1, 3)
    assert st == oracle_mod.ERR_NAN
    # equal scores keep storage order (stable sort); sums start from +0.0 so 0*(-1) stays +0.0
    z = np.array([[-1.0, 0.0], [1.0, 0.0]], dtype=np.float64)
    st, oi, os_ = oracle_mod.flat_search(z, None, [0.0, 1.0], 2, 3)
    assert [int(x) for x in oi] == [0, 1] and list(os_) == [0.0, 0.0]


def test_hnsw_functor_kats(oracle_mod, kats):
    c = kats["hnsw"]["id_mapping"]
    for r, d in zip(c["rows"], c["quantised_distances"]):
        assert oracle_mod.hnsw_distance(c["metric"], r["values"], c["query"]) == d
    assert c["quantised_distances"] == [173, 1424, 1424, 911]
    # zero vector cosine → 1000 (hnsw.rs:139-141); dot clamp (hnsw.rs:172)
    assert oracle_mod.hnsw_distance(0, [0.0, 0.0], [1.0, 0.0]) == 1000
    assert oracle_mod.hnsw_distance(3, [2000.0], [1.0]) == 0
    assert oracle_mod.hnsw_distance(3, [-2000.0], [1.0]) == 2000


def test_hnsw_index_kats(oracle_mod, kats):
    c = kats["hnsw"]["id_mapping"]
    h = oracle_mod.HNSW(3, c["metric"])
    for r in c["rows"]:
        assert h.add(r["id"], r["values"]) == 0
    assert len(h) == 4
    st, ids, scores, _ = h.search(c["query"], c["k"], c["metric"])
    assert st == 0 and len(ids) >= 1 and int(ids[0]) == c["expect_first_id"]  # hnsw.rs:633
    assert scores[0] == c["scores_by_row"][0]
    assert all(scores[i - 1] >= scores[i] for i in range(1, len(scores)))
    # hnsw.rs:547-557 dim mismatch on add; 637-646 duplicate id; 649-662 delete semantics
    assert h.add(7, [1.0, 2.0]) == oracle_mod.ERR_DIM
    assert h.add(100, [4.0, 5.0, 6.0]) == oracle_mod.ERR_DUP_ID
    assert h.delete(999) == oracle_mod.ERR_NOT_FOUND
    assert h.delete(100) == 0 and len(h) == 3
    st, ids, _, _ = h.search(c["query"], 4, c["metric"])
    assert st == 0 and 100 not in [int(x) for x in ids]          # soft delete filters results
    # hnsw.rs:425-430 metric mismatch; 416-421 dimension mismatch (always checked)
    assert h.search(c["query"], 2, 0)[0] == oracle_mod.ERR_METRIC_MISMATCH
    assert h.search([1.0, 2.0], 2, c["metric"])[0] == oracle_mod.ERR_DIM
    # hnsw.rs:596-602 empty index; 776-805 k > n does not fail
    e = oracle_mod.HNSW(3, 1)
    st, ids, _, _ = e.search([1.0, 2.0, 3.0], 5, 1)
    assert st == 0 and len(ids) == 0
    st, ids, _, _ = h.search(c["query"], 50, c["metric"])
    assert st == 0 and 1 <= len(ids) <= 3


def test_hnsw_level_distribution(oracle_mod):
    # crate random_level(): P(level >= 1) = 1/M; deterministic (zero-seeded ChaCha12)
    for M in (8, 16, 32):
        lv = oracle_mod.hnsw_levels(M, 200000)
        frac = float((lv >= 1).mean())
        assert abs(frac - 1.0 / M) < 0.15 / M
        assert np.array_equal(lv[:1000], oracle_mod.hnsw_levels(M, 1000))


def test_chacha12_block0_matches_python():
    """ChaCha with 12 rounds, zero key/nonce/counter: check the C++ stream against an
    independent pure-Python implementation (guards the PRNG restatement)."""
    import oracle

    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF

    def qr(s, a, b, c, d):
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 16)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 12)
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 8)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 7)

    def block(ctr):
        inp = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + [0] * 8 + [ctr & 0xFFFFFFFF, ctr >> 32, 0, 0]
        s = list(inp)
        for _ in range(6):
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15)
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14)
        return [(a + b) & 0xFFFFFFFF for a, b in zip(s, inp)]

    words = block(0) + block(1)
    M = 16
    exp = []
    for i in range(16):
        u = (words[2 * i + 1] << 32) | words[2 * i]
        uniform = float(u) / 18446744073709551616.0
        exp.append(int(-math.log(uniform) * (1.0 / math.log(float(M)))))
    assert list(oracle.hnsw_levels(M, 16)) == exp


def test_hnsw_recall_sanity(oracle_mod):
    """The restated graph must behave like an ANN index: on 2000 clustered points the
    reference setting (ef = k) finds most true neighbours under the QUANTISED metric, and a
    larger ef does not hurt."""
    n, dim, k = 2000, 32, 10
    rows = oracle_mod.synth_rows(42, 0, n, dim, clusters=16)
    qs = oracle_mod.synth_rows(43, 0, 50, dim, clusters=16)
    h = oracle_mod.HNSW(dim, 1, M=16, M0=32, ef_construction=100)
    assert h.add_batch(None, rows) == 0
    assert h.layer_len(0) == n and h.num_layers() >= 1
    hit = {0: 0, 64: 0}
    for q in qs:
        st, gt, _ = oracle_mod.flat_search(rows, None, q, k, 1)
        for ef in hit:
            st, ids, sc, vis = h.search(q.astype(np.float64), k, 1, ef)
            assert st == 0 and vis > 0 and len(ids) == k
            hit[ef] += len(set(int(x) for x in ids) & set(int(x) for x in gt))
    assert hit[0] / (50 * k) > 0.5
    assert hit[64] >= hit[0] - 5


def test_synth_rows_properties(oracle_mod):
    a = oracle_mod.synth_rows(42, 0, 64, 384)
    b = oracle_mod.synth_rows(42, 32, 32, 384)
    assert np.array_equal(a[32:], b)                      # counter-based: any row regenerable
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=1e-6)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1 / math.sqrt(384)) < 0.005
    c = oracle_mod.synth_rows(43, 0, 64, 384)
    assert not np.array_equal(a, c)


def test_bench_reference_arm_helpers_run_on_cpu(oracle_mod):
    """bench.py's reference-arm legs (CPU only): the flat baseline with its clone / single-thread brackets and the
    restated-HNSW block, on a tiny store."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vl_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rows = bench.synth_host_rows(oracle_mod, 42, 3000, bench.DIM, 2)
    assert np.array_equal(rows, oracle_mod.synth_rows(42, 0, 3000, bench.DIM))
    q = oracle_mod.synth_rows(43, 0, 4, bench.DIM)
    qps, dt, ids = bench.cpu_flat_qps(oracle_mod, rows, q, 10, 0, 2)
    qps_c, _, ids_c = bench.cpu_flat_qps(oracle_mod, rows, q, 10, 0, 2, clone_bytes=16)
    assert qps > 0 and qps_c > 0 and np.array_equal(ids, ids_c) and ids.shape == (4, 10)
    h = bench.hnsw_reference_cpu(oracle_mod, 2, n=600, efc=40, clusters=16, nq=32)
    assert h["rows"] == 600 and set(h["sweep"]) == {"0", "64"}
    for rec in h["sweep"].values():
        assert 0.0 <= rec["recall_at_10"] <= 1.0 and rec["qps_1_thread"] > 0 and rec["qps_2_threads"] > 0
    assert h["sweep"]["64"]["recall_at_10"] >= h["sweep"]["0"]["recall_at_10"]


def test_hnsw_fast_functors_identical(oracle_mod):
    """The restated insert evaluates the reference's u64 functors (hnsw.rs:113-174) through a guard-banded
    vectorised form (oracle/vl_oracle_hnsw.cpp): whenever the pre-floor value is within 1e-6 of an integer the strict
    left-to-right functor decides.  Whole graphs — every adjacency list of every layer — and the search results must
    be identical to the strict evaluation (VLO_HNSW_STRICT=1), for all four metrics, f32- and f64-valued rows."""
    import os
    n, dim = 1500, 96
    rows = oracle_mod.synth_rows(42, 0, n, dim, 16)
    q = oracle_mod.synth_rows(43, 0, 20, dim, 16)
    rng = np.random.default_rng(3)
    rows64 = rows.astype(np.float64) + 1e-9 * rng.standard_normal(rows.shape)     # not f32-representable
    for metric in (0, 1, 2, 3):
        for data in (rows, rows64):
            graphs = []
            for strict in ("1", "0"):
                os.environ["VLO_HNSW_STRICT"] = strict
                try:
                    h = oracle_mod.HNSW(dim, metric, 8, 16, 60)
                finally:
                    os.environ.pop("VLO_HNSW_STRICT", None)
                if data.dtype == np.float32:
                    assert h.add_batch(None, data) == 0
                else:
                    for i in range(n):
                        assert h.add(i, data[i]) == 0
                layers = [h.export_layer(l, 8) for l in range(1, h.num_layers() + 1)]
                res = h.search_batch(q, 10, 24, nthreads=1)
                graphs.append((h.export_zero(16), layers, res[1], res[2], h.strict_evals()))
            a, b = graphs
            assert a[4] == 0 and np.array_equal(a[0], b[0]), (metric, data.dtype)
            assert len(a[1]) == len(b[1])
            for la, lb in zip(a[1], b[1]):
                assert all(np.array_equal(x, y) for x, y in zip(la, lb))
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3].view(np.uint64), b[3].view(np.uint64))
