"""Flat search on data whose top-k gaps are below the bf16 bound (1024-centre mixture) next to
i.i.d. rows: lone-query latency and B = 1024 batch time through the HOST API (vl_index_search: certificate levels
included), which level answered (stats deltas), and parity of sampled queries against the CPU oracle.
  python scripts/clustered_probe.py [N]  → JSON on the last line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, vectorlite_b200 as vl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim, k = 384, 10
out = {"n": n, "dim": dim, "k": k}
keys = ("fast_queries", "exact_queries", "bf16_scans", "bf16_retries", "fp32_retries", "boosted_queries")
for name, clusters in (("iid", 0), ("clustered_1024", 1024)):
    idx = vl.FlatIndex(dim)
    idx.fill_synthetic(42, n, clusters=clusters)
    q = oracle.synth_rows(43, 0, 1024, dim, clusters)
    res = {}
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.Euclidean, vl.SimilarityMetric.DotProduct,
                   vl.SimilarityMetric.Manhattan):
        r = {}
        idx.search_batch(q[:2], k, metric); idx.search_batch(q[:1], k, metric)     # mirrors built
        b = idx.stats()
        t = time.perf_counter()
        single = [idx.search_batch(q[j:j + 1], k, metric) for j in range(64)]
        r["single_us"] = (time.perf_counter() - t) / 64 * 1e6
        a = idx.stats(); r["single_stats"] = {x: a[x] - b[x] for x in keys}
        if metric != vl.SimilarityMetric.Manhattan or os.environ.get("L1_BATCH"):
            times = []
            for rep in range(4):
                b = idx.stats()
                t = time.perf_counter(); gi, gs, gc = idx.search_batch(q, k, metric); times.append((time.perf_counter() - t) * 1e3)
                a = idx.stats()
                r["batch_stats_rep%d" % rep] = {x: a[x] - b[x] for x in keys}
            r["batch_ms"] = [round(x, 3) for x in times]
            for j in range(64):
                assert np.array_equal(single[j][0][0], gi[j]) and np.array_equal(single[j][1][0].view(np.uint64), gs[j].view(np.uint64)), (name, metric, j)
        res[metric.name] = r
        print(name, metric.name, r, flush=True)
    # oracle parity of sampled queries (cosine + L2)
    rows = oracle.synth_rows(42, 0, n, dim, clusters)
    sample = [0, 1, 2, 3, 500, 1023]
    for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.Euclidean):
        gi, gs, gc = idx.search_batch(q, k, metric)
        st, oi, os_ = oracle.flat_search_batch(rows, None, q[sample], k, int(metric), nthreads=os.cpu_count() or 8)
        assert st == 0 and np.array_equal(gi[sample], oi) and np.array_equal(gs[sample].view(np.uint64), os_.view(np.uint64)), (name, metric)
    res["oracle_parity_sampled"] = "ok"
    out[name] = res
    del rows
    idx.close()
print(json.dumps(out))
