"""Small end-to-end pass over every kernel family (a target for compute-sanitizer where that tool is available):
flat single-query, batched CUDA-core and tensor-core pipelines (all metrics), exact path, HNSW device + host
build, HNSW search (1-warp and 4-warp CTAs), reference score mode.  Checks results against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, vectorlite_b200 as vl

n, dim, k = int(os.environ.get("N", 9000)), 384, 10
rows = oracle.synth_rows(42, 0, n, dim)
q = oracle.synth_rows(43, 0, 136, dim)
idx = vl.FlatIndex(dim)
idx.add_batch(np.arange(n, dtype=np.uint64), rows)
for metric in vl.SimilarityMetric:
    for qs in (q[:1], q[:3], q[:136]):
        gi, gs, gc = idx.search_batch(qs, k, metric)
        st, oi, os_ = oracle.flat_search_batch(rows, None, qs, k, int(metric), nthreads=4)
        assert np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), (metric, len(qs))
idx.set_mode(vl.Mode.Exact)
gi, gs, gc = idx.search_batch(q[:2], k, vl.SimilarityMetric.Cosine)
print("flat ok", idx.stats(), flush=True)
hn = int(os.environ.get("HN", 5000))
for builder in ("device", "host"):
    h = vl.HNSWIndex(dim, vl.SimilarityMetric.Cosine, ef_construction=48)
    h.set_builder(builder)
    h.add_batch(np.arange(hn, dtype=np.uint64), rows[:hn])
    assert h.graph_check()["invalid"] == 0
    for nq in (3, 1100):
        qq = np.concatenate([q] * 9)[:nq]
        gi, gs, gc = h.search_batch(qq, k, vl.SimilarityMetric.Cosine, 0)
        assert np.all(gc == k)
    h.set_score_mode("reference")
    h.search_batch(q[:2], k, vl.SimilarityMetric.Cosine, 16)
    print("hnsw ok", builder, h.build_info(), flush=True)
print("SANITIZE_PROBE_OK")
