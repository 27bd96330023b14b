"""A handful of single-query HNSW searches (ef = k) on a device-built graph: the workload of
scripts/ncu_hnsw_single.sh (ncu capture of the nq = 1 search kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorlite_b200 as vl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim, k = 384, 10
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qidx = vl.FlatIndex(dim); qidx.fill_synthetic(43, 16, clusters=1024)
queries = qidx.export()[1]
ids, rows = flat.export(); flat.close()
h = vl.HNSWIndex(dim, metric, ef_construction=200)
h.add_batch(ids, rows); h.build()
for i in range(8):
    print(h.search_batch(queries[i:i + 1], k, metric, 0)[0][0][:3], h.stats()["hnsw_visited"], flush=True)
