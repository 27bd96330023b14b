"""Single-query HNSW latency through the host API (ef = k) on a device-built 1M x 384 graph, per CTA width."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim, k = 384, 10
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qidx = vl.FlatIndex(dim); qidx.fill_synthetic(43, 256, clusters=1024)
queries = qidx.export()[1]
ids, rows = flat.export(); flat.close()
h = vl.HNSWIndex(dim, metric, ef_construction=200)
h.add_batch(ids, rows); h.build()
for warps in ("4", "2", "1"):
    os.environ["VL_HNSW_WARPS"] = warps
    out = {"warps": int(warps)}
    for nq in (1, 16, 64, 128):
        for i in range(10):
            h.search_batch(queries[:nq], k, metric, 0)
        t0 = time.perf_counter()
        for i in range(100):
            h.search_batch(queries[:nq], k, metric, 0)
        out[f"nq{nq}_us"] = round((time.perf_counter() - t0) / 100 * 1e6, 1)
    print(json.dumps(out), flush=True)
