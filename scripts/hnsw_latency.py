"""Single-query / small-batch HNSW latency through the host API (ef = k) on a device-built 1M x 384 graph, per CTA
width (VL_HNSW_WARPS is read at every launch).  One JSON line per width."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
widths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto", "4", "16", "1"]
dim, k = 384, 10
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qidx = vl.FlatIndex(dim); qidx.fill_synthetic(43, 256, clusters=1024)
queries = qidx.export()[1]
truth, _, _ = flat.search_batch(queries, k, metric)
ids, rows = flat.export(); flat.close()
h = vl.HNSWIndex(dim, metric, ef_construction=200)
h.add_batch(ids, rows); h.build()
configs = [(w, f, e) for w in widths for f in [int(x) for x in os.environ.get("BEAM_FACTORS", "8").split(",")]
           for e in os.environ.get("EXPANDS", "auto").split(",")]
for warps, factor, expand in configs:
    if warps == "auto":
        os.environ.pop("VL_HNSW_WARPS", None)
    else:
        os.environ["VL_HNSW_WARPS"] = warps
    if expand == "auto":
        os.environ.pop("VL_HNSW_EXPAND", None)
    else:
        os.environ["VL_HNSW_EXPAND"] = expand
    h.set_beam_factor(factor)
    out = {"warps": warps, "beam_factor": factor, "expand": expand}
    for nq in (1, 16, 64, 128):
        for i in range(10):
            h.search_batch(queries[:nq], k, metric, 0)
        t0 = time.perf_counter()
        for i in range(100):
            h.search_batch(queries[(i % 2) * 128:(i % 2) * 128 + nq], k, metric, 0)
        out[f"nq{nq}_us"] = round((time.perf_counter() - t0) / 100 * 1e6, 1)
    gi, _, gc = h.search_batch(queries, k, metric, 0)
    out["visited_per_query"] = h.stats()["hnsw_visited"] / 256
    out["recall_at_10"] = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(256)) / (256 * k)
    print(json.dumps(out), flush=True)
