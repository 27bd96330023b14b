"""Device HNSW build time + recall at ef = k on N x 384 clustered rows (efC = 400, M/M0 = 16/32).
Env: N, VL_HNSW_BUILD_ONE_WARP=1 (construction searches on the one-warp kernel without the rank merge)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl
n, nq, dim, k = int(os.environ.get("N", 1_000_000)), 4096, 384, 10
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qi = vl.FlatIndex(dim); qi.fill_synthetic(43, nq, clusters=1024)
ids, rows = flat.export(); q = qi.export()[1]
h = vl.HNSWIndex(dim, vl.SimilarityMetric.Cosine, ef_construction=400)
t0 = time.perf_counter(); h.add_batch(ids, rows); info = h.build(); dt = time.perf_counter() - t0
truth, _, _ = flat.search_batch(q, k, vl.SimilarityMetric.Cosine)
gi, gs, gc = h.search_batch(q, k, vl.SimilarityMetric.Cosine, 0)
hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
print(json.dumps({"n": n, "one_warp_build": bool(os.environ.get("VL_HNSW_BUILD_ONE_WARP")), "build_seconds": round(dt, 3),
                  "builder": info, "recall_at_10_ef_k": hit / (nq * k), "visited_per_query": h.stats()["hnsw_visited"] / nq}))
