"""Builds a 200K x 384 GMM HNSW index and runs two 4096-query searches (ef=32) — used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl
n, nq, dim = int(os.environ.get("N", 200000)), int(os.environ.get("NQ", 4096)), 384
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qi = vl.FlatIndex(dim); qi.fill_synthetic(43, nq, clusters=1024)
ids, rows = flat.export(); q = qi.export()[1]
h = vl.HNSWIndex(dim, vl.SimilarityMetric.Cosine, ef_construction=100)
h.add_batch(ids, rows)
import time
truth, _, _ = flat.search_batch(q, 10, vl.SimilarityMetric.Cosine)
efs = [int(x) for x in os.environ.get("EF", "32").split(",")]
for ef in efs:
    h.search_batch(q, 10, vl.SimilarityMetric.Cosine, ef)
    t = time.time(); gi, gs, gc = h.search_batch(q, 10, vl.SimilarityMetric.Cosine, ef); dt = time.time() - t
    hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
    print(f"ef={ef} qps={nq/dt:.0f} recall={hit/(nq*10):.4f} visited/q={h.stats()['hnsw_visited']/nq:.0f}", flush=True)
print("ok", h.stats())
