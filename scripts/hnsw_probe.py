"""Builds a 200K x 384 GMM HNSW index and runs two 4096-query searches (ef=32) — used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl
n, nq, dim = int(os.environ.get("N", 200000)), 4096, 384
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=1024)
qi = vl.FlatIndex(dim); qi.fill_synthetic(43, nq, clusters=1024)
ids, rows = flat.export(); q = qi.export()[1]
h = vl.HNSWIndex(dim, vl.SimilarityMetric.Cosine, ef_construction=100)
h.add_batch(ids, rows)
for _ in range(2):
    h.search_batch(q, 10, vl.SimilarityMetric.Cosine, int(os.environ.get("EF", 32)))
print("ok", h.stats())
