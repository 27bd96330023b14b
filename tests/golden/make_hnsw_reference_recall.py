#!/usr/bin/env python3
"""Recall@10 of the ORACLE RESTATEMENT of the reference's HNSW (crate hnsw 0.11 semantics + u64
quantised functors, oracle/vl_oracle_hnsw.cpp) against exact flat results, on the bench's synthetic
data.  CPU-only and slow (single-threaded inserts at ef_construction = 400), so it is run once here
and the numbers are committed as tests/golden/hnsw_reference_recall.json; tests and bench.py compare
the CUDA HNSW against them at equal (M, M0, ef_construction, ef).

  python tests/golden/make_hnsw_reference_recall.py N CLUSTERS [M M0 EFC]
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import oracle

n = int(sys.argv[1]); clusters = int(sys.argv[2])
M, M0, efc = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (16, 32, 400)
dim, k, nq, metric = 384, 10, 500, 0
rows = oracle.synth_rows(42, 0, n, dim, clusters)
q = oracle.synth_rows(43, 0, nq, dim, clusters)
st, truth, _ = oracle.flat_search_batch(rows, None, q, k, metric, nthreads=4)
h = oracle.HNSW(dim, metric, M, M0, efc)
t = time.time(); h.add_batch(None, rows); build_s = time.time() - t
out = {"n": n, "dim": dim, "clusters": clusters, "M": M, "M0": M0, "ef_construction": efc, "metric": "cosine",
       "k": k, "nq": nq, "build_seconds_1thread": build_s, "sweep": {}}
for ef in (0, 16, 32, 64, 128, 256):
    t = time.time()
    st, ri, _, rc, vis = h.search_batch(q, k, ef, nthreads=4)
    dt = time.time() - t
    hit = sum(len(set(map(int, ri[i, :rc[i]])) & set(map(int, truth[i]))) for i in range(nq))
    out["sweep"][str(ef)] = {"recall_at_10": hit / (nq * k), "visited_per_query": vis / nq, "qps_4threads": nq / dt}
    print(ef, out["sweep"][str(ef)], flush=True)
path = os.path.join(os.path.dirname(__file__), f"hnsw_reference_recall_n{n}_c{clusters}_M{M}.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
