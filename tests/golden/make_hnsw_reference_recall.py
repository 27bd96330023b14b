#!/usr/bin/env python3
"""Recall@10 of the ORACLE RESTATEMENT of the reference's HNSW (crate hnsw 0.11 semantics + u64
quantised functors, oracle/vl_oracle_hnsw.cpp) against exact flat results, on the bench's synthetic
data.  CPU-only and slow (single-threaded inserts at ef_construction = 400), so it is run once here
and the numbers are committed as tests/golden/hnsw_reference_recall.json; tests and bench.py compare
the CUDA HNSW against them at equal (M, M0, ef_construction, ef).

  python tests/golden/make_hnsw_reference_recall.py N CLUSTERS [M M0 EFC]

Build cost of the restated insert (one thread, as the reference's; the functors are evaluated by the
guard-banded vectorised form, identical results): clustered rows scale ~linearly (the LIFO search stays
inside a cluster), i.i.d. rows ~n^1.8 (it visits most of the graph) - 50K i.i.d. rows took 434 s with the
strict functors, so 1M i.i.d. rows are out of reach (about a day); fixtures: 1M clustered, 200K i.i.d.
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import oracle

n = int(sys.argv[1]); clusters = int(sys.argv[2])
M, M0, efc = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (16, 32, 400)
dim, k, nq, metric = 384, 10, int(os.environ.get("VL_FIXTURE_NQ", "1000")), 0
rows = oracle.synth_rows(42, 0, n, dim, clusters)
q = oracle.synth_rows(43, 0, nq, dim, clusters)
st, truth, _ = oracle.flat_search_batch(rows, None, q, k, metric, nthreads=4)
h = oracle.HNSW(dim, metric, M, M0, efc)
t = time.time(); h.add_batch(None, rows); build_s = time.time() - t
out = {"n": n, "dim": dim, "clusters": clusters, "M": M, "M0": M0, "ef_construction": efc, "metric": "cosine",
       "k": k, "nq": nq, "build_seconds_1thread": build_s, "strict_functor_evals": int(h.strict_evals()), "sweep": {}}
# structure of the restated graph's layer 0 (explains recall plateaus: closest-M0 lists without a diversity
# heuristic fall apart into per-cluster islands once a cluster holds more than M0 rows)
try:
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    z = h.export_zero(M0)
    src = np.repeat(np.arange(z.shape[0], dtype=np.int64), M0)
    dst = z.reshape(-1)
    keep = dst != np.uint64(0xFFFFFFFFFFFFFFFF)
    g = csr_matrix((np.ones(int(keep.sum()), dtype=np.int8), (src[keep], dst[keep].astype(np.int64))), shape=(n, n))
    nweak, lab = connected_components(g, directed=True, connection="weak")
    nstrong, labs = connected_components(g, directed=True, connection="strong")
    out["layer0"] = {"edges": int(keep.sum()), "weak_components": int(nweak),
                     "largest_weak_component": int(np.bincount(lab).max()),
                     "strong_components": int(nstrong), "largest_strong_component": int(np.bincount(labs).max()),
                     "upper_layers": int(h.num_layers()), "layer1_nodes": int(h.layer_len(1)) if h.num_layers() else 0}
    print(out["layer0"], flush=True)
except Exception as e:  # scipy is test-side only
    out["layer0"] = {"error": str(e)}
for ef in (0, 16, 32, 64, 128, 256, 512):
    t = time.time()
    st, ri, _, rc, vis = h.search_batch(q, k, ef, nthreads=4)
    dt = time.time() - t
    hit = sum(len(set(map(int, ri[i, :rc[i]])) & set(map(int, truth[i]))) for i in range(nq))
    out["sweep"][str(ef)] = {"recall_at_10": hit / (nq * k), "visited_per_query": vis / nq, "qps_4threads": nq / dt}
    print(ef, out["sweep"][str(ef)], flush=True)
path = os.path.join(os.path.dirname(__file__), f"hnsw_reference_recall_n{n}_c{clusters}_M{M}.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
