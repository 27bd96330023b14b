"""HNSW build time, QPS, recall@10 vs exact flat and visited nodes per query over an ef sweep.
  python scripts/hnsw_bench.py N CLUSTERS [EFC M M0 NQ]  → JSON on the last line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl

n = int(sys.argv[1]); clusters = int(sys.argv[2])
efc = int(sys.argv[3]) if len(sys.argv) > 3 else 400
M = int(sys.argv[4]) if len(sys.argv) > 4 else 16
M0 = int(sys.argv[5]) if len(sys.argv) > 5 else 32
nq = int(sys.argv[6]) if len(sys.argv) > 6 else 4096
dim, k = 384, 10
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim)
flat.fill_synthetic(42, n, clusters=clusters)
qidx = vl.FlatIndex(dim)
qidx.fill_synthetic(43, nq, clusters=clusters)
queries = qidx.export()[1]
t = time.time(); truth, _, _ = flat.search_batch(queries, k, metric); t_flat = time.time() - t
ids, rows = flat.export()
h = vl.HNSWIndex(dim, metric, M=M, M0=M0, ef_construction=efc)
h.set_builder(os.environ.get("BUILDER", "auto"))
t = time.time(); h.add_batch(ids, rows); h.build(); build_s = time.time() - t
out = {"n": n, "clusters": clusters, "M": M, "M0": M0, "ef_construction": efc, "nq": nq, "build_seconds": build_s,
       "builder": h.build_info(), "graph": h.graph_check(), "build_threads": os.cpu_count(), "flat_exact_batch_seconds": t_flat, "sweep": {}}
print("built in", build_s, flush=True)
factors = [int(x) for x in os.environ.get("BEAM_FACTORS", "1,8").split(",")]
for f in factors:
    h.set_beam_factor(f)
    sweep = {}
    for ef in (0, 16, 32, 64, 128, 256):
        h.search_batch(queries[:256], k, metric, ef)
        t = time.time(); gi, gs, gc = h.search_batch(queries, k, metric, ef); dt = time.time() - t
        hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
        sweep[str(ef)] = {"recall_at_10": hit / (nq * k), "qps_e2e": nq / dt,
                          "visited_per_query": h.stats()["hnsw_visited"] / nq, "beam": f * (ef or k)}
        print("factor", f, "ef", ef, sweep[str(ef)], flush=True)
    out["sweep"]["factor_%d" % f] = sweep
out["flat_stats"] = flat.stats()
print(json.dumps(out))
