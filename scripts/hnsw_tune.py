"""Sweeps the HNSW search kernel's CTA width (VL_HNSW_WARPS) and visited-cache size (VL_HNSW_VIS_DIV) on one
device-built 1M x 384 graph: QPS (host API, 4096-query batches) and recall@10 per ef.  JSON lines."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorlite_b200 as vl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
efc = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dim, k, nq, clusters = 384, 10, 4096, 1024
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim); flat.fill_synthetic(42, n, clusters=clusters)
qidx = vl.FlatIndex(dim); qidx.fill_synthetic(43, nq, clusters=clusters)
import torch
queries = torch.from_numpy(qidx.export()[1]).pin_memory().numpy() if os.environ.get('PINNED', '1') == '1' else qidx.export()[1]
truth, _, _ = flat.search_batch(queries, k, metric)
ids, rows = flat.export()
flat.close()
h = vl.HNSWIndex(dim, metric, M=16, M0=32, ef_construction=efc)
h.add_batch(ids, rows); h.build()
print(json.dumps({"build": h.build_info()}), flush=True)
import itertools
for warps, vdiv, regpool in itertools.product(os.environ.get("WARPS", "4,2,1").split(","),
                                              os.environ.get("VDIVS", "1,2").split(","),
                                              os.environ.get("REGPOOLS", "1").split(",")):
    if True:
        os.environ["VL_HNSW_WARPS"] = warps
        os.environ["VL_HNSW_VIS_DIV"] = vdiv
        os.environ["VL_HNSW_REGPOOL"] = regpool
        row = {"warps": int(warps), "vis_div": int(vdiv), "regpool": int(regpool)}
        for ef in (0, 16, 32, 64, 128):
            h.search_batch(queries, k, metric, ef)
            t = time.perf_counter()
            for _ in range(3):
                gi, gs, gc = h.search_batch(queries, k, metric, ef)
            dt = (time.perf_counter() - t) / 3
            hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
            row[str(ef)] = {"qps": round(nq / dt), "recall": round(hit / (nq * k), 4),
                            "visited": round(h.stats()["hnsw_visited"] / nq)}
        print(json.dumps(row), flush=True)

if os.environ.get("REBUILD"):
    for regpool in os.environ["REBUILD"].split(","):
        os.environ["VL_HNSW_REGPOOL"] = regpool
        os.environ.pop("VL_HNSW_WARPS", None); os.environ.pop("VL_HNSW_VIS_DIV", None)
        h2 = vl.HNSWIndex(dim, metric, M=16, M0=32, ef_construction=efc)
        h2.add_batch(ids, rows)
        print(json.dumps({"rebuild_regpool": int(regpool), "build": h2.build_info(), "graph": h2.graph_check()}), flush=True)
        gi, gs, gc = h2.search_batch(queries, k, metric, 0)
        hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
        print(json.dumps({"rebuild_regpool": int(regpool), "recall_ef0": hit / (nq * k)}), flush=True)
        h2.close()
