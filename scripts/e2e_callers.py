"""e2e through vl_index_search with host buffers: N native caller threads, one query per call (the bench's e2e leg alone).
Env: N (rows), CLUSTERS (0 = i.i.d. rows, 1024 = the clustered mixture), CALLERS (comma list), TOTAL (queries per run), HNSW=1 for the HNSW index (ef = k)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, vectorlite_b200 as vl
n = int(os.environ.get("N", 1_000_000)); dim = 384; k = 10
cl = int(os.environ.get("CLUSTERS", 0))
idx = vl.FlatIndex(dim); idx.fill_synthetic(42, n, clusters=cl)
qi = vl.FlatIndex(dim); qi.fill_synthetic(43, 4096, first_row=0, clusters=cl)
q = np.ascontiguousarray(qi.export()[1], dtype=np.float32); qi.close()
idx.search_batch(q[:2], k, vl.SimilarityMetric.Cosine); idx.search_batch(q[:1], k, vl.SimilarityMetric.Cosine)
out = {}
for c in [int(x) for x in os.environ.get("CALLERS", "1,8,16,64").split(",")]:
    total = int(os.environ.get("TOTAL", 16384)) if c > 1 else 2048
    bench.native_callers(idx, q, k, vl.SimilarityMetric.Cosine, 0, c, total // 4)
    best = 0.0
    for _ in range(3):
        r = bench.native_callers(idx, q, k, vl.SimilarityMetric.Cosine, 0, c, total)
        best = max(best, r[0]) if r else best
    out[str(c)] = round(best)
print(json.dumps({"clusters": cl, "spin_us": os.environ.get("VL_COMBINE_SPIN_US", "default"), "qps_by_callers": out, "stats": idx.stats()}))
