"""Single-process shard-aware Flat index over ALL visible GPUs (vectorlite_b200/multi_gpu.py): parity with the
oracle on the whole store and lone-caller timings next to one GPU holding everything.
  python scripts/multi_gpu_check.py [ROWS_TOTAL]   → JSON on the last line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
import vectorlite_b200 as vl
from vectorlite_b200.multi_gpu import MultiGpuFlatIndex

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
G = torch.cuda.device_count()
dim, k = 384, 10
rows = oracle.synth_rows(42, 0, n, dim)
rows[n - 7] = rows[3]                                    # a tie across the first and the last shard
q = oracle.synth_rows(43, 0, 64, dim)
q[1] = rows[3]
ids = np.arange(n, dtype=np.uint64)
idx = MultiGpuFlatIndex(dim, list(range(G)))
t0 = time.perf_counter(); idx.add_batch(ids, rows); load_s = time.perf_counter() - t0
one = vl.FlatIndex(dim, device=0); one.add_batch(ids, rows)
out = {"gpus": G, "rows_total": n, "shard_sizes": idx.shard_sizes(), "bulk_load_s": round(load_s, 3), "parity": {}, "us": {}}
for metric in vl.SimilarityMetric:
    PQ = 8 if n <= 500_000 else 2                        # the oracle takes ~0.5 s per query per million rows
    gi, gs, gc = idx.search_batch(q[:PQ], k, metric)
    ok = True
    for j in range(PQ):
        st, oi, os_ = oracle.flat_search(rows, ids, q[j], k, int(metric))
        ok &= list(map(int, gi[j])) == list(map(int, oi)) and [float(x).hex() for x in gs[j]] == [float(x).hex() for x in os_]
    out["parity"][metric.name] = bool(ok)
for name, ix in (("sharded", idx), ("one_gpu", one)):
    for nq in (1, 64):
        for _ in range(5):
            ix.search_batch(q[:nq], k, vl.SimilarityMetric.Cosine)
        t0 = time.perf_counter()
        for i in range(50):
            ix.search_batch(q[:nq], k, vl.SimilarityMetric.Cosine)
        out["us"][f"{name}_nq{nq}"] = round((time.perf_counter() - t0) / 50 * 1e6, 1)
print(json.dumps(out))
assert all(out["parity"].values()), out
