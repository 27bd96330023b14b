import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex
import bench
dev = torch.device("cuda", 0)
idx = ShardedFlatIndex(384, rank=0, world=1, device=0)
idx.fill_synthetic(42, 1_000_000)
idx.local.set_pipelined(True)
queries = bench.device_synth_rows(vl, 43, 0, 1024, 0)
d_queries = torch.from_numpy(queries).to(dev)
k = 10
mode = sys.argv[1] if len(sys.argv) > 1 else "b1first"
def per_rep(m2, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        idx.search_device(d_bq, k, m2)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(reps)]
bq = bench.device_synth_rows(vl, 43, 1000, 1024, 0)
d_bq = torch.from_numpy(bq).to(dev)
if mode == "b1first":
    for name, mid in bench.METRIC_NAMES.items():
        m2 = vl.SimilarityMetric(mid)
        for rep in range(2):
            for qi in range(1024):
                idx.search_device(d_queries[qi:qi + 1], k, m2)
        torch.cuda.synchronize()
for name in ("cosine", "euclidean", "dot", "cosine"):
    m2 = vl.SimilarityMetric(bench.METRIC_NAMES[name])
    r = idx.search_device(d_bq, k, m2); torch.cuda.synchronize()
    failed = int((r[3] & 1).sum().item())
    print(mode, name, failed, per_rep(m2), flush=True)
