"""Times a pipelined stream of single-query flat searches (device-resident queries) at a given row width.
Env: N (rows), DIM, METRICS (comma list of metric ids), K, REPS (queries per timed stream), VL_DISABLE_BF16_SCAN=1 for
the fp32-arena scan."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex
n = int(os.environ.get("N", 1_000_000)); dim = int(os.environ.get("DIM", 384)); k = int(os.environ.get("K", 10))
reps = int(os.environ.get("REPS", 256))
idx = ShardedFlatIndex(dim, rank=0, world=1, device=0)
idx.fill_synthetic(42, n)
idx.local.set_pipelined(True)
q = torch.from_numpy(oracle.synth_rows(43, 1000, reps, dim)).cuda()
out = {}
for m in [vl.SimilarityMetric(int(x)) for x in os.environ.get("METRICS", "0").split(",")]:
    for j in range(8):
        r = idx.search_device(q[j:j + 1], k, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(reps):
        idx.search_device(q[j:j + 1], k, m)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    bpe = 4 if os.environ.get("VL_DISABLE_BF16_SCAN") else 2
    out[m.name] = {"us_per_query": round(us, 2), "qps": round(1e6 / us), "GBps_at_%dB_per_element" % bpe: round(n * dim * bpe / us / 1e3, 1)}
print(json.dumps({"n": n, "dim": dim, "k": k, "bf16_scan": not os.environ.get("VL_DISABLE_BF16_SCAN"), **out, "stats": idx.local.stats()}))
