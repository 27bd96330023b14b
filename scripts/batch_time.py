"""Times the batched (B=1024) flat pipeline on a 1M x 384 synthetic index: CUDA events over REPS batches.
Env: N (rows), DIM (row width, default 384), METRICS (comma list of metric ids), K, REPS, VL_TC_CLUSTER (multicast cluster size of the tensor-core kernel), VL_TC_PAIR (0 = no CTA pairs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex
n = int(os.environ.get("N", 1_000_000)); B = int(os.environ.get("B", 1024)); k = int(os.environ.get("K", 10))
reps = int(os.environ.get("REPS", 10)); dim = int(os.environ.get("DIM", 384))
idx = ShardedFlatIndex(dim, rank=0, world=1, device=0)
idx.fill_synthetic(42, n)
if os.environ.get("PIPELINED", "1") != "0":
    idx.local.set_pipelined(True)     # consecutive batches are chained with programmatic dependent launch
q = torch.from_numpy(oracle.synth_rows(43, 1000, B, dim)).cuda()
out = {}
for m in [vl.SimilarityMetric(int(x)) for x in os.environ.get("METRICS", "0").split(",")]:
    for _ in range(3):
        r = idx.search_device(q, k, m)
    torch.cuda.synchronize()
    failed = int((r[3] & 1).sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx.search_device(q, k, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[m.name] = {"ms_per_batch": round(ms, 4), "qps": round(B / ms * 1e3), "tflops": round(2.0 * B * n * dim / ms / 1e9, 1),
                   "cert_failed": failed}
print(json.dumps({"n": n, "dim": dim, "B": B, "k": k, "cluster": os.environ.get("VL_TC_CLUSTER", "1"), "pair": os.environ.get("VL_TC_PAIR", "default"), **out}))
