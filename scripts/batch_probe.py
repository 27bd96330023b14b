"""One batched (B=1024) search per metric on a 1M x 384 synthetic flat index — used under ncu to get
the per-kernel breakdown of the batched pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex
n = int(os.environ.get("N", 1_000_000)); B = 1024; k = 10
idx = ShardedFlatIndex(384, rank=0, world=1, device=0)
idx.fill_synthetic(42, n)
q = torch.from_numpy(oracle.synth_rows(43, 1000, B, 384)).cuda()
metrics = [vl.SimilarityMetric(int(m)) for m in os.environ.get("METRICS", "0").split(",")]
for m in metrics:
    for _ in range(2):
        idx.search_device(q, k, m)
    torch.cuda.synchronize()
print("ok", idx.local.stats())
