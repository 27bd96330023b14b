"""Small end-to-end pass over every kernel family (a target for compute-sanitizer where that tool is available):
flat single-query, batched CUDA-core and tensor-core pipelines (all metrics), exact path, HNSW device + host
build, HNSW search (1-warp and 4-warp CTAs), reference score mode.  Checks results against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, vectorlite_b200 as vl

n, dim, k = int(os.environ.get("N", 9000)), 384, 10
rows = oracle.synth_rows(42, 0, n, dim)
q = oracle.synth_rows(43, 0, 136, dim)
idx = vl.FlatIndex(dim)
idx.add_batch(np.arange(n, dtype=np.uint64), rows)
for metric in vl.SimilarityMetric:
    for qs in (q[:1], q[:3], q[:136]):
        gi, gs, gc = idx.search_batch(qs, k, metric)
        st, oi, os_ = oracle.flat_search_batch(rows, None, qs, k, int(metric), nthreads=4)
        assert np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), (metric, len(qs))
# round 2: threshold-estimation stage + tau kernel (needs >= 64K rows), certificate levels on clustered rows
# (larger over-selection, fp32 retry), chained device batches on a pipelined handle
n2 = int(os.environ.get("N2", 70000))
rows2 = oracle.synth_rows(42, 0, n2, dim, 64)
q2 = oracle.synth_rows(43, 0, 40, dim, 64)
big = vl.FlatIndex(dim)
big.fill_synthetic(42, n2, clusters=64)
for metric in (vl.SimilarityMetric.Cosine, vl.SimilarityMetric.Euclidean, vl.SimilarityMetric.DotProduct):
    for qs in (q2[:1], q2[:40]):
        for rep in range(2):
            gi, gs, gc = big.search_batch(qs, k, metric)
        st, oi, os_ = oracle.flat_search_batch(rows2, None, qs, k, int(metric), nthreads=4)
        assert np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), ("clustered", metric, len(qs))
import torch
big.set_pipelined(True)
dq = torch.from_numpy(q2).cuda()
o_ids = torch.zeros((40, k), dtype=torch.int64, device="cuda"); o_sc = torch.zeros((40, k), dtype=torch.float64, device="cuda")
o_cnt = torch.zeros(40, dtype=torch.int32, device="cuda"); o_flg = torch.zeros(40, dtype=torch.int32, device="cuda")
for rep in range(3):
    big.search_device(dq.data_ptr(), 40, k, vl.SimilarityMetric.Cosine, o_ids.data_ptr(), o_sc.data_ptr(), 0, o_cnt.data_ptr(),
                      o_flg.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
torch.cuda.synchronize()
print("flat round-2 paths ok", big.stats(), flush=True)
idx.set_mode(vl.Mode.Exact)
gi, gs, gc = idx.search_batch(q[:2], k, vl.SimilarityMetric.Cosine)
print("flat ok", idx.stats(), flush=True)
hn = int(os.environ.get("HN", 5000))
for builder in ("device", "host"):
    h = vl.HNSWIndex(dim, vl.SimilarityMetric.Cosine, ef_construction=48)
    h.set_builder(builder)
    h.add_batch(np.arange(hn, dtype=np.uint64), rows[:hn])
    assert h.graph_check()["invalid"] == 0
    for nq in (3, 1100):
        qq = np.concatenate([q] * 9)[:nq]
        gi, gs, gc = h.search_batch(qq, k, vl.SimilarityMetric.Cosine, 0)
        assert np.all(gc == k)
    h.set_score_mode("reference")
    h.search_batch(q[:2], k, vl.SimilarityMetric.Cosine, 16)
    print("hnsw ok", builder, h.build_info(), flush=True)
print("SANITIZE_PROBE_OK")
