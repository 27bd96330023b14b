"""BASELINE config 5: flat index, ROWS x 384 row-sharded over G GPUs (12.5M rows per GPU at 100M/8),
batch 1024, k = 100, cosine, NCCL all-gather + merge.  torchrun -n G scripts/config5.py [ROWS_PER_GPU]
Checks: all certificates hold, stored rows query back to themselves, tensor-core batched path ==
fp32 single-query path on sampled queries.  Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import oracle, vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
n, dim, k, B = per * world, 384, 100, 1024
idx = ShardedFlatIndex(dim, rank=rank, world=world, device=lr)
t0 = time.time(); idx.fill_synthetic(42, n); torch.cuda.synchronize(); t_fill = time.time() - t0
q = oracle.synth_rows(43, 0, B, dim)
probe_rows = [0, per - 1, per, n // 2 + 17, n - 1][: (5 if world > 1 else 2)]
for i, r in enumerate(probe_rows):
    q[i] = oracle.synth_rows(42, r, 1, dim)[0]
d_q = torch.from_numpy(q).to(dev)
metric = vl.SimilarityMetric.Cosine
ids, sc, cnt, flg = idx.search_device(d_q, k, metric)          # warm-up: builds the bf16 mirror
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    ids, sc, cnt, flg = idx.search_device(d_q, k, metric)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
h_ids = ids.cpu().numpy().view(np.uint64); h_sc = sc.cpu().numpy()
failed = int((flg.cpu() & 1).sum())
ok_probe = all(int(h_ids[i, 0]) == r and abs(h_sc[i, 0] - 1.0) < 1e-6 for i, r in enumerate(probe_rows))
# independent path: fp32 single-query scans for 8 sampled queries
idx.local.set_mode(vl.Mode.Fp32)
ok_cross = True
for i in range(0, 8):
    i2, s2, c2, f2 = idx.search_device(d_q[i:i + 1], k, metric)
    torch.cuda.synchronize()
    ok_cross = ok_cross and np.array_equal(i2.cpu().numpy().view(np.uint64)[0], h_ids[i]) and np.array_equal(s2.cpu().numpy()[0], h_sc[i])
if rank == 0:
    flops = 2.0 * B * n * dim
    print(json.dumps({"config": "flat sharded cosine", "rows_total": n, "rows_per_gpu": per, "gpus": world, "batch": B, "k": k,
                      "ms_per_batch": ms, "qps": B / (ms * 1e-3), "tflops_aggregate": flops / (ms * 1e-3) / 1e12,
                      "tflops_per_gpu": flops / world / (ms * 1e-3) / 1e12, "cert_failed": failed, "probe_ok": ok_probe,
                      "tensor_vs_fp32_paths_equal": bool(ok_cross), "fill_seconds": t_fill}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
