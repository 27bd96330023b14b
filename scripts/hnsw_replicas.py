"""BASELINE config 4: HNSW memory-optimized (M/M0 = 8/16) and high-accuracy (32/64) profiles, 1M x 384,
cosine, k = 10, one full replica per GPU (replicas only: graph traversal does not shard), every
rank serving its own 4096-query batches.  torchrun -n G scripts/hnsw_replicas.py [ROWS] [EFC]
Rank 0 prints one JSON line: per-profile build seconds (device builder), recall@10 vs exact flat and the
aggregate q/s (sum over ranks of batch / max-over-ranks time)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import vectorlite_b200 as vl

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
efc = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dim, k, nq, clusters = 384, 10, 4096, 1024
metric = vl.SimilarityMetric.Cosine
flat = vl.FlatIndex(dim, device=lr); flat.fill_synthetic(42, n, clusters=clusters)
qsrc = vl.FlatIndex(dim, device=lr); qsrc.fill_synthetic(43, nq, first_row=rank * nq, clusters=clusters)   # each replica its own queries
queries = torch.from_numpy(qsrc.export()[1]).pin_memory().numpy()
truth, _, _ = flat.search_batch(queries, k, metric)
ids, rows = flat.export(); flat.close()
out = {"rows": n, "gpus": world, "batch_per_gpu": nq, "k": k, "ef_construction": efc, "profiles": {}}
for name in ("memory-optimized", "default", "high-accuracy"):
    h = vl.HNSWIndex(dim, metric, ef_construction=efc, device=lr, profile=name)
    h.add_batch(ids, rows); h.build()
    info = h.build_info()
    prof = {"M_M0": list(vl.HNSWIndex.PROFILES[name]), "build_seconds": info["seconds"], "builder": info["builder"], "sweep": {}}
    for ef in (0, 32, 128):
        h.search_batch(queries, k, metric, ef)
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); reps = 5
        for _ in range(reps):
            gi, gs, gc = h.search_batch(queries, k, metric, ef)
        dt = torch.tensor([(time.perf_counter() - t0) / reps], device=f"cuda:{lr}", dtype=torch.float64)
        hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
        rec = torch.tensor([hit / (nq * k)], device=f"cuda:{lr}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(rec, op=dist.ReduceOp.MIN)
        prof["sweep"][str(ef)] = {"qps_aggregate_e2e": world * nq / float(dt.item()), "recall_at_10_min_over_ranks": float(rec.item()),
                                  "visited_per_query": h.stats()["hnsw_visited"] / nq}
    out["profiles"][name] = prof
    h.close()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
