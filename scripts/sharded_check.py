"""torchrun -n G scripts/sharded_check.py — row-sharded flat search on G real GPUs vs the unsharded
oracle (ids + f64 scores bit-exact, all metrics, B=1 and batched), through ShardedFlatIndex."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import oracle, vectorlite_b200 as vl
from vectorlite_b200.sharded import ShardedFlatIndex

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, dim, k = 200_000, 384, 10
idx = ShardedFlatIndex(dim, rank=rank, world=world, device=lr)   # VL_EXCHANGE=p2p (default) | nccl
idx.fill_synthetic(42, n)
rows = oracle.synth_rows(42, 0, n, dim) if rank == 0 else None
queries = oracle.synth_rows(43, 0, 40, dim)
ok = True
for metric in vl.SimilarityMetric:
    for qs in (queries[:1], queries[:5], queries):          # B=1, small, batched (tensor / CUDA-core tiles)
        gi, gs, gc = idx.search(qs, k, metric)
        if rank == 0:
            st, oi, os_ = oracle.flat_search_batch(rows, None, qs, k, int(metric), nthreads=8)
            good = np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64))
            ok = ok and good
            print(f"metric={metric.name} nq={qs.shape[0]} {'OK' if good else 'MISMATCH'}", flush=True)
# pipelined exchange (side stream, ring of buffers) must give the same answers as the blocking one
d_q = torch.from_numpy(queries).cuda()
outs = []
for i in range(12):
    outs.append(idx.search_device_pipelined(d_q[i:i + 1], k, vl.SimilarityMetric.Cosine))
    if i % 4 == 3:            # consume a ring's worth before its slots are reused
        idx.drain(); torch.cuda.synchronize()
        for j, o in enumerate(outs):
            qi = i - len(outs) + 1 + j
            ref_ids, ref_sc, _, _ = idx.search_device(d_q[qi:qi + 1], k, vl.SimilarityMetric.Cosine)
            torch.cuda.synchronize()
            good = torch.equal(o[0], ref_ids) and torch.equal(o[1], ref_sc)
            ok = ok and good
        outs = []
if rank == 0:
    print("pipelined exchange", "OK" if ok else "MISMATCH", flush=True)
if rank == 0:
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", "world", world, "exchange", idx.exchange)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
