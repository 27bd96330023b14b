"""World-size-2 gloo test (CPU) of the row-sharded flat protocol used by vectorlite_b200/sharded.py:
contiguous shard ranges, per-shard top-k with GLOBAL positions, one all-gather, merge ordered by
(score desc, global position asc).  The per-shard search is played by the CPU oracle here (the CUDA
shard kernel is covered by tests/test_flat_gpu.py::test_search_device_and_sharded_merge and by
scripts/sharded_check.py under torchrun on real GPUs); what this pins is the exchange + merge logic
and that the result equals the unsharded reference search, ties across shards included."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_topk_reference(ids, scores, pos, counts, k):
    """NumPy statement of merge_topk_kernel (arena.cu): [G, nq, k] lists → global top-k."""
    G, nq, _ = ids.shape
    out_ids = np.full((nq, k), 2**64 - 1, dtype=np.uint64)
    out_sc = np.zeros((nq, k))
    out_cnt = np.zeros(nq, dtype=np.uint32)
    for q in range(nq):
        ent = [(-scores[g, q, i], pos[g, q, i], ids[g, q, i]) for g in range(G) for i in range(counts[g, q])]
        ent.sort(key=lambda t: (t[0], t[1]))
        m = min(k, len(ent))
        out_cnt[q] = m
        for i in range(m):
            out_ids[q, i], out_sc[q, i] = ent[i][2], -ent[i][0]
    return out_ids, out_sc, out_cnt


def _worker(rank, world, port, n, dim, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from vectorlite_b200.sharded import shard_range
    rows = oracle.synth_rows(42, 0, n, dim)
    rows[n // 2 + 3] = rows[5]            # the same vector in both shards → cross-shard tie
    queries = np.concatenate([oracle.synth_rows(43, 0, 5, dim), rows[5:6]])
    lo, hi = shard_range(n, world, rank)
    nq = queries.shape[0]
    ok = True
    for metric in range(4):
        st, li, ls = oracle.flat_search_batch(rows[lo:hi], np.arange(lo, hi, dtype=np.uint64), queries, k, metric)
        mine = torch.from_numpy(np.stack([li.astype(np.int64), ls.view(np.int64), li.astype(np.int64)]))  # ids, scores, pos
        cnt = torch.full((nq,), min(k, hi - lo), dtype=torch.int64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        gcnt = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(gathered, mine)
        dist.all_gather(gcnt, cnt)
        g = torch.stack(gathered).numpy()
        oi, os_, oc = merge_topk_reference(g[:, 0].astype(np.uint64), g[:, 1].view(np.float64), g[:, 2].astype(np.uint64),
                                           torch.stack(gcnt).numpy(), k)
        st, ri, rs = oracle.flat_search_batch(rows, None, queries, k, metric)
        ok = ok and np.array_equal(oi, ri) and np.array_equal(os_.view(np.uint64), rs.view(np.uint64))
    ret[rank] = ok
    dist.destroy_process_group()


def test_shard_ranges_cover_and_are_contiguous():
    from vectorlite_b200.sharded import shard_range
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_exchange_and_merge_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 4000, 48, 10, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
