"""Shard-aware indexes behind the same VectorIndex interface (src/lib.rs:224-245), ONE process driving several GPUs.

The reference is a single-process server whose `Collection` owns one index (`client.rs:243-247`); a CUDA-backed
collection on an 8-GPU box owns one of these instead, and nothing above the index changes:

* `MultiGpuFlatIndex` — the flat store ROW-SHARDED over the devices.  Shard g holds the contiguous
  storage-order range [base_g, base_g + n_g): every row of shard g precedes every row of shard g+1, so
  "(shard, local position) ascending" IS global insertion order and the stable-sort tie-break of `flat.rs:116`
  stays exact after the merge.  Routing: a bulk load into an empty index is split evenly; incremental adds go to
  the shard that owns the tail (the last non-empty shard, moving on when it reaches `shard_rows`); when the last
  shard is full the store is re-split evenly with 25 % headroom per shard (`rebalance`).  Deletes are
  order-preserving inside their shard (`flat.rs:93-96`, missing id = Ok).  A search runs on every shard
  concurrently (`vl_group_search`, csrc/group.cpp; each shard's answer is already exact and certified) and the
  per-shard top-k lists are merged on the host by a stable sort in shard order.  The multi-PROCESS deployment of the same
  layout, with the exchange fused into the kernels over NVLink peer memory, is `sharded.ShardedFlatIndex`.
* `MultiGpuHnswIndex` — HNSW does not shard (a graph traversal is sequential per query and partitioning the graph
  changes recall): one full REPLICA per device, mutations applied to all, queries split across the replicas.

Both are host-side routing over per-device `vl_index` handles of the C ABI; they add no CPU compute path.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np

import ctypes as C

from . import (VL_ERR_DIM, VL_OK, DimensionMismatch, FlatIndex, HNSWIndex, IndexType, SearchResult, SimilarityMetric,
               Vector, VectorLiteError, _err, _ptr, lib)

_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


class MultiGpuFlatIndex:
    """FlatIndex (src/index/flat.rs:59-136) row-sharded over `devices` (repeat a device id to place several shards
    on one GPU — that is how the single-GPU tests exercise the routing)."""

    MIN_SPLIT_ROWS = 4096     # a bulk load is not split below this many rows per shard

    def __init__(self, dim: int, devices: Sequence[int], shard_rows: int = 1 << 20,
                 data: Optional[Sequence[Vector]] = None, shard_factory=None):
        """`shard_factory(dim, device)` builds one shard (default: a `FlatIndex` on that device); the CPU tests of
        the routing bookkeeping pass a stand-in, searches always go through the real handles."""
        if not devices:
            raise ValueError("at least one device is required")
        self._dim = int(dim)
        self._devices = [int(d) for d in devices]
        self._make = shard_factory or (lambda dim_, dev: FlatIndex(dim_, device=dev))
        self._shards: List[FlatIndex] = [self._make(self._dim, d) for d in self._devices]
        self._shard_rows = max(1, int(shard_rows))
        self._where: Dict[int, int] = {}          # id -> shard
        self._meta: Dict[int, tuple] = {}         # id -> (text, metadata): gathered for the hits only
        self._tail = 0                            # shard that owns the tail of the storage order
        self._pool = ThreadPoolExecutor(max_workers=len(self._shards))
        self._L = None                            # the C ABI, loaded with the first search
        self._grp = None
        self._grp_lock = threading.Lock()
        if data:
            self.add_batch(np.array([v.id for v in data], dtype=np.uint64),
                           np.array([np.asarray(v.values, dtype=np.float32) for v in data], dtype=np.float32),
                           [v.text for v in data], [v.metadata for v in data])

    # -- lifecycle / shape -------------------------------------------------------------------------------------
    def close(self) -> None:
        self._drop_group()
        for s in self._shards:
            s.close()
        self._pool.shutdown(wait=False)

    def _group(self):
        """The vl_group over the current shard handles (re-created after a re-split replaces them)."""
        if self._grp is not None:
            return self._grp
        with self._grp_lock:                      # searches are re-entrant (read lock): create the group once
            if self._L is None:
                self._L = lib()
            if self._grp is None:
                arr = (C.c_void_p * len(self._shards))(*[s.handle for s in self._shards])
                g = C.c_void_p()
                st = self._L.vl_group_create(arr, len(self._shards), C.byref(g))
                if st != VL_OK:
                    raise VectorLiteError(st, _err())
                self._grp = g
        return self._grp

    def _drop_group(self) -> None:
        if self._grp is not None:
            self._L.vl_group_destroy(self._grp)
            self._grp = None

    def num_shards(self) -> int:
        return len(self._shards)

    def shard_sizes(self) -> List[int]:
        return [s.len() for s in self._shards]

    def shard_of(self, id: int) -> Optional[int]:
        return self._where.get(int(id))

    def dimension(self) -> int:
        return self._dim

    def len(self) -> int:
        return len(self._where)

    __len__ = len

    def is_empty(self) -> bool:
        return not self._where

    def max_id(self) -> Optional[int]:            # flat.rs:76-78
        return max(self._where) if self._where else None

    def metric(self) -> Optional[SimilarityMetric]:
        return None

    def index_type(self):
        return IndexType.Flat

    def _rebase(self) -> None:
        """global position of shard g's first row (device searches report base + local position)"""
        base = 0
        for s in self._shards:
            s.set_pos_base(base)
            base += s.len()

    # -- mutation ------------------------------------------------------------------------------------------------
    def _tail_shard(self) -> int:
        """The shard an appended row goes to; re-splits the store when the last shard is full."""
        g = self._tail
        while self._shards[g].len() >= self._shard_rows:
            if g + 1 < len(self._shards):
                g += 1
            else:
                self.rebalance()
                g = self._tail
        self._tail = g
        return g

    def add(self, vector: Vector) -> None:
        """flat.rs:82-91: dimension check, duplicate-id check over the WHOLE store, append."""
        vals = np.asarray(vector.values, dtype=np.float32)
        if vals.size != self._dim:
            raise ValueError("Vector dimension mismatch")                      # flat.rs:84
        if int(vector.id) in self._where:
            raise ValueError(f"Vector ID {int(vector.id)} already exists")    # flat.rs:87
        g = self._tail_shard()
        self._shards[g].add(Vector(id=int(vector.id), values=vals, text="", metadata=None))
        self._where[int(vector.id)] = g
        if vector.text or vector.metadata is not None:
            self._meta[int(vector.id)] = (vector.text, vector.metadata)
        self._rebase()

    def add_batch(self, ids, rows, texts=None, metadata=None, per: Optional[int] = None) -> None:
        """Bulk load (FlatIndex::new, flat.rs:68 / the persistence load path): an empty index is split evenly
        (`per` rows per shard when given), otherwise the rows fill the tail shard by shard."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self._dim:
            raise ValueError("Vector dimension mismatch")
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        n = rows.shape[0]
        assert ids.shape[0] == n
        id_list = [int(i) for i in ids]
        if len(set(id_list)) != n or any(i in self._where for i in id_list):
            dup = next(i for i in id_list if i in self._where or id_list.count(i) > 1)
            raise ValueError(f"Vector ID {dup} already exists")
        G = len(self._shards)
        if self.is_empty():
            if per is None:                                    # tiny loads stay together
                per = max(-(-n // G), min(self.MIN_SPLIT_ROWS, self._shard_rows)) if n else 0
            self._shard_rows = max(self._shard_rows, per)
            cuts = [min(n, g * per) for g in range(G + 1)]
        else:
            cuts = None
        if cuts is not None:
            futs = [self._pool.submit(self._shards[g].add_batch, ids[cuts[g]:cuts[g + 1]], rows[cuts[g]:cuts[g + 1]])
                    for g in range(G) if cuts[g + 1] > cuts[g]]
            for f in futs:
                f.result()
            for g in range(G):
                for i in id_list[cuts[g]:cuts[g + 1]]:
                    self._where[i] = g
            self._tail = max([g for g in range(G) if cuts[g + 1] > cuts[g]], default=0)
        else:
            done = 0
            while done < n:
                g = self._tail_shard()
                take = min(n - done, self._shard_rows - self._shards[g].len())
                self._shards[g].add_batch(ids[done:done + take], rows[done:done + take])
                for i in id_list[done:done + take]:
                    self._where[i] = g
                done += take
        if texts is not None or metadata is not None:
            for j, i in enumerate(id_list):
                self._meta[i] = (texts[j] if texts is not None else "", metadata[j] if metadata is not None else None)
        self._rebase()

    def delete(self, id: int) -> None:
        """flat.rs:93-96: `retain` — order-preserving, deleting a missing id is Ok."""
        g = self._where.pop(int(id), None)
        if g is None:
            return
        self._shards[g].delete(int(id))
        self._meta.pop(int(id), None)
        self._rebase()

    def rebalance(self, headroom: float = 0.25) -> None:
        """Re-split the store evenly over the shards, keeping the global storage order, and leave `headroom` of a
        shard's share free on each: appends only ever reach the tail shard, so the imbalance between the last
        shard and the others is bounded by `headroom` and a re-split (one export + one bulk upload) happens once per
        `headroom`·n/G appended rows."""
        ids, rows = self.export()
        per = -(-ids.shape[0] // len(self._devices))
        self._shard_rows = max(self._shard_rows, int(per * (1.0 + headroom)) + 1)
        self._drop_group()                       # it borrows the handles that are about to be replaced
        for s in self._shards:
            s.close()
        self._shards = [self._make(self._dim, d) for d in self._devices]
        self._where.clear()
        self._tail = 0
        meta, self._meta = self._meta, {}
        if ids.shape[0]:
            self.add_batch(ids, rows, per=per)
        self._meta = meta

    # -- queries -------------------------------------------------------------------------------------------------
    def search_batch(self, queries, k: int, similarity_metric, ef: int = 0):
        """(ids[nq,k] u64, scores[nq,k] f64, counts[nq] u32): every shard searched concurrently, lists merged by a
        stable sort in shard order (== score desc, global insertion order asc, flat.rs:116)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, qdim = q.shape
        if not self.is_empty() and qdim != self._dim:                          # flat.rs:99-104
            raise DimensionMismatch(self._dim, qdim, f"Dimension mismatch: expected {self._dim}, got {qdim}")
        k = max(int(k), 0)
        out_ids = np.full((nq, k), _NONE, dtype=np.uint64)
        out_sc = np.zeros((nq, k), dtype=np.float64)
        out_cnt = np.zeros(nq, dtype=np.uint32)
        if self.is_empty() or k == 0 or nq == 0:
            return out_ids, out_sc, out_cnt
        # vl_group_search: every non-empty shard searched concurrently, stable merge in shard order (csrc/group.cpp)
        grp = self._group()
        st = self._L.vl_group_search(grp, _ptr(q, C.c_float), nq, qdim, k, int(similarity_metric),
                                     _ptr(out_ids, C.c_uint64), _ptr(out_sc, C.c_double), _ptr(out_cnt, C.c_uint32))
        if st == VL_ERR_DIM:
            raise DimensionMismatch(self._dim, qdim, _err())
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return out_ids, out_sc, out_cnt

    def search(self, query, k: int, similarity_metric: SimilarityMetric, ef: int = 0) -> List[SearchResult]:
        ids, scores, counts = self.search_batch(np.asarray(query, dtype=np.float32)[None, :], k, similarity_metric, ef)
        return [self._result(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def _result(self, id_: int, score: float) -> SearchResult:
        text, md = self._meta.get(id_, ("", None))
        return SearchResult(id=id_, score=score, text=text, metadata=md)

    def get_vector(self, id: int) -> Optional[Vector]:
        g = self._where.get(int(id))
        if g is None:
            return None
        v = self._shards[g].get_vector(int(id))
        if v is None:
            return None
        text, md = self._meta.get(int(id), ("", None))
        return Vector(id=int(id), values=v.values, text=text, metadata=md)

    def export(self):
        """(ids, rows) of the whole store in global storage order."""
        parts = [s.export() for s in self._shards if s.len() > 0]
        if not parts:
            return np.empty(0, dtype=np.uint64), np.empty((0, self._dim), dtype=np.float32)
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def stats(self) -> List[dict]:
        return [s.stats() for s in self._shards]


class MultiGpuHnswIndex:
    """HNSWIndex (src/index/hnsw.rs:197-518) as one full replica per device: `add` / `delete` go to every replica,
    a batch of queries is split across the replicas, single queries rotate over them."""

    def __init__(self, dim: int, metric: SimilarityMetric, devices: Sequence[int], M: int = 16, M0: int = 32,
                 ef_construction: int = 400):
        if not devices:
            raise ValueError("at least one device is required")
        self._replicas: List[HNSWIndex] = [HNSWIndex(dim, metric, M=M, M0=M0, ef_construction=ef_construction, device=int(d))
                                           for d in devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self._replicas))
        self._rr = 0
        self._rr_lock = threading.Lock()

    @property
    def _meta(self):
        return self._replicas[0]._meta

    def close(self) -> None:
        for r in self._replicas:
            r.close()
        self._pool.shutdown(wait=False)

    def num_replicas(self) -> int:
        return len(self._replicas)

    def _all(self, fn_name: str, *args):
        futs = [self._pool.submit(getattr(r, fn_name), *args) for r in self._replicas]
        err = None
        for f in futs:
            try:
                f.result()
            except Exception as e:       # every replica sees the same inputs: the same error, raised once
                err = e
        if err is not None:
            raise err

    def add(self, vector: Vector) -> None:
        self._all("add", vector)

    def add_batch(self, ids, rows, texts=None, metadata=None) -> None:
        self._all("add_batch", ids, rows, texts, metadata)

    def build(self) -> None:
        self._all("build")

    def delete(self, id: int) -> None:
        self._all("delete", id)

    def _next(self) -> HNSWIndex:
        with self._rr_lock:
            r = self._replicas[self._rr % len(self._replicas)]
            self._rr += 1
        return r

    def search_batch(self, queries, k: int, similarity_metric, ef: int = 0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, R = q.shape[0], len(self._replicas)
        if nq < 2 * R or R == 1:
            return self._next().search_batch(q, k, similarity_metric, ef)
        cuts = [nq * g // R for g in range(R + 1)]
        parts = [f.result() for f in [self._pool.submit(self._replicas[g].search_batch, q[cuts[g]:cuts[g + 1]], k,
                                                        similarity_metric, ef) for g in range(R)]]
        return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
                np.concatenate([p[2] for p in parts]))

    def search(self, query, k: int, similarity_metric: SimilarityMetric, ef: int = 0) -> List[SearchResult]:
        return self._next().search(query, k, similarity_metric, ef)

    def _result(self, id_: int, score: float) -> SearchResult:
        return self._replicas[0]._result(id_, score)

    def get_vector(self, id: int) -> Optional[Vector]:
        return self._replicas[0].get_vector(id)

    def len(self) -> int:
        return self._replicas[0].len()

    __len__ = len

    def is_empty(self) -> bool:
        return self._replicas[0].is_empty()

    def dimension(self) -> int:
        return self._replicas[0].dimension()

    def max_id(self) -> Optional[int]:
        return self._replicas[0].max_id()

    def metric(self) -> Optional[SimilarityMetric]:
        return self._replicas[0].metric()

    def index_type(self):
        return IndexType.HNSW

    def export(self, *a, **kw):
        return self._replicas[0].export(*a, **kw)
