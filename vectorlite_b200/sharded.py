"""Row-sharded flat index: one process per GPU, contiguous row ranges.  The per-shard top-k lists are
exchanged over NVLink peer memory: the kernel that produces a shard's final top-k stores it directly
into every peer's exchange slot and a merge kernel waits on per-query stamps (``PeerExchange``,
csrc/exchange.cu) — no collective call on the data path.  ``exchange="nccl"`` keeps the baseline
(ONE packed all-gather over NCCL + merge kernel); torch.distributed is only the plumbing that
carries the 64-byte IPC handles at setup.

The reference has no multi-device path; this is the shard-aware routing the north
star adds.  Rows are independent and top-k is a decomposable reduction, so the only exchange is
nq·k·(8+8+8) bytes per rank.  Global storage position = shard base + local position keeps the
reference's stable-sort tie-break (flat.rs:116) exact across shards.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import FlatIndex, SimilarityMetric, VectorLiteError, lib, _err, VL_OK


def _current_stream() -> int:
    """torch's current stream as a cudaStream_t; the legacy default stream is handle 0 in torch,
    which the C ABI reads as "use the handle's own stream", so map it to cudaStreamLegacy (0x1)."""
    return torch.cuda.current_stream().cuda_stream or 1


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous row range [lo, hi) owned by `rank` (same rule on every rank)."""
    per, rem = divmod(n_total, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class PeerExchange:
    """One rank's end of the peer-memory exchange (vl_exchange, include/vectorlite_cuda.h)."""

    HANDLE_BYTES = 64

    def __init__(self, device: int, world: int, rank: int, max_nq: int = 1024, max_k: int = 128):
        self._L = lib()
        self._x = C.c_void_p()
        self.device, self.world, self.rank, self.max_nq, self.max_k = device, world, rank, max_nq, max_k
        st = self._L.vl_exchange_create(device, world, rank, max_nq, max_k, C.byref(self._x))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def local_handle(self) -> bytes:
        buf = C.create_string_buffer(self.HANDLE_BYTES)
        st = self._L.vl_exchange_local_handle(self._x, buf)
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return bytes(buf.raw)

    def connect(self, handles) -> None:
        """handles: the local_handle() of every rank, in rank order."""
        table = b"".join(handles)
        assert len(table) == self.world * self.HANDLE_BYTES
        st = self._L.vl_exchange_connect(self._x, C.c_char_p(table))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    @staticmethod
    def connect_local(exchanges) -> None:
        """Wire exchanges that live in this process (rank order)."""
        arr = (C.c_void_p * len(exchanges))(*[e._x for e in exchanges])
        st = lib().vl_exchange_connect_local(arr, len(exchanges))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def search(self, index: FlatIndex, d_queries: torch.Tensor, k: int, metric: SimilarityMetric, outs,
               stream: int) -> None:
        """Enqueue shard search + push + wait/merge on `stream`; `outs` holds the device output tensors."""
        nq = d_queries.shape[0]
        st = self._L.vl_index_search_exchange(index._h, self._x, d_queries.data_ptr(), nq, k, int(metric),
                                              outs["o_ids"].data_ptr(), outs["o_sc"].data_ptr(),
                                              outs["o_pos"].data_ptr(), outs["o_cnt"].data_ptr(),
                                              outs["xflg"].data_ptr(), C.c_void_p(stream))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def close(self) -> None:
        if self._x and self._x.value:
            self._L.vl_exchange_destroy(self._x)
            self._x = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedFlatIndex:
    def __init__(self, dim: int, rank: Optional[int] = None, world: Optional[int] = None,
                 device: Optional[int] = None, group=None, exchange: Optional[str] = None):
        self.group = group
        self.exchange = exchange or os.environ.get("VL_EXCHANGE", "p2p")   # "p2p" | "nccl"
        self._px = None
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.device = torch.cuda.current_device() if device is None else device
        self.dim = dim
        self.local = FlatIndex(dim, device=self.device)
        self.base = 0
        self._bufs = {}
        self._ring = {}          # (nq, k) -> pipelined buffer sets
        self._side = None        # side stream for the exchange (all-gather + merge)
        self._seq = 0
        self.last_retried = 0    # queries of the last search() that were re-run after a failed certificate

    def fill_synthetic(self, seed: int, n_total: int, clusters: int = 0):
        lo, hi = shard_range(n_total, self.world, self.rank)
        self.base = lo
        self.local.fill_synthetic(seed, hi - lo, first_row=lo, clusters=clusters, first_id=lo)
        self.local.set_pos_base(lo)

    def add_shard(self, ids, rows, base: int):
        self.base = base
        self.local.add_batch(ids, rows)
        self.local.set_pos_base(base)

    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            G = self.world
            blk = int(lib().vl_packed_result_bytes(nq, k))       # bytes per shard block (multiple of 8)
            packed = torch.zeros((G, blk // 8), dtype=torch.int64, device=dev)
            # merged outputs live in ONE blob [ids nk | scores nk | counts nq (i32) | flags nq (i32)] so that the
            # host path brings everything back with a single D2H copy
            nk = nq * k
            blob = torch.zeros(2 * nk + nq, dtype=torch.int64, device=dev)
            tail = blob[2 * nk:].view(torch.int32)
            self._bufs[key] = dict(
                packed=packed, blk=blk, blob=blob,
                o_ids=blob[:nk].view(nq, k),
                o_sc=blob[nk:2 * nk].view(torch.float64).view(nq, k),
                o_pos=torch.zeros((nq, k), dtype=torch.int64, device=dev),
                o_cnt=tail[:nq],
                xflg=tail[nq:2 * nq].view(1, nq),   # OR of the shards' flags (peer-memory merge)
                flg=torch.zeros((G, nq), dtype=torch.int32, device=dev),
                h_blob=torch.zeros(2 * nk + nq, dtype=torch.int64).pin_memory(),
                h_q=torch.zeros((nq, self.dim), dtype=torch.float32).pin_memory(),
                d_q=torch.zeros((nq, self.dim), dtype=torch.float32, device=dev),
            )
        return self._bufs[key]

    def _peer_exchange(self, nq: int, k: int):
        """The connected PeerExchange (collective on first use / when capacities grow), or None when the
        NCCL baseline is selected or peer mapping is unavailable on this box."""
        if self.exchange != "p2p" or self.world == 1:
            return None
        px = self._px
        if px is not None and nq <= px.max_nq and k <= px.max_k:
            return px
        if px is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            px.close()
            self._px = None
        ok, handle, err = True, b"", ""
        try:
            px = PeerExchange(self.device, self.world, self.rank, max(nq, 1024), max(k, 128))
            handle = px.local_handle()
        except VectorLiteError as e:   # noqa: PERF203
            ok, err, px = False, str(e), None
        table = [None] * self.world
        dist.all_gather_object(table, (ok, handle), group=self.group)
        if all(t[0] for t in table):
            try:
                px.connect([t[1] for t in table])
            except VectorLiteError as e:
                ok, err = False, str(e)
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)    # also the barrier before the first search
        if not all(oks):
            if px is not None:
                px.close()
            if self.rank == 0:
                warnings.warn(f"peer-memory exchange unavailable ({err or 'a peer failed'}); using the NCCL all-gather")
            self.exchange = "nccl"
            return None
        self._px = px
        return px

    def search_device(self, d_queries: torch.Tensor, k: int, metric: SimilarityMetric):
        """d_queries: [nq, dim] fp32 CUDA tensor (replicated on every rank).  Returns device tensors
        (ids, scores, counts, flags) holding the GLOBAL top-k on every rank; flags is [1, nq] (OR over
        shards, peer-memory exchange) or [G, nq] (NCCL baseline).  No host synchronisation."""
        nq = d_queries.shape[0]
        b = self._buffers(nq, k)
        r = self.rank
        stream = _current_stream()
        if self.world == 1:  # single shard: the local result is already the global one
            self.local.search_device(d_queries.data_ptr(), nq, k, metric, b["o_ids"].data_ptr(),
                                     b["o_sc"].data_ptr(), b["o_pos"].data_ptr(), b["o_cnt"].data_ptr(),
                                     b["xflg"].data_ptr(), stream)
            return b["o_ids"], b["o_sc"], b["o_cnt"], b["xflg"]
        px = self._peer_exchange(nq, k)
        if px is not None:
            px.search(self.local, d_queries, k, metric, b, stream)
            return b["o_ids"], b["o_sc"], b["o_cnt"], b["xflg"]
        packed = b["packed"]
        base = packed[r].data_ptr()
        nk8 = nq * k * 8
        self.local.search_device(d_queries.data_ptr(), nq, k, metric, base, base + nk8, base + 2 * nk8,
                                 base + 3 * nk8, base + 3 * nk8 + nq * 4, stream)
        dist.all_gather_into_tensor(packed.view(-1), packed[r].clone(), group=self.group)
        st = lib().vl_merge_topk_packed_device(self.device, self.world, nq, k, packed.data_ptr(),
                                               b["o_ids"].data_ptr(), b["o_sc"].data_ptr(), b["o_pos"].data_ptr(),
                                               b["o_cnt"].data_ptr(), C.c_void_p(stream))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        # flags of every shard: the last nq u32 of each block
        flg = packed.view(torch.int32)[:, (3 * nk8 + nq * 4) // 4:(3 * nk8 + nq * 8) // 4]
        return b["o_ids"], b["o_sc"], b["o_cnt"], flg

    # ---- throughput mode: the exchange of search i overlaps the scan of search i+1 -----------------------
    def search_device_pipelined(self, d_queries: torch.Tensor, k: int, metric: SimilarityMetric, depth: int = 4):
        """Like search_device, but the all-gather + merge run on a side stream so the next search's
        scan can start immediately.  Results land in a ring of `depth` buffer sets; the returned tensors
        are valid once `drain()` (or the returned event) has been waited on."""
        if self.world == 1:
            return self.search_device(d_queries, k, metric) + (None,)
        nq = d_queries.shape[0]
        key = (nq, k)
        if key not in self._ring:
            saved = self._bufs.pop(key, None)
            sets = []
            for _ in range(depth):
                self._bufs.pop(key, None)
                sets.append(self._buffers(nq, k))
            self._bufs.pop(key, None)
            if saved is not None:
                self._bufs[key] = saved
            self._ring[key] = dict(sets=sets, events=[None] * depth)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ring = self._ring[key]
        slot = self._seq % len(ring["sets"])
        self._seq += 1
        b = ring["sets"][slot]
        main = torch.cuda.current_stream()
        px = self._peer_exchange(nq, k)
        if px is not None:   # one stream: on a pipelined handle the merge rides the PDL chain, no side stream
            px.search(self.local, d_queries, k, metric, b, _current_stream())
            return b["o_ids"], b["o_sc"], b["o_cnt"], b["xflg"], None
        if ring["events"][slot] is not None:
            main.wait_event(ring["events"][slot])          # this buffer set's previous exchange is done
        r = self.rank
        packed = b["packed"]
        base = packed[r].data_ptr()
        nk8 = nq * k * 8
        self.local.search_device(d_queries.data_ptr(), nq, k, metric, base, base + nk8, base + 2 * nk8,
                                 base + 3 * nk8, base + 3 * nk8 + nq * 4, _current_stream())
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            dist.all_gather_into_tensor(packed.view(-1), packed[r].clone(), group=self.group)
            st = lib().vl_merge_topk_packed_device(self.device, self.world, nq, k, packed.data_ptr(),
                                                   b["o_ids"].data_ptr(), b["o_sc"].data_ptr(), b["o_pos"].data_ptr(),
                                                   b["o_cnt"].data_ptr(), C.c_void_p(self._side.cuda_stream))
            if st != VL_OK:
                raise VectorLiteError(st, _err())
            done = torch.cuda.Event()
            done.record(self._side)
        ring["events"][slot] = done
        flg = packed.view(torch.int32)[:, (3 * nk8 + nq * 4) // 4:(3 * nk8 + nq * 8) // 4]
        return b["o_ids"], b["o_sc"], b["o_cnt"], flg, done

    def drain(self):
        """Make the current stream wait for every exchange issued by search_device_pipelined."""
        if self._side is not None:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)

    def search(self, queries: np.ndarray, k: int, metric: SimilarityMetric):
        """Host in / host out (the e2e path): H2D of the queries, sharded search, D2H of the merged
        top-k.  Queries for which some shard's certificate failed (bit0 of the OR-ed flags, identical on every
        rank) are re-run — those queries only — through each shard's host search, which walks the larger
        over-selection / fp32 / exact levels by itself, and merged again.  A peer whose results did not arrive
        within the exchange's bounded wait (bit4) leaves the merged list incomplete: that raises."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        b = self._buffers(nq, k)
        b["h_q"].numpy()[:] = queries                      # pinned staging, reused across calls
        b["d_q"].copy_(b["h_q"], non_blocking=True)
        ids, sc, cnt, flg = self.search_device(b["d_q"], k, metric)
        nk = nq * k
        if flg.shape[0] == 1:                              # peer-memory exchange: flags sit in the blob
            b["h_blob"].copy_(b["blob"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hb = b["h_blob"].numpy()
            h_flg = hb[2 * nk:].view(np.int32)[nq:2 * nq].reshape(1, nq)
        else:                                              # NCCL baseline: per-shard flags [G, nq]
            b["h_blob"].copy_(b["blob"], non_blocking=True)
            h_flg = flg.cpu().numpy()
            hb = b["h_blob"].numpy()
        any_flg = np.bitwise_or.reduce(h_flg, axis=0)
        if int((any_flg & 16).max()) != 0:
            raise VectorLiteError(6, "row-sharded exchange: a peer shard's results did not arrive within the bounded "
                                     "wait (flag bit4); the merged top-k would be incomplete")
        out_ids = hb[:nk].view(np.uint64).reshape(nq, k).copy()
        out_sc = hb[nk:2 * nk].view(np.float64).reshape(nq, k).copy()
        out_cnt = hb[2 * nk:].view(np.uint32)[:nq].copy()
        failed = np.nonzero(any_flg & 1)[0]
        self.last_retried = int(failed.size)
        if failed.size:
            ri, rs, rc = self._search_retry(queries[failed], k, metric)
            out_ids[failed], out_sc[failed], out_cnt[failed] = ri, rs, rc
        return out_ids, out_sc, out_cnt

    def _search_retry(self, queries, k, metric):
        # every rank saw the same OR-ed flags → all ranks take this branch together, with the same queries
        gi, gs, gc = self.local.search_batch(queries, k, metric)  # host API: larger K' → fp32 → exact, failing queries only
        nq = queries.shape[0]
        if self.world == 1:
            return gi, gs, gc
        dev = torch.device("cuda", self.device)
        r = self.rank
        nk = nq * k
        blk_words = int(lib().vl_packed_result_bytes(nq, k)) // 8
        packed = torch.zeros((self.world, blk_words), dtype=torch.int64, device=dev)
        blk = torch.zeros(blk_words, dtype=torch.int64)
        blk[0:nk] = torch.from_numpy(gi.view(np.int64).reshape(-1))
        blk[nk:2 * nk] = torch.from_numpy(gs.view(np.int64).reshape(-1))
        blk[2 * nk:3 * nk] = torch.from_numpy(gi.astype(np.int64).reshape(-1))   # sharded stores: id == global position
        blk.view(torch.int32)[6 * nk:6 * nk + nq] = torch.from_numpy(gc.view(np.int32))
        mine = blk.to(dev)
        dist.all_gather_into_tensor(packed.view(-1), mine, group=self.group)
        o_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        o_sc = torch.zeros((nq, k), dtype=torch.float64, device=dev)
        o_pos = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        o_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
        st = lib().vl_merge_topk_packed_device(self.device, self.world, nq, k, packed.data_ptr(), o_ids.data_ptr(),
                                               o_sc.data_ptr(), o_pos.data_ptr(), o_cnt.data_ptr(),
                                               C.c_void_p(_current_stream()))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return (o_ids.cpu().numpy().view(np.uint64), o_sc.cpu().numpy(), o_cnt.cpu().numpy().view(np.uint32))
