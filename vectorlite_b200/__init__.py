"""vectorlite_b200 — host-side mirror of VectorLite's index interface over the B200 C ABI.

The reference is Rust (no toolchain in this image), so the host side that tests and the bench
drive is this thin ctypes layer over ``libvectorlite_cuda.so`` (``include/vectorlite_cuda.h``).
It mirrors, name for name, the reference's vector-level API:

* ``SimilarityMetric``            — src/lib.rs:363-378
* ``Vector`` / ``SearchResult``   — src/lib.rs:163-174, 193-203
* ``VectorIndex`` protocol: ``add, delete, search, len, is_empty, get_vector, dimension``
                                  — src/lib.rs:224-245
* ``FlatIndex``                   — src/index/flat.rs:59-136
* ``HNSWIndex``                   — src/index/hnsw.rs:197-518
* ``VectorIndexWrapper``          — src/lib.rs:270-346 (``metric()``, ``index_type()``)

Text and metadata stay on the host (as ``HNSWIndex.metadata`` does at hnsw.rs:78-82,210) and are
attached to the <= k hits only — the reference clones them for all n rows (flat.rs:111-112).

There is NO CPU fallback: if the CUDA library is missing or no device is usable, construction
raises.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvectorlite_cuda.so")

__all__ = ["SimilarityMetric", "IndexType", "Vector", "SearchResult", "FlatIndex", "HNSWIndex",
           "VectorIndexWrapper", "VectorLiteError", "DimensionMismatch", "MetricMismatch", "lib",
           "Mode"]


class SimilarityMetric(enum.IntEnum):  # declaration order of src/lib.rs:364-378
    Cosine = 0
    Euclidean = 1
    Manhattan = 2
    DotProduct = 3

    @classmethod
    def default(cls):  # #[default] Cosine, lib.rs:367
        return cls.Cosine

    def calculate(self, a, b):
        raise NotImplementedError("scalar metrics run on the device inside search(); "
                                  "there is no host implementation in the product")


class IndexType(enum.IntEnum):  # src/lib.rs:248-268
    Flat = 0
    HNSW = 1


class Mode(enum.IntEnum):
    Auto = 0
    Exact = 1
    Fp32 = 2


VL_OK, VL_ERR_DIM, VL_ERR_DUP_ID, VL_ERR_NOT_FOUND, VL_ERR_METRIC_MISMATCH = 0, 1, 2, 3, 4
VL_ERR_INVALID, VL_ERR_CUDA, VL_ERR_OOM, VL_ERR_NAN, VL_ERR_UNSUPPORTED = 5, 6, 7, 8, 9


class VectorLiteError(Exception):  # src/errors.rs:10-67 (the variants the hot path raises)
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class DimensionMismatch(VectorLiteError):  # errors.rs DimensionMismatch{expected, actual}
    def __init__(self, expected: int, actual: int, message: str = ""):
        super().__init__(VL_ERR_DIM, message or f"Dimension mismatch: expected {expected}, got {actual}")
        self.expected, self.actual = expected, actual


class MetricMismatch(VectorLiteError):  # errors.rs MetricMismatch{requested, index}
    def __init__(self, requested, index, message: str = ""):
        super().__init__(VL_ERR_METRIC_MISMATCH, message or f"Metric mismatch: requested {requested}, index {index}")
        self.requested, self.index = requested, index


@dataclass
class Vector:  # src/lib.rs:163-174
    id: int
    values: Sequence[float]
    text: str = ""
    metadata: Optional[Any] = None


@dataclass
class SearchResult:  # src/lib.rs:193-203
    id: int
    score: float
    text: str = ""
    metadata: Optional[Any] = None


# ---------------------------------------------------------------------------------------------
_lib = None


def lib():
    """The loaded C ABI.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(
            f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  vectorlite_b200 has no CPU fallback.")
    L = C.CDLL(_SO)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    fp, dp, u64p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    sig = {
        "vl_flat_create": (i32, [u32, i32, C.POINTER(vp)]),
        "vl_hnsw_create": (i32, [u32, i32, u32, u32, u32, i32, C.POINTER(vp)]),
        "vl_index_destroy": (None, [vp]),
        "vl_index_add": (i32, [vp, u64, fp, u32]),
        "vl_index_add_f64": (i32, [vp, u64, dp, u32]),
        "vl_index_add_batch": (i32, [vp, u64p, fp, u64]),
        "vl_index_delete": (i32, [vp, u64]),
        "vl_index_fill_synthetic": (i32, [vp, u64, u64, u64, u32, u64]),
        "vl_index_build": (i32, [vp]),
        "vl_hnsw_set_builder": (i32, [vp, i32]),
        "vl_hnsw_build_info": (i32, [vp, u64p, u64p]),
        "vl_hnsw_set_score_mode": (i32, [vp, i32]),
        "vl_hnsw_set_beam_factor": (i32, [vp, u32]),
        "vl_hnsw_graph_bytes": (i32, [vp, u64p]),
        "vl_hnsw_export_graph": (i32, [vp, vp, u64, u64p]),
        "vl_hnsw_import_graph": (i32, [vp, u64p, fp, u64, vp, u64]),
        "vl_hnsw_graph_check": (i32, [vp, u64p]),
        "vl_index_search": (i32, [vp, fp, u32, u32, u32, i32, u32, u64p, dp, u32p]),
        "vl_index_search_f64": (i32, [vp, dp, u32, u32, u32, i32, u32, u64p, dp, u32p]),
        "vl_index_search_device": (i32, [vp, vp, u32, u32, i32, u32, vp, vp, vp, vp, vp, vp]),
        "vl_merge_topk_device": (i32, [i32, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "vl_packed_result_bytes": (u64, [u32, u32]),
        "vl_merge_topk_packed_device": (i32, [i32, u32, u32, u32, vp, vp, vp, vp, vp, vp]),
        "vl_exchange_create": (i32, [i32, u32, u32, u32, u32, C.POINTER(vp)]),
        "vl_exchange_destroy": (None, [vp]),
        "vl_exchange_local_handle": (i32, [vp, vp]),
        "vl_exchange_connect": (i32, [vp, vp]),
        "vl_exchange_connect_local": (i32, [C.POINTER(vp), u32]),
        "vl_index_search_exchange": (i32, [vp, vp, vp, u32, u32, i32, vp, vp, vp, vp, vp, vp]),
        "vl_group_create": (i32, [C.POINTER(vp), u32, C.POINTER(vp)]),
        "vl_group_destroy": (None, [vp]),
        "vl_group_size": (u32, [vp]),
        "vl_group_search": (i32, [vp, fp, u32, u32, u32, i32, u64p, dp, u32p]),
        "vl_index_len": (u64, [vp]),
        "vl_index_dim": (u32, [vp]),
        "vl_index_type_of": (i32, [vp]),
        "vl_index_metric": (i32, [vp]),
        "vl_index_device": (i32, [vp]),
        "vl_index_max_id": (i32, [vp, u64p]),
        "vl_index_get_vector": (i32, [vp, u64, fp]),
        "vl_index_export": (i32, [vp, u64, u64, u64p, fp, u64p]),
        "vl_index_set_mode": (i32, [vp, i32]),
        "vl_index_set_pos_base": (i32, [vp, u64]),
        "vl_index_stats": (i32, [vp, u64p, u32]),
        "vl_index_set_pipelined": (i32, [vp, i32]),
        "vl_index_set_profiling": (i32, [vp, i32]),
        "vl_index_profile_read": (i32, [vp, dp, u64p]),
        "vl_index_device_rows": (i32, [vp, C.POINTER(vp), u32p]),
        "vl_last_error": (C.c_char_p, []),
        "vl_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError == the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    L._vl_signatures = sig
    _lib = L
    return L


def _err() -> str:
    return (lib().vl_last_error() or b"").decode("utf-8", "replace")


def _ptr(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class _CudaIndex:
    """Shared VectorIndex plumbing (src/lib.rs:224-245) over one vl_index handle."""

    def __init__(self):
        self._h = C.c_void_p()
        self._L = lib()
        self._meta = {}  # id -> (text, metadata): host-side, gathered for the hits only

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.vl_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- VectorIndex -------------------------------------------------------------------------
    def add(self, vector: Vector) -> None:
        """Err(String) → ValueError with the reference's message (flat.rs:84,87; hnsw.rs:365,369)."""
        vals = np.ascontiguousarray(vector.values, dtype=np.float32)
        st = self._L.vl_index_add(self._h, int(vector.id), _ptr(vals, C.c_float), vals.size)
        if st != VL_OK:
            if st in (VL_ERR_DIM, VL_ERR_DUP_ID):
                raise ValueError(_err())
            raise VectorLiteError(st, _err())
        if vector.text or vector.metadata is not None:
            self._meta[int(vector.id)] = (vector.text, vector.metadata)

    def add_batch(self, ids, rows, texts=None, metadata=None) -> None:
        """Bulk load == FlatIndex::new(dim, data) (flat.rs:68) / the persistence load path."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dimension():
            raise ValueError("Vector dimension mismatch")
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        assert ids.shape[0] == rows.shape[0]
        st = self._L.vl_index_add_batch(self._h, _ptr(ids, C.c_uint64), _ptr(rows, C.c_float), rows.shape[0])
        if st != VL_OK:
            if st in (VL_ERR_DIM, VL_ERR_DUP_ID):
                raise ValueError(_err())
            raise VectorLiteError(st, _err())
        if texts is not None or metadata is not None:
            for i, id_ in enumerate(ids):
                self._meta[int(id_)] = (texts[i] if texts is not None else "",
                                        metadata[i] if metadata is not None else None)

    def delete(self, id: int) -> None:
        st = self._L.vl_index_delete(self._h, int(id))
        if st == VL_ERR_NOT_FOUND:
            raise ValueError(_err())  # hnsw.rs:402 "Vector ID {} does not exist"
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        self._meta.pop(int(id), None)

    def search(self, query, k: int, similarity_metric: SimilarityMetric, ef: int = 0) -> List[SearchResult]:
        ids, scores, counts = self.search_batch(np.asarray(query, dtype=np.float32)[None, :], k,
                                                similarity_metric, ef)
        return [self._result(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def search_batch(self, queries, k: int, similarity_metric, ef: int = 0):
        """Batched search (additive to the reference API): returns (ids[nq,k] u64, scores[nq,k] f64,
        counts[nq] u32)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, qdim = q.shape
        ids = np.empty((nq, max(k, 0)), dtype=np.uint64)
        scores = np.empty((nq, max(k, 0)), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.uint32)
        st = self._L.vl_index_search(self._h, _ptr(q, C.c_float), nq, qdim, k, int(similarity_metric), ef,
                                     _ptr(ids, C.c_uint64), _ptr(scores, C.c_double), _ptr(counts, C.c_uint32))
        if st == VL_ERR_DIM:
            raise DimensionMismatch(self.dimension(), qdim, _err())
        if st == VL_ERR_METRIC_MISMATCH:
            raise MetricMismatch(SimilarityMetric(int(similarity_metric)), self.metric(), _err())
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return ids, scores, counts

    def _result(self, id_: int, score: float) -> SearchResult:
        text, md = self._meta.get(id_, ("", None))
        return SearchResult(id=id_, score=score, text=text, metadata=md)

    def len(self) -> int:
        return int(self._L.vl_index_len(self._h))

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def get_vector(self, id: int) -> Optional[Vector]:
        out = np.empty(self.dimension(), dtype=np.float32)
        st = self._L.vl_index_get_vector(self._h, int(id), _ptr(out, C.c_float))
        if st == VL_ERR_NOT_FOUND:
            return None
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        text, md = self._meta.get(int(id), ("", None))
        return Vector(id=int(id), values=out, text=text, metadata=md)

    def dimension(self) -> int:
        return int(self._L.vl_index_dim(self._h))

    def max_id(self) -> Optional[int]:  # flat.rs:76-78, hnsw.rs:267-269
        out = C.c_uint64(0)
        st = self._L.vl_index_max_id(self._h, C.byref(out))
        return None if st != VL_OK else int(out.value)

    def metric(self) -> Optional[SimilarityMetric]:
        m = self._L.vl_index_metric(self._h)
        return None if m < 0 else SimilarityMetric(m)

    # -- extras ------------------------------------------------------------------------------
    def export(self, first: int = 0, count: Optional[int] = None):
        n = self.len() - first if count is None else count
        n = max(n, 0)
        ids = np.empty(n, dtype=np.uint64)
        rows = np.empty((n, self.dimension()), dtype=np.float32)
        got = C.c_uint64(0)
        st = self._L.vl_index_export(self._h, first, n, _ptr(ids, C.c_uint64), _ptr(rows, C.c_float), C.byref(got))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return ids[:got.value], rows[:got.value]

    def stats(self) -> dict:
        out = np.zeros(12, dtype=np.uint64)
        self._L.vl_index_stats(self._h, _ptr(out, C.c_uint64), 12)
        keys = ["launches", "fast_queries", "exact_queries", "h2d_bytes", "d2h_bytes", "hnsw_visited", "bf16_scans",
                "combined_queries", "bf16_retries", "fp32_retries", "boosted_queries", "tensor_queries"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def set_mode(self, mode: Mode) -> None:
        st = self._L.vl_index_set_mode(self._h, int(mode))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def set_pipelined(self, on: bool) -> None:
        self._L.vl_index_set_pipelined(self._h, 1 if on else 0)

    def set_profiling(self, on: bool) -> None:
        self._L.vl_index_set_profiling(self._h, 1 if on else 0)

    def profile_read(self):
        ms, n = C.c_double(0), C.c_uint64(0)
        st = self._L.vl_index_profile_read(self._h, C.byref(ms), C.byref(n))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return ms.value, int(n.value)

    @property
    def handle(self):
        return self._h


class FlatIndex(_CudaIndex):
    """FlatIndex (src/index/flat.rs:59-136) backed by the device arena."""

    def __init__(self, dim: int, data: Optional[Sequence[Vector]] = None, device: int = 0):
        super().__init__()
        st = self._L.vl_flat_create(int(dim), int(device), C.byref(self._h))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        if data:
            ids = np.array([v.id for v in data], dtype=np.uint64)
            rows = np.array([np.asarray(v.values, dtype=np.float32) for v in data], dtype=np.float32)
            self.add_batch(ids, rows, [v.text for v in data], [v.metadata for v in data])

    def fill_synthetic(self, seed: int, n: int, first_row: int = 0, clusters: int = 0, first_id: Optional[int] = None):
        st = self._L.vl_index_fill_synthetic(self._h, seed, first_row, n, clusters,
                                             first_row if first_id is None else first_id)
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def set_pos_base(self, base: int):
        self._L.vl_index_set_pos_base(self._h, int(base))

    def search_device(self, d_queries_ptr: int, nq: int, k: int, metric, d_ids: int, d_scores: int,
                      d_pos: int, d_counts: int, d_flags: int, stream: int = 0):
        """Enqueue a search whose queries and outputs are device pointers (no host sync)."""
        st = self._L.vl_index_search_device(self._h, d_queries_ptr, nq, k, int(metric), 0, d_ids, d_scores,
                                            d_pos or None, d_counts, d_flags, stream or None)
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def index_type(self):
        return IndexType.Flat


class HNSWIndex(_CudaIndex):
    """HNSWIndex (src/index/hnsw.rs:197-518).  The reference's compile-time profiles
    (hnsw.rs:95-109) are runtime parameters: M/M0 = 16/32 (default), 8/16 (memory-optimized),
    32/64 (high-accuracy)."""

    PROFILES = {"default": (16, 32), "memory-optimized": (8, 16), "high-accuracy": (32, 64)}

    def __init__(self, dim: int, metric: SimilarityMetric, M: int = 0, M0: int = 0, ef_construction: int = 0,
                 device: int = 0, profile: Optional[str] = None):
        super().__init__()
        if dim == 0:
            raise ValueError("HNSW index dimension cannot be 0")  # hnsw.rs:217-219 panics
        if profile:
            M, M0 = self.PROFILES[profile]
        st = self._L.vl_hnsw_create(int(dim), int(metric), M, M0, ef_construction, int(device), C.byref(self._h))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def build(self):
        st = self._L.vl_index_build(self._h)
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    BUILDERS = {"auto": 0, "host": 1, "device": 2}

    def set_builder(self, builder: str) -> None:
        """Where bulk adds into an empty index build the graph: "auto" (device from 4096 rows), "host", "device"."""
        st = self._L.vl_hnsw_set_builder(self._h, self.BUILDERS[builder])
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def export_graph(self) -> bytes:
        """Levels + adjacency of the current graph (vl_hnsw_export_graph); raises VectorLiteError (UNSUPPORTED) when
        the graph holds soft-deleted nodes."""
        nb = C.c_uint64(0)
        st = self._L.vl_hnsw_graph_bytes(self._h, C.byref(nb))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        buf = np.empty(nb.value, dtype=np.uint8)
        w = C.c_uint64(0)
        st = self._L.vl_hnsw_export_graph(self._h, buf.ctypes.data_as(C.c_void_p), nb.value, C.byref(w))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return buf[:w.value].tobytes()

    def import_graph(self, ids, rows, blob: bytes) -> None:
        """Restore a saved graph into this EMPTY index over the rows it was exported with (same order)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        b = np.frombuffer(blob, dtype=np.uint8)
        st = self._L.vl_hnsw_import_graph(self._h, _ptr(ids, C.c_uint64), _ptr(rows, C.c_float), rows.shape[0],
                                          b.ctypes.data_as(C.c_void_p), b.size)
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        for i in ids:
            self._meta.setdefault(int(i), ("", None))

    def set_beam_factor(self, factor: int) -> None:
        """Device beam width = factor x ef (ef = the reference's min(k, len), hnsw.rs:437, or the `ef` argument);
        1 = equal ef."""
        st = self._L.vl_hnsw_set_beam_factor(self._h, int(factor))
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def set_score_mode(self, mode: str) -> None:
        """"exact" (default): Flat similarity of the returned ids; "reference": the reference's quantised HNSW score."""
        st = self._L.vl_hnsw_set_score_mode(self._h, {"exact": 0, "reference": 1}[mode])
        if st != VL_OK:
            raise VectorLiteError(st, _err())

    def build_info(self) -> dict:
        b, us = C.c_uint64(0), C.c_uint64(0)
        st = self._L.vl_hnsw_build_info(self._h, C.byref(b), C.byref(us))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        return {"builder": {0: "none", 1: "host", 2: "device"}[int(b.value)], "seconds": us.value / 1e6}

    def graph_check(self) -> dict:
        out = np.zeros(6, dtype=np.uint64)
        st = self._L.vl_hnsw_graph_check(self._h, _ptr(out, C.c_uint64))
        if st != VL_OK:
            raise VectorLiteError(st, _err())
        keys = ["nodes", "edges0", "self_loops", "duplicates", "invalid", "isolated0"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def index_type(self):
        return IndexType.HNSW


class VectorIndexWrapper:
    """enum VectorIndexWrapper { Flat, HNSW } (src/lib.rs:270-346): static dispatch + metric()/index_type()."""

    def __init__(self, index: _CudaIndex):
        self.index = index

    def __getattr__(self, name):
        return getattr(self.index, name)

    def metric(self):
        return self.index.metric() if isinstance(self.index, HNSWIndex) else None

    def index_type(self):
        return self.index.index_type()
