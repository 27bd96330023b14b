"""CPU ORACLE — test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  The product (``vectorlite_b200``) never does.

``oracle.lib`` is a ctypes view of ``oracle/libvl_oracle.so`` (C++ restatement of the
reference's f64 search, see ``vl_oracle.h`` for the file:line map); ``oracle.py_oracle`` is an
independent pure-Python restatement used to cross-check the C++ one on small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvl_oracle.so")

COSINE, EUCLIDEAN, MANHATTAN, DOT = 0, 1, 2, 3
OK, ERR_DIM, ERR_DUP_ID, ERR_NOT_FOUND, ERR_METRIC_MISMATCH, ERR_INVALID, ERR_NAN = 0, 1, 2, 3, 4, 5, 8


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("vl_oracle_flat.cpp", "vl_oracle_hnsw.cpp", "vl_oracle.h")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libvl_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    dp, fp, u64p, szp, u32p = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint64),
                               C.POINTER(C.c_size_t), C.POINTER(C.c_uint32))
    L.vlo_metric.restype = C.c_double
    L.vlo_metric.argtypes = [C.c_int, dp, dp, C.c_size_t]
    L.vlo_flat_search.restype = C.c_int
    L.vlo_flat_search.argtypes = [dp, u64p, C.c_size_t, C.c_size_t, dp, C.c_size_t, C.c_size_t,
                                  C.c_int, u64p, dp, szp]
    L.vlo_flat_search_f32.restype = C.c_int
    L.vlo_flat_search_f32.argtypes = [fp, u64p, C.c_size_t, C.c_size_t, fp, C.c_size_t, C.c_size_t,
                                      C.c_int, u64p, dp, szp]
    L.vlo_flat_search_batch_f32.restype = C.c_int
    L.vlo_flat_search_batch_f32.argtypes = [fp, u64p, C.c_size_t, C.c_size_t, fp, C.c_size_t,
                                            C.c_size_t, C.c_int, C.c_int, C.c_size_t, u64p, dp]
    L.vlo_hnsw_distance.restype = C.c_uint64
    L.vlo_hnsw_distance.argtypes = [C.c_int, dp, dp, C.c_size_t]
    L.vlo_convert_distance_to_similarity.restype = C.c_double
    L.vlo_convert_distance_to_similarity.argtypes = [C.c_double, C.c_int]
    L.vlo_hnsw_create.restype = C.c_void_p
    L.vlo_hnsw_create.argtypes = [C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t]
    L.vlo_hnsw_destroy.restype = None
    L.vlo_hnsw_destroy.argtypes = [C.c_void_p]
    L.vlo_hnsw_add.restype = C.c_int
    L.vlo_hnsw_add.argtypes = [C.c_void_p, C.c_uint64, dp, C.c_size_t]
    L.vlo_hnsw_add_batch_f32.restype = C.c_int
    L.vlo_hnsw_add_batch_f32.argtypes = [C.c_void_p, u64p, fp, C.c_size_t]
    L.vlo_hnsw_delete.restype = C.c_int
    L.vlo_hnsw_delete.argtypes = [C.c_void_p, C.c_uint64]
    L.vlo_hnsw_len.restype = C.c_size_t
    L.vlo_hnsw_len.argtypes = [C.c_void_p]
    L.vlo_hnsw_search.restype = C.c_int
    L.vlo_hnsw_search.argtypes = [C.c_void_p, dp, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                  u64p, dp, szp, u64p]
    L.vlo_hnsw_search_batch_f32.restype = C.c_int
    L.vlo_hnsw_search_batch_f32.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.c_int, u64p, dp, u32p, u64p]
    L.vlo_hnsw_num_layers.restype = C.c_size_t
    L.vlo_hnsw_num_layers.argtypes = [C.c_void_p]
    L.vlo_hnsw_layer_len.restype = C.c_size_t
    L.vlo_hnsw_layer_len.argtypes = [C.c_void_p, C.c_size_t]
    L.vlo_hnsw_export_zero.restype = None
    L.vlo_hnsw_export_zero.argtypes = [C.c_void_p, u64p]
    L.vlo_hnsw_export_layer.restype = None
    L.vlo_hnsw_export_layer.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
    L.vlo_hnsw_strict_evals.restype = C.c_uint64
    L.vlo_hnsw_strict_evals.argtypes = [C.c_void_p]
    L.vlo_hnsw_levels.restype = None
    L.vlo_hnsw_levels.argtypes = [C.c_size_t, C.c_size_t, u32p]
    L.vlo_synth_rows_f32.restype = None
    L.vlo_synth_rows_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, C.c_size_t, C.c_uint32, fp]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def metric(m: int, a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, "Vectors must have the same length"  # lib.rs:381
    return _load().vlo_metric(m, _p(a, C.c_double), _p(b, C.c_double), a.size)


def flat_search(rows, ids, q, k: int, m: int):
    """FlatIndex::search (flat.rs:98-119).  rows: [n, dim] f64 or f32 (widened).  Returns
    (status, ids[<=k], scores[<=k])."""
    rows = np.ascontiguousarray(rows)
    n = rows.shape[0]
    dim = rows.shape[1] if rows.ndim == 2 else 0
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    kk = max(min(k, n), 1)
    oi = np.zeros(kk, dtype=np.uint64)
    os_ = np.zeros(kk, dtype=np.float64)
    cnt = C.c_size_t(0)
    L = _load()
    if rows.dtype == np.float32:
        q = np.ascontiguousarray(q, dtype=np.float32)
        st = L.vlo_flat_search_f32(_p(rows, C.c_float), _p(ids_a, C.c_uint64), n, dim,
                                   _p(q, C.c_float), q.size, k, m, _p(oi, C.c_uint64),
                                   _p(os_, C.c_double), C.byref(cnt))
    else:
        rows = rows.astype(np.float64, copy=False)
        q = np.ascontiguousarray(q, dtype=np.float64)
        st = L.vlo_flat_search(_p(rows, C.c_double), _p(ids_a, C.c_uint64), n, dim,
                               _p(q, C.c_double), q.size, k, m, _p(oi, C.c_uint64),
                               _p(os_, C.c_double), C.byref(cnt))
    return st, oi[:cnt.value].copy(), os_[:cnt.value].copy()


def flat_search_batch(rows, ids, queries, k: int, m: int, nthreads: int = 1, clone_bytes: int = 0):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n, dim = rows.shape
    nq = queries.shape[0]
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    oi = np.zeros((nq, k), dtype=np.uint64)
    os_ = np.zeros((nq, k), dtype=np.float64)
    st = _load().vlo_flat_search_batch_f32(_p(rows, C.c_float), _p(ids_a, C.c_uint64), n, dim,
                                           _p(queries, C.c_float), nq, k, m, nthreads, clone_bytes,
                                           _p(oi, C.c_uint64), _p(os_, C.c_double))
    return st, oi, os_


def hnsw_distance(m: int, a, b) -> int:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return _load().vlo_hnsw_distance(m, _p(a, C.c_double), _p(b, C.c_double), a.size)


def convert_distance_to_similarity(d: float, m: int) -> float:
    return _load().vlo_convert_distance_to_similarity(d, m)


def hnsw_levels(M: int, n: int):
    out = np.zeros(n, dtype=np.uint32)
    _load().vlo_hnsw_levels(M, n, _p(out, C.c_uint32))
    return out


def synth_rows(seed: int, row0: int, n: int, dim: int, clusters: int = 0):
    out = np.empty((n, dim), dtype=np.float32)
    _load().vlo_synth_rows_f32(seed, row0, n, dim, clusters, _p(out, C.c_float))
    return out


class HNSW:
    """HNSWIndex (hnsw.rs:197-518) over the restated crate graph."""

    def __init__(self, dim: int, metric: int, M: int = 16, M0: int = 32, ef_construction: int = 400):
        self._L = _load()
        self._h = self._L.vlo_hnsw_create(dim, metric, M, M0, ef_construction)
        if not self._h:
            raise ValueError("HNSW index dimension cannot be 0")  # hnsw.rs:217-219 panics
        self.dim, self.metric = dim, metric

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.vlo_hnsw_destroy(self._h)
            self._h = None

    def add(self, id_: int, v) -> int:
        v = np.ascontiguousarray(v, dtype=np.float64)
        return self._L.vlo_hnsw_add(self._h, id_, _p(v, C.c_double), v.size)

    def add_batch(self, ids, rows) -> int:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        return self._L.vlo_hnsw_add_batch_f32(self._h, _p(ids_a, C.c_uint64), _p(rows, C.c_float),
                                              rows.shape[0])

    def delete(self, id_: int) -> int:
        return self._L.vlo_hnsw_delete(self._h, id_)

    def __len__(self):
        return self._L.vlo_hnsw_len(self._h)

    def search(self, q, k: int, metric: int, ef: int = 0):
        q = np.ascontiguousarray(q, dtype=np.float64)
        oi = np.zeros(max(k, 1), dtype=np.uint64)
        os_ = np.zeros(max(k, 1), dtype=np.float64)
        cnt = C.c_size_t(0)
        vis = C.c_uint64(0)
        st = self._L.vlo_hnsw_search(self._h, _p(q, C.c_double), q.size, k, metric, ef,
                                     _p(oi, C.c_uint64), _p(os_, C.c_double), C.byref(cnt),
                                     C.byref(vis))
        return st, oi[:cnt.value].copy(), os_[:cnt.value].copy(), vis.value

    def search_batch(self, queries, k: int, ef: int = 0, nthreads: int = 1):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        oi = np.zeros((nq, k), dtype=np.uint64)
        os_ = np.zeros((nq, k), dtype=np.float64)
        cnts = np.zeros(nq, dtype=np.uint32)
        vis = C.c_uint64(0)
        st = self._L.vlo_hnsw_search_batch_f32(self._h, _p(queries, C.c_float), nq, k, ef, nthreads,
                                               _p(oi, C.c_uint64), _p(os_, C.c_double),
                                               _p(cnts, C.c_uint32), C.byref(vis))
        return st, oi, os_, cnts, vis.value

    def export_zero(self, M0: int):
        n = self.layer_len(0)
        out = np.zeros((n, M0), dtype=np.uint64)
        self._L.vlo_hnsw_export_zero(self._h, _p(out, C.c_uint64))
        return out

    def export_layer(self, l: int, M: int):
        n = self.layer_len(l)
        zn = np.zeros(n, dtype=np.uint64)
        nn = np.zeros(n, dtype=np.uint64)
        nb = np.zeros((n, M), dtype=np.uint64)
        self._L.vlo_hnsw_export_layer(self._h, l, _p(zn, C.c_uint64), _p(nn, C.c_uint64), _p(nb, C.c_uint64))
        return zn, nn, nb

    def strict_evals(self):
        return self._L.vlo_hnsw_strict_evals(self._h)

    def num_layers(self):
        return self._L.vlo_hnsw_num_layers(self._h)

    def layer_len(self, l: int):
        return self._L.vlo_hnsw_layer_len(self._h, l)
