"""Pure-Python second oracle (test infrastructure only — see oracle/__init__.py).

Python floats are IEEE-754 binary64 and CPython never fuses a*b+c, so plain loops obey the
same arithmetic rules as the Rust reference.  Independent of the C++ restatement; used to
cross-check it on small cases and to hold the reference's known-answer tests.

Follows: src/lib.rs:425-444 (cosine), 476-489 (euclidean), 521-532 (manhattan), 565-572 (dot),
380-391 (dispatch); src/index/flat.rs:82-119 (add / delete / search);
src/index/hnsw.rs:113-174 (u64 functors), 51-75 (score conversion).
"""
from __future__ import annotations

import math

COSINE, EUCLIDEAN, MANHATTAN, DOT = 0, 1, 2, 3


def cosine_similarity(a, b) -> float:
    assert len(a) == len(b), "Vectors must have the same length"
    dot = na = nb = 0.0
    for x, y in zip(a, b):
        dot += x * y
        na += x * x
        nb += y * y
    norm_a, norm_b = math.sqrt(na), math.sqrt(nb)
    if norm_a == 0.0 or norm_b == 0.0:
        return 0.0
    return dot / (norm_a * norm_b)


def euclidean_similarity(a, b) -> float:
    assert len(a) == len(b), "Vectors must have the same length"
    s = 0.0
    for x, y in zip(a, b):
        d = x - y
        s += d * d
    return 1.0 / (1.0 + math.sqrt(s))


def manhattan_similarity(a, b) -> float:
    assert len(a) == len(b), "Vectors must have the same length"
    s = 0.0
    for x, y in zip(a, b):
        s += abs(x - y)
    return 1.0 / (1.0 + s)


def dot_product(a, b) -> float:
    assert len(a) == len(b), "Vectors must have the same length"
    s = 0.0
    for x, y in zip(a, b):
        s += x * y
    return s


_FN = {COSINE: cosine_similarity, EUCLIDEAN: euclidean_similarity,
       MANHATTAN: manhattan_similarity, DOT: dot_product}


def calculate(metric: int, a, b) -> float:
    return _FN[metric](a, b)


class DimensionMismatch(Exception):
    def __init__(self, expected, actual):
        super().__init__(f"Dimension mismatch: expected {expected}, got {actual}")
        self.expected, self.actual = expected, actual


class FlatIndex:
    """src/index/flat.rs:59-136."""

    def __init__(self, dim: int, data=None):
        self.dim = dim
        self.data = list(data or [])  # [(id, values)]

    def add(self, id_: int, values):
        if len(values) != self.dim:
            raise ValueError("Vector dimension mismatch")          # flat.rs:83-85
        if any(e[0] == id_ for e in self.data):
            raise ValueError(f"Vector ID {id_} already exists")     # flat.rs:86-88
        self.data.append((id_, [float(v) for v in values]))

    def delete(self, id_: int):
        self.data = [e for e in self.data if e[0] != id_]           # flat.rs:94 (missing id: Ok)

    def search(self, query, k: int, metric: int):
        if self.data and len(query) != self.dim:                    # flat.rs:99-104
            raise DimensionMismatch(self.dim, len(query))
        sims = [(e[0], calculate(metric, e[1], query)) for e in self.data]  # flat.rs:106-114
        if len(sims) >= 2 and any(math.isnan(s) for _, s in sims):
            raise FloatingPointError("NaN score: the reference panics (flat.rs:116)")
        # flat.rs:116 — stable, descending; Python's sort is stable and -0.0 == 0.0.
        sims.sort(key=lambda t: -t[1])
        return sims[:k]                                             # flat.rs:117


def _as_u64(d: float) -> int:
    if math.isnan(d) or d <= 0.0:
        return 0
    if d >= 18446744073709551616.0:
        return (1 << 64) - 1
    return int(d)


def hnsw_distance(metric: int, a, b) -> int:
    """src/index/hnsw.rs:113-174."""
    if metric == EUCLIDEAN:
        s = 0.0
        for x, y in zip(a, b):
            d = x - y
            s += d * d
        return _as_u64(math.sqrt(s) * 1000.0)
    if metric == COSINE:
        dot = na = nb = 0.0
        for x, y in zip(a, b):
            dot = dot + x * y
            na = na + x * x
            nb = nb + y * y
        norm_a, norm_b = math.sqrt(na), math.sqrt(nb)
        if norm_a == 0.0 or norm_b == 0.0:
            return 1000
        return _as_u64((1.0 - dot / (norm_a * norm_b)) * 1000.0)
    if metric == MANHATTAN:
        s = 0.0
        for x, y in zip(a, b):
            s += abs(x - y)
        return _as_u64(s * 1000.0)
    s = 0.0
    for x, y in zip(a, b):
        s += x * y
    return _as_u64(1000.0 - min(max(s, -1000.0), 1000.0))


def convert_distance_to_similarity(distance: float, metric: int) -> float:
    """src/index/hnsw.rs:51-75."""
    if metric in (EUCLIDEAN, MANHATTAN):
        return 1.0 / (1.0 + distance)
    if metric == COSINE:
        return 1.0 - distance / 1000.0
    return min(max((1000.0 - distance) / 1000.0, 0.0), 1.0)
