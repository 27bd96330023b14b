#!/usr/bin/env python3
"""Writes tests/golden/reference_kats.json: every known-answer test the reference's own
unit / integration tests hold for the search hot path (SURVEY.md §8c ①-⑩ and the HNSW toys),
transcribed from the cited reference test sources.  Expected values are what the reference
test ASSERTS (id, count, score to 1e-10, inequality); "exact" fields add the IEEE-f64 value of
the reference formula evaluated in pure Python (oracle/py_oracle.py), which obeys the same
arithmetic rules as rustc output.

Run from the repo root:  python tests/golden/make_reference_kats.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import py_oracle as po  # noqa: E402

COS, EUC, MAN, DOT = 0, 1, 2, 3


def flat_case(name, cite, rows, query, k, metric, expect_ids_prefix, expect_len, asserts):
    idx = po.FlatIndex(len(rows[0][1]), rows)
    res = idx.search(query, k, metric)
    return {
        "name": name, "cite": cite, "rows": [{"id": i, "values": v} for i, v in rows],
        "query": query, "k": k, "metric": metric,
        "expect_len": expect_len, "expect_ids_prefix": expect_ids_prefix, "asserts": asserts,
        "exact_ids": [r[0] for r in res], "exact_scores": [r[1] for r in res],
        "exact_scores_hex": [float(r[1]).hex() for r in res],
    }


unit3 = [(1, [1.0, 0.0, 0.0]), (2, [0.0, 1.0, 0.0]), (3, [0.0, 0.0, 1.0])]
pts2 = [(1, [0.0, 0.0]), (2, [3.0, 4.0]), (3, [6.0, 8.0])]

flat = [
    flat_case("flat_cosine_identical", "src/index/flat.rs:187-201", unit3, [1.0, 0.0, 0.0], 2, COS,
              [1], 2, [{"index": 0, "score": 1.0, "tol": 1e-10}]),
    flat_case("flat_serde_roundtrip_search", "src/index/flat.rs:145-184", unit3, [1.1, 0.1, 0.1], 2,
              COS, [1], 2, [{"index": 0, "score_gt": 0.99}, {"non_increasing": True}]),
    flat_case("wrapper_roundtrip_search", "src/lib.rs:697-724",
              [(1, [1.0, 0.0, 0.0]), (2, [0.0, 1.0, 0.0])], [1.1, 0.1, 0.1], 1, COS, [1], 1, []),
    flat_case("flat_euclidean", "src/index/flat.rs:204-218", pts2, [0.0, 0.0], 2, EUC, [1], 2,
              [{"index": 0, "score": 1.0, "tol": 1e-10}]),
    flat_case("flat_manhattan", "src/index/flat.rs:221-235", pts2, [0.0, 0.0], 2, MAN, [1], 2,
              [{"index": 0, "score": 1.0, "tol": 1e-10}]),
    flat_case("flat_dot", "src/index/flat.rs:238-252",
              [(1, [1.0, 2.0]), (2, [2.0, 1.0]), (3, [0.0, 0.0])], [1.0, 2.0], 2, DOT, [1], 2,
              [{"index": 0, "score": 5.0, "tol": 1e-10}]),
    flat_case("flat_cosine_vs_dot_cos", "src/index/flat.rs:255-274",
              [(1, [1.0, 2.0]), (2, [2.0, 1.0])], [1.0, 2.0], 1, COS, [1], 1, []),
    flat_case("flat_cosine_vs_dot_dot", "src/index/flat.rs:255-274",
              [(1, [1.0, 2.0]), (2, [2.0, 1.0])], [1.0, 2.0], 1, DOT, [1], 1, []),
    flat_case("persistence_search", "src/persistence.rs:247-249,279-281",
              [(0, [1.0, 2.0, 3.0]), (1, [4.0, 5.0, 6.0])], [1.1, 2.1, 3.1], 1, COS, [0], 1, []),
    flat_case("vector_store_search", "src/lib.rs:681-694",
              [(0, [1.0, 0.0, 0.0]), (1, [0.0, 1.0, 0.0]), (2, [0.0, 0.0, 1.0])], [1.0, 0.0, 0.0], 2,
              COS, [0], 2, [{"index": 0, "score": 1.0, "tol": 1e-10}]),
    # mock embedder returns the same vector for every text → all scores tie → first inserted wins
    flat_case("all_equal_embeddings_tie", "src/client.rs:665-667; tests/http_integration_test.rs:181-209",
              [(0, [1.0, 2.0, 3.0]), (1, [1.0, 2.0, 3.0]), (2, [1.0, 2.0, 3.0])], [1.0, 2.0, 3.0], 1,
              COS, [0], 1, []),
]

metric_kats = [  # src/lib.rs:579-662, all to 1e-10
    {"cite": "src/lib.rs:579-583", "metric": COS, "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expect": 1.0},
    {"cite": "src/lib.rs:586-590", "metric": COS, "a": [1.0, 0.0], "b": [0.0, 1.0], "expect": 0.0},
    {"cite": "src/lib.rs:593-597", "metric": COS, "a": [1.0, 2.0, 3.0], "b": [-1.0, -2.0, -3.0], "expect": -1.0},
    {"cite": "src/lib.rs:600-604", "metric": EUC, "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expect": 1.0},
    {"cite": "src/lib.rs:607-612", "metric": EUC, "a": [0.0, 0.0], "b": [3.0, 4.0], "expect": 1.0 / 6.0},
    {"cite": "src/lib.rs:615-619", "metric": MAN, "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expect": 1.0},
    {"cite": "src/lib.rs:622-627", "metric": MAN, "a": [0.0, 0.0], "b": [3.0, 4.0], "expect": 1.0 / 8.0},
    {"cite": "src/lib.rs:630-635", "metric": DOT, "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expect": 14.0},
    {"cite": "src/lib.rs:638-642", "metric": DOT, "a": [1.0, 0.0], "b": [0.0, 1.0], "expect": 0.0},
    {"cite": "src/lib.rs:645-650", "metric": DOT, "a": [1.0, 2.0, 3.0], "b": [-1.0, -2.0, -3.0], "expect": -14.0},
]
for m in metric_kats:
    m["tol"] = 1e-10
    m["exact_hex"] = float(po.calculate(m["metric"], m["a"], m["b"])).hex()

convert_kats = [  # src/index/hnsw.rs:807-1032
    {"metric": EUC, "d": 0.0, "expect": 1.0, "tol": 0.0}, {"metric": EUC, "d": 0.5, "expect": 1 / 1.5, "tol": 1e-10},
    {"metric": EUC, "d": 1.0, "expect": 0.5, "tol": 1e-10}, {"metric": EUC, "d": 10.0, "expect": 1 / 11.0, "tol": 1e-10},
    {"metric": COS, "d": 0.0, "expect": 1.0, "tol": 0.0}, {"metric": COS, "d": 100.0, "expect": 0.9, "tol": 1e-10},
    {"metric": COS, "d": 500.0, "expect": 0.5, "tol": 1e-10}, {"metric": COS, "d": 2000.0, "expect": -1.0, "tol": 0.0},
    {"metric": MAN, "d": 0.0, "expect": 1.0, "tol": 0.0}, {"metric": MAN, "d": 1.0, "expect": 0.5, "tol": 1e-10},
    {"metric": DOT, "d": 0.0, "expect": 1.0, "tol": 0.0}, {"metric": DOT, "d": 2000.0, "expect": 0.0, "tol": 0.0},
    {"metric": DOT, "d": 500.0, "expect": 0.5, "tol": 1e-10},
]

hnsw_rows = [(100, [1.0, 0.0, 0.0]), (200, [0.0, 1.0, 0.0]), (300, [0.0, 0.0, 1.0]), (400, [1.0, 1.0, 0.0])]
hq = [1.1, 0.1, 0.1]
hnsw = {
    "id_mapping": {  # src/index/hnsw.rs:605-634
        "cite": "src/index/hnsw.rs:605-634", "metric": EUC, "rows": [{"id": i, "values": v} for i, v in hnsw_rows],
        "query": hq, "k": 2, "expect_first_id": 100,
        "quantised_distances": [po.hnsw_distance(EUC, v, hq) for _, v in hnsw_rows],
        "scores_by_row": [po.convert_distance_to_similarity(po.hnsw_distance(EUC, v, hq) / 1000.0, EUC)
                          for _, v in hnsw_rows],
    },
    "search_basic": {  # src/index/hnsw.rs:566-593
        "cite": "src/index/hnsw.rs:566-593", "metric": EUC,
        "rows": [{"id": i + 1, "values": v} for i, (_, v) in enumerate(hnsw_rows)], "query": hq, "k": 2,
    },
}

out = {"flat": flat, "metrics": metric_kats, "convert": convert_kats, "hnsw": hnsw}
path = os.path.join(os.path.dirname(__file__), "reference_kats.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
