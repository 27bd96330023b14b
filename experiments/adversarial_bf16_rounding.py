"""A constructed input on which a constant bf16 bound of 0.0040·‖x‖‖q‖ with BOTH operands rounded (the tensor-core
batched path's constant until the end of round 1) certifies a wrong top-k — and on which the worst-case constant
0.0079 now in csrc/batch.h, like the measured-norm bound of experiments/measured_norm_certificate.patch, refuses the
certificate (→ exact path → right answer).  DESIGN.md §3.

Dot product, d = 384.  Every element of the query q and of row A is u = 1 + 2^-8 − 2^-20, which bf16 rounds DOWN to
1.0 (relative error ≈ −2^-8 on both operands → −2^-7 on A·q).  Rows C (10 of them) and B (≥ 54) are bf16-exact mixes
of 1.0 and 1.0078125 whose exact scores lie just below A's, while their approximate scores lie above A's:
    exact:  A 387.005 > C 386.809 > B 385.994        approx:  C 385.305 > B 384.492 > A 384.000
so A — the true best row — is not among the K' = 64 kept rows, and kth_exact(C) > worst_approx(B) + 0.004·‖x‖max‖q‖.

  python experiments/adversarial_bf16_rounding.py          # CPU: the arithmetic of the counter-example
  python experiments/adversarial_bf16_rounding.py gpu      # GPU: batched search vs oracle (must AGREE)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
D = 384
U = np.float32(1 + 2.0 ** -8 - 2.0 ** -20)


def mix(n_hi):
    v = np.ones(D, np.float32)
    v[:n_hi] = np.float32(1.0078125)
    return v


def dataset(n_fill=4000):
    a, c, b = np.full(D, U, np.float32), mix(167), mix(63)
    fill = np.full((n_fill, D), 0.5, np.float32)
    rows = np.concatenate([fill[: n_fill // 2], b[None].repeat(60, 0), c[None].repeat(10, 0), a[None], fill[n_fill // 2:]])
    return rows, np.full(D, U, np.float32), n_fill // 2 + 70      # rows, query, position of A


def bf16(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def cpu_demo():
    rows, q, pos_a = dataset()
    a, c, b = rows[pos_a], rows[pos_a - 1], rows[pos_a - 11]
    ex = lambda x: float(np.dot(x.astype(np.float64), q.astype(np.float64)))
    ap = lambda x: float(np.dot(bf16(x).astype(np.float64), bf16(q).astype(np.float64)))
    maxn = float(np.sqrt((rows.astype(np.float64) ** 2).sum(1).max()))
    qn = float(np.linalg.norm(q.astype(np.float64)))
    worst, kth = ap(b), ex(c)
    e_x = max(float(np.linalg.norm(bf16(r) - r) / np.linalg.norm(r)) for r in (a, b, c))
    e_q = float(np.linalg.norm(bf16(q) - q) / np.linalg.norm(q))
    measured = e_x + e_q + e_x * e_q
    print(f"exact  A {ex(a):.3f}  C {ex(c):.3f}  B {ex(b):.3f}")
    print(f"approx A {ap(a):.3f}  C {ap(c):.3f}  B {ap(b):.3f}")
    print("A excluded by the approximate scan:", ap(a) < worst, "| A is the true best row:", ex(a) > kth)
    print("constant bound 0.0040 certifies:", kth > worst + 0.0040 * maxn * qn)
    print("worst-case constant 0.0079 certifies:", kth > worst + 0.0079 * maxn * qn)
    print(f"measured bound {measured:.5f} certifies:", kth > worst + measured * maxn * qn)
    assert ap(a) < worst and ex(a) > kth and kth > worst + 0.0040 * maxn * qn
    assert not kth > worst + measured * maxn * qn and not kth > worst + 0.0079 * maxn * qn


def gpu_check():
    import oracle
    import vectorlite_b200 as vl
    rows, q, pos_a = dataset()
    idx = vl.FlatIndex(D)
    idx.add_batch(np.arange(rows.shape[0], dtype=np.uint64), rows)
    queries = np.stack([q, q])                                   # nq = 2 → tensor-core batched path
    gi, gs, gc = idx.search_batch(queries, 10, vl.SimilarityMetric.DotProduct)
    st, oi, os_ = oracle.flat_search(rows, None, q, 10, int(vl.SimilarityMetric.DotProduct))
    print("oracle ids ", list(map(int, oi)))
    print("device ids ", list(map(int, gi[0])), "stats", idx.stats())
    same = list(map(int, gi[0])) == list(map(int, oi))
    print("AGREE" if same else "DIFFER (a certificate passed for a top-k without row %d)" % pos_a)
    return same


if __name__ == "__main__":
    cpu_demo()
    if len(sys.argv) > 1 and sys.argv[1] == "gpu":
        sys.exit(0 if gpu_check() else 1)
