#!/usr/bin/env python3
"""bench.py — headline benchmark of the VectorLite B200 search hot path.

Workload (BASELINE.json configs[1]): Flat index, 1M × 384-d f32 synthetic unit vectors, cosine,
k = 10, single-query scans (B = 1).  One "step" = QUERIES_PER_STEP independent single-query
searches back to back (each one a full HBM-bound scan of the 1.536 GB store + fused top-k +
fp64 rescore + certificate).  With N GPUs the store is row-sharded: every rank holds a 1M-row
shard (weak scaling: N·1M rows in total), every query is searched on all shards and the per-shard
top-k are exchanged with one NCCL all-gather and merged on the device.

  value     queries/s × shards (1M-row shard-scans per second, whole job), inputs resident in HBM
  e2e       same metric through the host C-ABI call (vl_index_search / ShardedFlatIndex.search):
            queries start in host memory, results end in host memory, copies inside the timed region
  roofline  flat_scan_kernel: algorithmic bytes per launch ÷ CUDA-event duration vs measured HBM peak
  cpu_baseline  the C++ oracle restatement of the reference's FlatIndex::search on the host cores

`--impl reference` times the reference arm: the reference is Rust and cannot be built in this
image (no cargo/rustc), so the arm is the oracle port (C++ restatement of flat.rs:98-119 +
lib.rs:425-444, f64, sequential) on all host threads, one query per thread.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DIM = 384
METRIC_NAMES = {"cosine": 0, "euclidean": 1, "manhattan": 2, "dot": 3}
QUERIES_PER_STEP = 1024      # a step list of 20 lasts > 2 s on the device (0.113 ms per query)
ORACLE_CHECK_QUERIES = 64     # N = 1: every one of these bench queries is compared with the CPU oracle (ids + f64 bits)
E2E_CALLERS = min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4))
# ^ concurrent host threads calling vl_index_search in the e2e leg (N = 1): as many as the reference arm uses
#   (one query per host thread)


def native_callers_group(L, group, queries, k, metric, n_threads, total):
    """native_callers for a shard group (vl_group_search)."""
    import ctypes as C
    so = os.path.join(ROOT, "scripts", "libvl_native_callers.so")
    if not os.path.exists(so):
        return None
    N = C.CDLL(so)
    if not hasattr(N, "vl_native_callers_group"):
        return None
    N.vl_native_callers_group.restype = C.c_double
    N.vl_native_callers_group.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                          C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq, dim = q.shape
    ids = np.zeros((nq, k), dtype=np.uint64)
    sc = np.zeros((nq, k), dtype=np.float64)
    cnt = np.zeros(nq, dtype=np.uint32)
    fn = C.cast(L.vl_group_search, C.c_void_p)
    dt = N.vl_native_callers_group(fn, group, q.ctypes.data, nq, dim, k, int(metric), n_threads, total,
                                   ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data)
    if dt <= 0:
        raise RuntimeError(f"native callers (group): vl_group_search failed with status {int(-dt)}")
    return total / dt, ids, sc


def native_callers(index, queries, k, metric, ef, n_threads, total):
    """`total` single-query vl_index_search calls (host buffers in, host results out) issued by `n_threads` plain
    host threads sharing one cursor (scripts/native_callers.cpp) — the reference's serving pattern without the
    Python interpreter lock between calls.  Returns (queries/s, ids[nq,k], scores[nq,k]) or None when the helper
    is not built."""
    import ctypes as C
    so = os.path.join(ROOT, "scripts", "libvl_native_callers.so")
    if not os.path.exists(so):
        return None
    N = C.CDLL(so)
    N.vl_native_callers.restype = C.c_double
    N.vl_native_callers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                    C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq, dim = q.shape
    ids = np.zeros((nq, k), dtype=np.uint64)
    sc = np.zeros((nq, k), dtype=np.float64)
    cnt = np.zeros(nq, dtype=np.uint32)
    fn = C.cast(index._L.vl_index_search, C.c_void_p)
    dt = N.vl_native_callers(fn, index._h, q.ctypes.data, nq, dim, k, int(metric), ef, n_threads, total,
                             ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data)
    if dt <= 0:
        raise RuntimeError(f"native callers: vl_index_search failed with status {int(-dt)}")
    return total / dt, ids, sc


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bench_config(args):
    """`config` of BOTH arms (identical dicts: the driver compares them); arm-specific notes live elsewhere."""
    return {"workload": f"flat {args.rows}x{DIM} f32 per GPU shard, {args.metric}, k={args.k}, B=1 "
                        f"({QUERIES_PER_STEP} single-query searches per step)",
            "rows_per_shard": args.rows, "dim": DIM, "k": args.k, "metric": args.metric, "batch": 1,
            "queries_per_step": QUERIES_PER_STEP,
            "data_recipe": "counter-based synthetic unit vectors (rows: stream 42, queries: stream 43)",
            "l2": "inputs larger than L2 (1.536 GB store vs 126 MB)"}


def device_synth_rows(vl, seed, first_row, n, device, clusters=0):
    """[n, DIM] f32 synthetic rows `first_row …` of stream `seed` from the PRODUCT's own counter-based generator
    (vl_index_fill_synthetic; bit-identical to the oracle's synth_rows, tests/test_flat_gpu.py
    test_device_generator_matches_oracle), so the measured arm does not need oracle/ for its inputs."""
    src = vl.FlatIndex(DIM, device=device)
    try:
        src.fill_synthetic(seed, n, first_row=first_row, clusters=clusters)
        rows = np.ascontiguousarray(src.export()[1], dtype=np.float32)
    finally:
        src.close()
    assert rows.shape == (n, DIM)
    return rows


def cpu_flat_qps(oracle, rows, queries, k, metric, threads, clone_bytes=0):
    t0 = time.perf_counter()
    st, ids, _ = oracle.flat_search_batch(rows, None, queries, k, metric, nthreads=threads, clone_bytes=clone_bytes)
    dt = time.perf_counter() - t0
    assert st == 0
    return queries.shape[0] / dt, dt, ids


def synth_host_rows(oracle, seed, n, dim, threads, first_row=0, clusters=0):
    """Host copy of rows [first_row, first_row + n) of the synthetic store (counter-based → chunks in parallel)."""
    out = np.empty((n, dim), dtype=np.float32)
    chunk = (n + threads - 1) // threads

    def work(t):
        lo, hi = t * chunk, min(n, (t + 1) * chunk)
        if lo < hi:
            out[lo:hi] = oracle.synth_rows(seed, first_row + lo, hi - lo, dim, clusters)
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    return out


def hnsw_section(vl, n, efc, device, nq=4096, k=10, clusters=1024):
    """HNSW default profile (M/M0 = 16/32) on the 1024-centre mixture: bulk build (device builder), then
    QPS (host API, 4096-query batches, copies included), recall@10 vs the exact flat result and
    evaluated nodes per query over the ef sweep.  ef = 0 is the reference's own setting (ef = k)."""
    metric = vl.SimilarityMetric.Cosine
    flat = vl.FlatIndex(DIM, device=device)
    flat.fill_synthetic(42, n, clusters=clusters)
    qsrc = vl.FlatIndex(DIM, device=device)
    qsrc.fill_synthetic(43, nq, clusters=clusters)
    import torch
    queries = torch.from_numpy(qsrc.export()[1]).pin_memory().numpy()   # e2e leg: inputs start in PINNED host memory
    truth, _, _ = flat.search_batch(queries, k, metric)          # exact, certified (tensor-core batched path)
    ids, rows = flat.export()
    flat.close()
    h = vl.HNSWIndex(DIM, metric, M=16, M0=32, ef_construction=efc, device=device)
    t0 = time.perf_counter()
    h.add_batch(ids, rows)       # bulk add into an empty index: built on the device (csrc/hnsw_build.cu)
    h.build()
    build_s = time.perf_counter() - t0
    build_info = h.build_info()
    del rows
    # `ef` is the reference's ef (hnsw.rs:437; 0 = min(k, len)); the device beam is beam_factor x ef.  Both the
    # default factor and beam = ef exactly are reported, each with the nodes it evaluates per query.
    def run_sweep(factor):
        h.set_beam_factor(factor)
        out = {}
        for ef in (0, 16, 32, 64, 128, 256):
            if (ef or k) * factor > 2048:
                continue
            h.search_batch(queries, k, metric, ef)      # warm-up at the timed batch size (same kernel variant)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                gi, gs, gc = h.search_batch(queries, k, metric, ef)
            dt = (time.perf_counter() - t0) / reps
            hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
            out[str(ef)] = {"recall_at_10": hit / (nq * k), "qps_e2e": nq / dt,
                            "visited_per_query": h.stats()["hnsw_visited"] / nq, "beam": factor * (ef or k)}
        return out
    sweep_equal_ef = run_sweep(1)
    sweep = run_sweep(8)
    # one query per call at the reference's own setting (ef = k): the lone-client latency of the host API
    for i in range(20):
        h.search_batch(queries[i:i + 1], k, metric, 0)
    t0 = time.perf_counter()
    for i in range(200):
        h.search_batch(queries[i:i + 1], k, metric, 0)
    single_us = (time.perf_counter() - t0) / 200 * 1e6
    # the same one-query calls from E2E_CALLERS concurrent threads (combined into shared launches by the handle)
    import itertools
    from concurrent.futures import ThreadPoolExecutor
    total_conc = 4096

    def conc_run(total):
        cursor = itertools.count()

        def work(_):
            while True:
                i = next(cursor)
                if i >= total:
                    return
                h.search_batch(queries[i % nq:i % nq + 1], k, metric, 0)
        with ThreadPoolExecutor(max_workers=E2E_CALLERS) as pool:
            list(pool.map(work, range(E2E_CALLERS)))
    conc_run(256)
    t0 = time.perf_counter()
    conc_run(total_conc)
    conc_qps = total_conc / (time.perf_counter() - t0)
    conc_impl = "python threads (ctypes)"
    nat = native_callers(h, queries[:1024], k, metric, 0, E2E_CALLERS, 2 * total_conc)
    if nat is not None:
        conc_py, conc_qps, conc_impl = conc_qps, nat[0], "native host threads (scripts/native_callers.cpp)"
    else:
        conc_py = None
    ref = {}
    for name in ("hnsw_reference_recall_n1000000_c1024_M16.json", "hnsw_reference_recall_n200000_c1024_M16.json",
                 "hnsw_reference_recall_n200000_c0_M16.json", "hnsw_reference_recall_n20000_c1024_M16.json",
                 "hnsw_reference_recall_n20000_c0_M16.json"):
        pth = os.path.join(ROOT, "tests", "golden", name)
        if os.path.exists(pth):
            d = json.load(open(pth))
            ref[name] = {e: round(v["recall_at_10"], 4) for e, v in d["sweep"].items()}
            if "qps_4threads" in next(iter(d["sweep"].values())):
                # timed when the golden file was generated (build container's host cores, 4 threads) — not this box
                ref[name]["qps_4threads_when_generated"] = {e: round(v["qps_4threads"]) for e, v in d["sweep"].items()}
                ref[name]["visited_per_query"] = {e: round(v["visited_per_query"]) for e, v in d["sweep"].items()}
            if "layer0" in d:
                ref[name]["layer0_graph"] = d["layer0"]
    h.close()
    return {"rows": n, "dim": DIM, "data": f"synthetic {clusters}-centre mixture, unit norm", "M": 16, "M0": 32,
            "ef_construction": efc, "k": k, "batch": nq, "build_seconds": build_s, "builder": build_info,
            "host_threads": cpu_threads(),
            "beam_factor_default": 8, "sweep": sweep, "sweep_beam_equals_ef": sweep_equal_ef,
            "single_query_latency_us_ef_k": single_us,
            "concurrent_single_query_callers": {"callers": E2E_CALLERS, "qps_e2e_ef_k": conc_qps, "callers_impl": conc_impl,
                                                "python_callers_qps": conc_py},
            "reference_restatement_recall": ref,
            "note": "reference recall = oracle restatement of crate hnsw 0.11 + u64-quantised functors at efC=400 "
                    "(tests/golden/, one CPU thread like the reference's insert: the 1M-row clustered set is this "
                    "section's data; i.i.d. rows stop at 200K because the restated insert visits most of the graph, "
                    "~n^1.8).  Device beam = beam_factor x ef: `sweep` uses the default factor 8, "
                    "`sweep_beam_equals_ef` factor 1; parity (ours >= reference, no slack) is asserted in "
                    "tests/test_hnsw_gpu.py at both mappings on clustered rows and at the default one on i.i.d. rows"}


def hnsw_reference_cpu(oracle, threads, n=10_000, efc=400, clusters=1024, nq=512, k=10):
    """The reference's HNSW search on the host cores: the oracle restatement of crate hnsw 0.11 + the u64
    milli-unit functors (hnsw.rs:113-174,415-496) on the bench's mixture data.  Bounded: the restated insert is
    single-threaded like the reference's (≈ 40 s for 10K rows at ef_construction = 400), so the graph is small;
    the 1M-row figures of the CUDA index are in the other arm's `hnsw` block."""
    rows = oracle.synth_rows(42, 0, n, DIM, clusters)
    q = oracle.synth_rows(43, 0, nq, DIM, clusters)
    st, truth, _ = oracle.flat_search_batch(rows, None, q, k, 0, nthreads=threads)
    assert st == 0
    h = oracle.HNSW(DIM, 0, 16, 32, efc)
    t0 = time.perf_counter()
    h.add_batch(None, rows)
    build_s = time.perf_counter() - t0
    sweep = {}
    for ef in (0, 64):
        rec = {}
        for th in (1, threads):
            t0 = time.perf_counter()
            st, ri, _, rc, vis = h.search_batch(q, k, ef, nthreads=th)
            dt = time.perf_counter() - t0
            rec[f"qps_{th}_threads" if th > 1 else "qps_1_thread"] = nq / dt
        hit = sum(len(set(map(int, ri[i, :rc[i]])) & set(map(int, truth[i]))) for i in range(nq))
        rec.update({"recall_at_10": hit / (nq * k), "visited_per_query": vis / nq})
        sweep[str(ef)] = rec
    return {"rows": n, "dim": DIM, "data": f"synthetic {clusters}-centre mixture, unit norm", "M": 16, "M0": 32,
            "ef_construction": efc, "k": k, "queries": nq, "build_seconds_1_thread": build_s, "host_threads": threads,
            "sweep": sweep, "kind": "port (restatement of the crate's published algorithm; graph parity unpinned)"}


def cpu_config1(oracle, threads, k, metric, nq=64):
    """BASELINE config 1 on the host cores: flat 10K x 384, one query at a time (the reference's own intended size,
    README.md:92-95 "<10K vectors"): microseconds per query on ONE thread (what one reference query uses) and the
    queries/s of `threads` threads with one query each."""
    rows = oracle.synth_rows(42, 0, 10_000, DIM)
    q = oracle.synth_rows(43, 0, nq, DIM)
    cpu_flat_qps(oracle, rows, q[:4], k, metric, 1)
    qps1, _, ids = cpu_flat_qps(oracle, rows, q, k, metric, 1)
    qpsT, _, _ = cpu_flat_qps(oracle, rows, np.tile(q, (4, 1)), k, metric, threads)
    return {"rows": 10_000, "dim": DIM, "k": k, "latency_us_1_thread": 1e6 / qps1, "qps_1_thread": qps1,
            f"qps_{threads}_threads": qpsT, "queries": nq}, ids


def run_reference(args):
    """The reference arm: CPU only, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = cpu_threads()
    n = args.rows
    metric = METRIC_NAMES[args.metric]
    rows = synth_host_rows(oracle, 42, n, DIM, threads)
    # bounded sample: one query per host thread per step (≈0.7 s of CPU per query at 1M rows)
    nq = threads
    queries = oracle.synth_rows(43, 0, nq * (args.steps + args.warmup), DIM)
    for w in range(args.warmup):
        cpu_flat_qps(oracle, rows, queries[w * nq:(w + 1) * nq], args.k, metric, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        off = (args.warmup + s) * nq
        cpu_flat_qps(oracle, rows, queries[off:off + nq], args.k, metric, threads)
    dt = time.perf_counter() - t0
    qps = nq * args.steps / dt
    # outside the timed steps: the same search with the reference's per-row String clone (flat.rs:111-112, a 16-byte
    # text per row) and on ONE thread (the reference's own per-query behaviour) — SURVEY §8d brackets
    qps_clone, _, _ = cpu_flat_qps(oracle, rows, queries[:nq], args.k, metric, threads, clone_bytes=16)
    qps_1t, _, _ = cpu_flat_qps(oracle, rows, queries[:1], args.k, metric, 1)
    del rows
    c1, _ = cpu_config1(oracle, threads, args.k, metric)
    hnsw_ref = None
    if args.hnsw_rows > 0:
        hnsw_ref = hnsw_reference_cpu(oracle, threads)
    line = {
        "impl": "reference", "metric": "flat_1m_384d_k10_qps", "value": qps,
        "unit": "queries/s x 1M-row shards", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{nq} queries per step (one per host thread; bounded sample of the {QUERIES_PER_STEP}-query "
                                   f"step), {nq * args.steps} queries over the full {n}-row store in the timed region",
                         "with_16B_text_clone_per_row_qps": qps_clone, "single_thread_qps": qps_1t,
                         "config1_flat_10k": c1,
                         "note": "reference is Rust (no toolchain here): C++ oracle restatement of flat.rs:98-119 + "
                                 "lib.rs:425-572, f32 rows widened to f64, -O2 -ffp-contract=off"},
        "e2e": {"value": qps, "unit": "queries/s x 1M-row shards", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0, "config1_flat_10k": c1},
        "gpu_launches": 0,
        "hnsw": hnsw_ref,
    }
    emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything that writes to fd 1 (NCCL's version / INFO banner, library chatter) to stderr; the
    one JSON line goes to the real stdout through emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="rows per GPU shard")
    ap.add_argument("--metric", default="cosine", choices=list(METRIC_NAMES))
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline and the oracle comparison")
    ap.add_argument("--extras", action="store_true", help="also time the other metrics / batch modes at N > 1")
    ap.add_argument("--hnsw-rows", type=int, default=1_000_000, help="HNSW section size (0 = skip; rank 0, N=1 only)")
    ap.add_argument("--hnsw-efc", type=int, default=400, help="ef_construction (reference default: 400)")
    ap.add_argument("--wide-rows", type=int, default=500_000, help="rows of the 768-d leg (0 = skip; rank 0, N=1 only)")
    ap.add_argument("--config5-rows", type=int, default=100_000_000,
                    help="BASELINE config 5 leg: total rows of the sharded B=1024, k=100 batch search (0 = skip); at "
                         "N = 1 one shard of the 8-GPU configuration (rows/8) is measured")
    ap.add_argument("--config4-rows", type=int, default=1_000_000,
                    help="BASELINE config 4 leg: HNSW memory-optimized / high-accuracy profiles, one replica per GPU, "
                         "4096 queries per replica (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import vectorlite_b200 as vl
    from vectorlite_b200.sharded import ShardedFlatIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: vectorlite_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side barriers that keep the GPUs idle while rank 0 measures
    vl.lib()

    metric = vl.SimilarityMetric(METRIC_NAMES[args.metric])
    n_shard, k = args.rows, args.k
    n_total = n_shard * world
    idx = ShardedFlatIndex(DIM, rank=rank, world=world, device=local_rank)
    idx.fill_synthetic(42, n_total)
    assert idx.local.len() == n_shard
    idx.local.set_pipelined(True)   # PDL: scan of query i+1 overlaps the rescore/certify kernel of query i

    nq_pool = QUERIES_PER_STEP
    queries = device_synth_rows(vl, 43, 0, nq_pool, local_rank)   # the measured arm never touches oracle/
    d_queries = torch.from_numpy(queries).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def step_device():
        # a stream of independent single-query searches; with N > 1 the per-query exchange (peer-memory pushes by
        # the finalize kernel + stamp-waiting merge kernel) rides the same PDL chain and overlaps the next scan
        for qi in range(QUERIES_PER_STEP):
            if world > 1:
                idx.search_device_pipelined(d_queries[qi:qi + 1], k, metric)
            else:
                idx.search_device(d_queries[qi:qi + 1], k, metric)
        if world > 1:
            idx.drain()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- correctness gate before any number counts: the TIMED path (one search per query) on the first
    # ORACLE_CHECK_QUERIES queries of the pool: certificates must hold; ids + scores are compared with the CPU
    # oracle further down (all of them at N = 1, 8 sampled ones through a distributed oracle merge at N > 1)
    n_check = ORACLE_CHECK_QUERIES if world == 1 else 8
    got_ids = np.zeros((n_check, k), dtype=np.uint64)
    got_sc = np.zeros((n_check, k), dtype=np.float64)
    for qi in range(n_check):
        o_ids, o_sc, o_cnt, flg = idx.search_device(d_queries[qi:qi + 1], k, metric)
        torch.cuda.synchronize()
        assert int((flg & 17).max()) == 0, "optimality certificate / exchange failed on the bench workload"
        got_ids[qi] = o_ids.cpu().numpy().view(np.uint64)[0]
        got_sc[qi] = o_sc.cpu().numpy()[0]

    # ---- device-resident throughput (value) -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    st0 = idx.local.stats()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches = idx.local.stats()["launches"] - st0
    merges = args.steps * QUERIES_PER_STEP
    n_queries = args.steps * QUERIES_PER_STEP
    qps_global = n_queries / (ms * 1e-3)
    value = qps_global * world

    # ---- dominant-kernel duration inside the same timed region (events around each scan launch) ---
    def scan_roofline(metric_):
        """(kernel name, bytes/element, algorithmic bytes per launch, avg kernel ms, launches, traffic) of the
        single-query scan as configured: events bracket every scan launch of a timed pass."""
        b0 = idx.local.stats()["bf16_scans"]
        idx.local.set_profiling(True)
        timed(step_device if metric_ == metric else (lambda: [idx.search_device(d_queries[qi:qi + 1], k, metric_)
                                                              for qi in range(QUERIES_PER_STEP)]),
              max(1, min(args.steps, 1024 // QUERIES_PER_STEP)))
        scan_ms, scan_n = idx.local.profile_read()
        idx.local.set_profiling(False)
        bf16 = idx.local.stats()["bf16_scans"] > b0
        cos, l2 = metric_ == vl.SimilarityMetric.Cosine, metric_ == vl.SimilarityMetric.Euclidean
        if bf16:   # bf16 mirror (cosine: pre-normalised rows, no norm array; L2: + fp32 ‖row‖²)
            ab = n_shard * DIM * 2 + (n_shard * 4 if l2 else 0) + DIM * 4
            name, tfile = "flat_scan_kernel<METRIC,3,BF16=true> (bf16 mirror)", "r01_flat_scan_bf16_traffic.json"
        else:
            ab = n_shard * DIM * 4 + (n_shard * 4 if cos else 0) + DIM * 4
            name, tfile = "flat_scan_kernel<METRIC,NCH,BF16=false> (fp32 arena)", "r01_flat_scan_traffic.json"
        tr = None
        tp = os.path.join(ROOT, "profiles", tfile)
        if os.path.exists(tp):
            try:
                tr = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                tr = None
        return name, (2 if bf16 else 4), ab, scan_ms / max(scan_n, 1), scan_n, tr

    scan_kernel, scan_elem_bytes, algo_bytes, scan_ms_avg, scan_n, traffic = scan_roofline(metric)
    peak, peak_src = load_peaks()
    achieved = algo_bytes / (scan_ms_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": scan_kernel, "kernel_ms": scan_ms_avg, "launches_timed": scan_n,
                "algorithmic_bytes_per_launch": algo_bytes, "scanned_copy_bytes_per_element": scan_elem_bytes,
                "frac_of_nominal_8000": achieved / 8000.0,
                "scanned_copy": ("bf16 mirror of the rows, 2 B/element (SURVEY §8d); the fp32-arena scan (4 B/element) is "
                                 "roofline.fp32_arena") if scan_elem_bytes == 2 else "fp32 arena"}

    # SURVEY §8d's 4 B/element headline: the same single-query search with scans pinned to the fp32 arena
    idx.local.set_mode(vl.Mode.Fp32)
    step_device()
    f32_name, _, f32_bytes, f32_ms, f32_n, f32_tr = scan_roofline(metric)
    t_ms = timed(step_device, 3)
    roofline["fp32_arena"] = {
        "qps": 3 * QUERIES_PER_STEP / (t_ms * 1e-3) * world, "kernel": f32_name, "kernel_ms": f32_ms,
        "achieved": f32_bytes / (f32_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": f32_bytes / (f32_ms * 1e-3) / 1e9 / peak, "frac_of_nominal_8000": f32_bytes / (f32_ms * 1e-3) / 1e9 / 8000.0,
        "algorithmic_bytes_per_launch": f32_bytes, "traffic": f32_tr, "launches_timed": f32_n}
    idx.local.set_mode(vl.Mode.Auto)

    # ---- end to end through the host API (host buffers in, host results out) -----------------------
    # One caller at a time (a lone client), and E2E_CALLERS concurrent callers, which is how the reference serves
    # searches (tokio workers under a read lock, client.rs:398) and how the reference arm is timed (one query per
    # host thread).  N = 1: callers on the handle (combined into batched launches).  N > 1: (a) one caller per rank
    # through the peer-memory exchange (one process per GPU), (b) rank 0 alone driving a shard group over all N GPUs
    # (vl_group_search: the reference's one-server-process deployment) with E2E_CALLERS native callers.
    def one_search(qi):
        if world == 1:
            idx.local.search_batch(queries[qi:qi + 1], k, metric)
        else:
            idx.search(queries[qi:qi + 1], k, metric)

    e2e_nq = 256   # queries per e2e step (one call each)

    def step_e2e():
        for qi in range(e2e_nq):
            one_search(qi)

    def time_e2e(step_fn):
        for _ in range(2):
            step_fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_fn()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return e2e_steps * e2e_nq / dt * world

    e2e_steps = max(4, args.steps // 2)
    e2e_single = time_e2e(step_e2e)
    e2e_qps, e2e_callers = e2e_single, 1
    e2e_python_callers, e2e_impl, e2e_single_python = None, "python threads (ctypes)", None
    e2e_extra = {}
    if world == 1:
        # E2E_CALLERS threads issue single-query searches back to back (a shared cursor); every call is still one
        # query in host memory → one result in host memory
        import itertools
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=E2E_CALLERS)

        def run_concurrent(total):
            cursor = itertools.count()

            def work(_):
                while True:
                    i = next(cursor)
                    if i >= total:
                        return
                    one_search(i % QUERIES_PER_STEP)
            list(pool.map(work, range(E2E_CALLERS)))
        run_concurrent(512)
        conc_total = max(args.steps, 8) * QUERIES_PER_STEP
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_concurrent(conc_total // 4)
        e2e_qps = (conc_total // 4) / (time.perf_counter() - t0)
        e2e_callers = E2E_CALLERS
        pool.shutdown()
        # the same calls from plain host threads (no interpreter lock between them): the headline e2e figure
        e2e_python_callers = e2e_qps
        want_ids, want_sc, _ = idx.local.search_batch(queries, k, metric)
        native_callers(idx.local, queries, k, metric, 0, E2E_CALLERS, 4096)
        nat = native_callers(idx.local, queries, k, metric, 0, E2E_CALLERS, conc_total)
        if nat is not None:
            assert np.array_equal(nat[1], want_ids) and np.array_equal(nat[2].view(np.uint64), want_sc.view(np.uint64))
            e2e_qps, e2e_impl = nat[0], "native host threads (scripts/native_callers.cpp)"
            # a lone native caller: the C-ABI latency without the Python wrapper's allocations
            e2e_single_python = e2e_single
            e2e_single = native_callers(idx.local, queries, k, metric, 0, 1, 4 * QUERIES_PER_STEP)[0]
            # more requests in flight than host cores (a server's worker pool): the combined batches grow with the
            # number of waiting callers, the device-side cost of a batch barely does (0.16 ms for 2 ... 128 queries)
            more = {}
            for nthreads in (64, 128):
                r_ = native_callers(idx.local, queries, k, metric, 0, nthreads, 2 * conc_total)
                more[str(nthreads)] = r_[0]
            e2e_extra["value_by_callers"] = {"16": e2e_qps, **more, "host_cores": cpu_threads()}
    else:
        e2e_extra["exchange_one_caller_per_rank"] = {
            "value": e2e_single, "unit": "queries/s x 1M-row shards",
            "api": "ShardedFlatIndex.search (one process per GPU, peer-memory exchange), one query per call"}
        # (b) one process, N GPUs: rank 0 builds its own shards on every device and serves E2E_CALLERS native callers
        host_barrier()
        if rank == 0 and torch.cuda.device_count() >= world:
            try:
                import ctypes as C
                L = vl.lib()
                shards = []
                for g in range(world):
                    sh = vl.FlatIndex(DIM, device=g)
                    sh.fill_synthetic(42, n_shard, first_row=g * n_shard, first_id=g * n_shard)
                    shards.append(sh)
                arr = (C.c_void_p * world)(*[sh.handle for sh in shards])
                grp = C.c_void_p()
                st = L.vl_group_create(arr, world, C.byref(grp))
                assert st == 0, "vl_group_create failed"
                native_callers_group(L, grp, queries, k, metric, E2E_CALLERS, 4096)          # mirrors, warm-up
                total = max(args.steps, 8) * QUERIES_PER_STEP
                nat = native_callers_group(L, grp, queries, k, metric, E2E_CALLERS, total)
                lone = native_callers_group(L, grp, queries, k, metric, 1, 2048)
                if nat is not None:
                    # the merged answers must be the sharded exchange's answers (checked against the oracle below)
                    assert np.array_equal(nat[1][:n_check], got_ids), "group search differs from the exchange path"
                    assert np.array_equal(nat[2][:n_check].view(np.uint64), got_sc.view(np.uint64))
                    e2e_extra["group_native_callers"] = {
                        "value": nat[0] * world, "unit": "queries/s x 1M-row shards", "global_qps": nat[0],
                        "callers": E2E_CALLERS, "single_caller_value": lone[0] * world,
                        "api": "vl_group_search (ONE process drives all N GPUs: the reference's single-server deployment), "
                               "one query per call from native host threads; concurrent callers are combined in front of "
                               "the shard fan-out (csrc/group.cpp)"}
                    e2e_qps, e2e_callers = nat[0] * world, E2E_CALLERS
                    e2e_impl = "native host threads on a shard group (vl_group_search)"
                L.vl_group_destroy(grp)
                for sh in shards:
                    sh.close()
            except Exception as e:  # noqa: BLE001 — keep the line
                e2e_extra["group_native_callers"] = {"error": repr(e)}
        host_barrier()

    # ---- BASELINE config 1: flat 10K x 384, one query (the reference's own intended size) -----------------
    c1_ids = None
    if rank == 0:
        try:
            small = vl.FlatIndex(DIM, device=local_rank)
            small.fill_synthetic(42, 10_000)
            small.set_pipelined(True)
            s_ids = torch.zeros((1, k), dtype=torch.int64, device=dev)
            s_sc = torch.zeros((1, k), dtype=torch.float64, device=dev)
            s_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            s_flg = torch.zeros(1, dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream().cuda_stream or 1

            def small_stream():
                for qi in range(QUERIES_PER_STEP):
                    small.search_device(d_queries[qi:qi + 1].data_ptr(), 1, k, metric, s_ids.data_ptr(), s_sc.data_ptr(), 0,
                                        s_cnt.data_ptr(), s_flg.data_ptr(), stream)
            small_stream()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                small_stream()
            e1.record()
            torch.cuda.synchronize()
            dev_us = e0.elapsed_time(e1) * 1e3 / (4 * QUERIES_PER_STEP)
            c1 = {"rows": 10_000, "dim": DIM, "k": k, "device_us_per_query_pipelined": dev_us}
            small.search_batch(queries[:2], k, metric)
            lone = native_callers(small, queries, k, metric, 0, 1, 4096)
            many = native_callers(small, queries, k, metric, 0, E2E_CALLERS, 8 * 4096)
            if lone is not None:
                c1.update({"latency_us_lone_caller": 1e6 / lone[0], "qps_lone_caller": lone[0],
                           f"qps_{E2E_CALLERS}_callers": many[0],
                           "api": "vl_index_search, host buffers in / host results out, one query per call"})
                c1_ids = lone[1]
            else:
                c1_ids = None
            small.close()
            e2e_extra["config1_flat_10k"] = c1
        except Exception as e:  # noqa: BLE001
            e2e_extra["config1_flat_10k"] = {"error": repr(e)}
            c1_ids = None
    host_barrier()

    # ---- extras: other metrics, batched (B=1024) tensor-core / CUDA-core pipelines ---------------------
    extras = {}
    if args.extras or world == 1:   # N = 1: always (a few seconds of GPU time); N > 1: only on request
        for name, mid in METRIC_NAMES.items():
            m2 = vl.SimilarityMetric(mid)

            def f():
                for qi in range(QUERIES_PER_STEP):
                    idx.search_device(d_queries[qi:qi + 1], k, m2)
            f()
            t_ms = timed(f, 3)
            extras[f"flat_b1_{name}_qps"] = 3 * QUERIES_PER_STEP / (t_ms * 1e-3)
        B = 1024
        bq = device_synth_rows(vl, 43, 1000, B, local_rank)
        d_bq = torch.from_numpy(bq).to(dev)
        tc_peak = None
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            tc_peak = json.load(open(pk)).get("bf16_tflops_sustained")
        for name, mid in METRIC_NAMES.items():
            m2 = vl.SimilarityMetric(mid)
            reps = 2 if name == "manhattan" else 10

            res = {}

            def g():
                res["r"] = idx.search_device(d_bq, k, m2)
            g()
            g()      # (pipelined handles alternate two scratch sets: both have been used before the timed reps)
            torch.cuda.synchronize()
            failed = int((res["r"][3] & 1).sum().item())
            t_ms = timed(g, reps) / reps
            qps = B / (t_ms * 1e-3) * world
            rec = {"qps": qps, "ms_per_batch": t_ms, "cert_failed_of_1024": failed}
            if name != "manhattan":
                tf = 2.0 * B * n_shard * DIM / (t_ms * 1e-3) / 1e12
                rec.update({"tflops_whole_pipeline": tf, "frac_of_measured_bf16_sustained": tf / tc_peak if tc_peak else None})
            else:
                rec.update({"lane_ops_per_s": 2.0 * B * n_shard * DIM / (t_ms * 1e-3)})
            extras[f"flat_b1024_{name}"] = rec
        extras["tc_cluster"] = int(os.environ.get("VL_TC_CLUSTER", "1"))
        extras["tc_cta_pairs"] = os.environ.get("VL_TC_PAIR", "default: cta_group::2 for even 128-query block counts")
        bt = extras.get(f"flat_b1024_{args.metric}")
        if bt and "tflops_whole_pipeline" in bt:
            roofline["batched_tensor"] = {"bound": "tensor", "workload": f"B=1024 x {n_shard} rows, {args.metric}, k={k}: whole pipeline "
                                          "(query conversion, 3 staged tcgen05 scans, selects, f64 rescore + certificate)",
                                          "achieved": bt["tflops_whole_pipeline"], "peak": tc_peak, "unit": "TFLOP/s",
                                          "frac": bt["frac_of_measured_bf16_sustained"], "ms_per_batch": bt["ms_per_batch"],
                                          "cert_failed_of_1024": bt["cert_failed_of_1024"]}

    # ---- clustered rows (1024-centre mixture): top-k gaps below the bf16 bound, the certificate levels at work -----
    if rank == 0 and world == 1:
        try:
            cl = vl.FlatIndex(DIM, device=local_rank)
            cl.fill_synthetic(42, n_shard, clusters=1024)
            cq = device_synth_rows(vl, 43, 0, 1024, local_rank, clusters=1024)
            keys = ("fast_queries", "exact_queries", "bf16_retries", "fp32_retries", "boosted_queries")
            cl.search_batch(cq[:2], k, metric)
            cl.search_batch(cq[:1], k, metric)
            s0 = cl.stats()
            lone = native_callers(cl, cq, k, metric, 0, 1, 2048)
            native_callers(cl, cq, k, metric, 0, E2E_CALLERS, 4096)   # warm-up like the main e2e leg (scratch at the larger K')
            many = native_callers(cl, cq, k, metric, 0, E2E_CALLERS, 16384)
            s1 = cl.stats()
            tb = []
            for _ in range(5):
                t0 = time.perf_counter()
                cl.search_batch(cq, k, metric)
                tb.append((time.perf_counter() - t0) * 1e3)
            s2 = cl.stats()
            ti = []
            bq_h = queries[:1024]
            for _ in range(5):
                t0 = time.perf_counter()
                idx.local.search_batch(bq_h, k, metric)
                ti.append((time.perf_counter() - t0) * 1e3)
            e2e_extra["clustered_1024"] = {
                "data": f"{n_shard} x {DIM}, 1024-centre mixture (top-10 / top-64 cosine gap ~0.003: below the bf16 bound)",
                "single_caller_value": lone[0] if lone else None, "value": many[0] if many else None,
                "callers": E2E_CALLERS, "unit": "queries/s",
                "single_query_stats": {x: s1[x] - s0[x] for x in keys},
                "batch_1024_ms_host_api": min(tb), "batch_1024_ms_host_api_iid_rows": min(ti),
                "batch_stats": {x: s2[x] - s1[x] for x in keys},
                "note": "queries whose base certificate fails are re-run together at a larger over-selection / on the "
                        "fp32 arena; exact_queries counts per-query exact scans (expected 0)"}
            cl.close()
        except Exception as e:  # noqa: BLE001
            e2e_extra["clustered_1024"] = {"error": repr(e)}

    # ---- wide rows (768-d: BERT-base-sized embeddings; the reference's DEFAULT_VECTOR_DIMENSION, lib.rs:142) -----------
    # rows wider than 384 elements: the tensor-core batch streams its query chunks through the TMA ring, single queries
    # scan the bf16 mirror.  Checked against the product's own exact f64 path (VL_MODE_EXACT) on 8 queries.
    if rank == 0 and world == 1 and args.wide_rows > 0:
        try:
            wd, wn = 768, args.wide_rows
            wi = ShardedFlatIndex(wd, rank=0, world=1, device=local_rank)
            wi.fill_synthetic(42, wn)
            wi.local.set_pipelined(True)
            src = vl.FlatIndex(wd, device=local_rank)
            src.fill_synthetic(43, 1024, first_row=1000)
            wq_h = np.ascontiguousarray(src.export()[1], dtype=np.float32)
            src.close()
            d_wq = torch.from_numpy(wq_h).to(dev)
            w0 = wi.local.stats()
            res = {}

            def wb():
                res["r"] = wi.search_device(d_wq, k, metric)
            wb(); wb()
            torch.cuda.synchronize()
            wfailed = int((res["r"][3] & 1).sum().item())
            ms_b = timed(wb, 10) / 10

            def ws():
                for qi in range(256):
                    wi.search_device(d_wq[qi:qi + 1], k, metric)
            ws()
            us_s = timed(ws, 2) / 2 / 256 * 1e3
            w1 = wi.local.stats()
            a_ids, a_sc, _ = wi.local.search_batch(wq_h[:8], k, metric)
            wi.local.set_mode(vl.Mode.Exact)
            x_ids, x_sc, _ = wi.local.search_batch(wq_h[:8], k, metric)
            e2e_extra["wide_rows_768"] = {
                "workload": f"flat {wn} x {wd} f32, {args.metric}, k={k}", "batch_1024_ms": ms_b,
                "batch_1024_tflops": 2.0 * 1024 * wn * wd / (ms_b * 1e-3) / 1e12, "batch_cert_failed_of_1024": wfailed,
                "tensor_queries": w1["tensor_queries"] - w0["tensor_queries"],
                "single_query_us_pipelined": us_s, "single_query_GBps_at_2B_per_element": wn * wd * 2 / (us_s * 1e-6) / 1e9,
                "bf16_mirror_scans": w1["bf16_scans"] - w0["bf16_scans"],
                "equals_exact_f64_path_on_8_queries": bool(np.array_equal(a_ids, x_ids) and
                                                           np.array_equal(a_sc.view(np.uint64), x_sc.view(np.uint64)))}
            wi.local.close()
        except Exception as e:  # noqa: BLE001
            e2e_extra["wide_rows_768"] = {"error": repr(e)}

    # ---- BASELINE config 5: 100M x 384 row-sharded, B = 1024, k = 100 (strong scaling over N) ------------------
    if args.config5_rows > 0:
        try:
            rows5 = args.config5_rows if world > 1 else args.config5_rows // 8
            per5 = rows5 // world
            idx.local.close()                      # free the 1M-row shard before the large one is filled
            i5 = ShardedFlatIndex(DIM, rank=rank, world=world, device=local_rank)
            i5.fill_synthetic(42, per5 * world)
            B5, k5 = 1024, 100
            q5 = device_synth_rows(vl, 43, 5000, B5, local_rank)
            probe_rows = [0, per5 - 1, (per5 * world) // 2 + 17, per5 * world - 1]
            for i, r in enumerate(probe_rows):      # stored rows must come back as their own top-1 (size-independent check)
                q5[i] = device_synth_rows(vl, 42, r, 1, local_rank)[0]
            d_q5 = torch.from_numpy(q5).to(dev)
            r5 = i5.search_device(d_q5, k5, metric)
            torch.cuda.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps5 = 5
            e0.record()
            for _ in range(reps5):
                r5 = i5.search_device(d_q5, k5, metric)
            e1.record()
            barrier()
            ms5 = e0.elapsed_time(e1) / reps5
            if world > 1:
                t = torch.tensor([ms5], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms5 = float(t.item())
            flg5 = r5[3].cpu().numpy()
            failed5 = int(((np.bitwise_or.reduce(flg5, axis=0) if flg5.ndim == 2 else flg5) & 1).sum())
            xfail5 = int(((np.bitwise_or.reduce(flg5, axis=0) if flg5.ndim == 2 else flg5) & 16).sum())
            t0 = time.perf_counter()
            h_ids5, h_sc5, h_cnt5 = i5.search(q5, k5, metric)       # host API: H2D, search, retries of failing queries, D2H
            host_ms5 = (time.perf_counter() - t0) * 1e3
            ok_probe = all(int(h_ids5[i, 0]) == r for i, r in enumerate(probe_rows))
            flops5 = 2.0 * B5 * per5 * world * DIM
            e2e_extra["config5"] = {
                "workload": f"flat {per5 * world} x {DIM} row-sharded over {world} GPU(s), cosine, B={B5}, k={k5}"
                            + ("" if world > 1 else " (ONE shard of the 8-GPU configuration: 100M rows do not fit one GPU)"),
                "rows_total": per5 * world, "rows_per_gpu": per5, "gpus": world, "scaling": "strong (total rows fixed)" if world > 1 else "n/a",
                "ms_per_batch_device": ms5, "qps": B5 / (ms5 * 1e-3), "tflops_per_gpu": flops5 / world / (ms5 * 1e-3) / 1e12,
                "cert_failed_of_1024": failed5, "exchange_timeouts": xfail5, "host_api_ms_per_batch": host_ms5,
                "host_api_queries_retried": i5.last_retried, "exact_queries": i5.local.stats()["exact_queries"],
                "stored_rows_query_back_to_themselves": bool(ok_probe),
                "exchange_bytes_per_rank_per_batch": int(vl.lib().vl_packed_result_bytes(B5, k5)),
                "exchange": ("peer-memory pushes + stamp-waiting merge kernel" if i5.exchange == "p2p" else "NCCL all-gather + merge kernel") if world > 1 else "none"}
            i5.local.close()
        except Exception as e:  # noqa: BLE001
            e2e_extra["config5"] = {"error": repr(e)}

    # ---- BASELINE config 4: HNSW profiles 8/16 and 32/64, ONE REPLICA PER GPU, queries split across the replicas ----
    if args.config4_rows > 0:
        try:
            n4, B4 = args.config4_rows, 4096
            cos = vl.SimilarityMetric.Cosine
            flat4 = vl.FlatIndex(DIM, device=local_rank)
            flat4.fill_synthetic(42, n4, clusters=1024)
            q4 = device_synth_rows(vl, 43, rank * B4, B4, local_rank, clusters=1024)     # this replica's share of the queries
            truth4, _, _ = flat4.search_batch(q4, k, cos)                                  # exact (certified) flat
            ids4, rows4 = flat4.export()
            flat4.close()
            c4 = {"rows": n4, "dim": DIM, "data": "synthetic 1024-centre mixture, unit norm", "ef_construction": args.hnsw_efc,
                  "k": k, "queries_per_replica": B4, "replicas": world, "ef": "k (the reference's setting), device beam = 8 x ef",
                  "parallelism": "replicas only: one full graph per GPU, queries split, no collective"}
            for prof, (M4, M04) in (("memory-optimized", (8, 16)), ("high-accuracy", (32, 64))):
                h4 = vl.HNSWIndex(DIM, cos, M=M4, M0=M04, ef_construction=args.hnsw_efc, device=local_rank)
                t0 = time.perf_counter()
                h4.add_batch(ids4, rows4)
                h4.build()
                build_s = time.perf_counter() - t0
                h4.search_batch(q4, k, cos, 0)
                host_barrier()
                t0 = time.perf_counter()
                reps4 = 3
                for _ in range(reps4):
                    gi4, _, gc4 = h4.search_batch(q4, k, cos, 0)
                dt4 = (time.perf_counter() - t0) / reps4
                hit4 = sum(len(set(map(int, gi4[i, :gc4[i]])) & set(map(int, truth4[i]))) for i in range(B4)) / (B4 * k)
                vis4 = h4.stats()["hnsw_visited"] / B4
                h4.close()
                if world > 1:
                    t = torch.tensor([dt4, -hit4, build_s], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt4, hit4, build_s = float(t[0]), -float(t[1]), float(t[2])
                c4[prof] = {"M": M4, "M0": M04, "qps_e2e_all_replicas": world * B4 / dt4, "qps_e2e_per_replica": B4 / dt4,
                            "recall_at_10_min_over_replicas": hit4, "visited_per_query": vis4,
                            "build_seconds_max_over_replicas": build_s}
            del rows4
            e2e_extra["config4_hnsw_replicas"] = c4
        except Exception as e:  # noqa: BLE001
            e2e_extra["config4_hnsw_replicas"] = {"error": repr(e)}

    # ---- CPU baseline + oracle comparison of the timed path ------------------------------------------------
    cpu = None
    oracle_check = None
    if not args.no_cpu_baseline:
        import oracle  # the checker / CPU baseline: the only leg of this arm that executes anything under oracle/
        oracle.build()
        threads = max(1, cpu_threads() // world)
        rows = synth_host_rows(oracle, 42, n_shard, DIM, threads, first_row=rank * n_shard)
        if world == 1:
            nq_cpu = ORACLE_CHECK_QUERIES      # ≈ 64 x 0.5 s of CPU at 1M rows, spread over the host threads
            qps_cpu, dt_cpu, cpu_ids = cpu_flat_qps(oracle, rows, queries[:nq_cpu], k, int(metric), threads)
            st, cpu_ids, cpu_sc = oracle.flat_search_batch(rows, None, queries[:nq_cpu], k, int(metric), nthreads=threads)
            assert st == 0
            assert np.array_equal(cpu_ids, got_ids), "GPU ids differ from the oracle"
            assert np.array_equal(cpu_sc.view(np.uint64), got_sc.view(np.uint64)), "GPU f64 scores differ from the oracle's bits"
            oracle_check = {"queries": int(nq_cpu), "ids": "identical", "scores": "bit-identical f64"}
            qps_1t, dt_1t, _ = cpu_flat_qps(oracle, rows, queries[:2], k, int(metric), 1)
            c1cpu, c1_oracle_ids = cpu_config1(oracle, threads, k, int(metric))
            if c1_ids is not None:
                assert np.array_equal(c1_ids[:c1_oracle_ids.shape[0]], c1_oracle_ids), "config 1: GPU ids differ from the oracle"
                c1cpu["gpu_ids_match_oracle"] = True
            cpu = {"value": qps_cpu, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": f"{nq_cpu} queries over the full {n_shard}-row store in {dt_cpu:.1f}s "
                             f"(one query per thread); single thread: {qps_1t:.2f} q/s",
                   "single_thread_qps": qps_1t, "config1_flat_10k": c1cpu}
        else:
            # distributed oracle: every rank scores ITS shard's rows on the host, rank 0 merges by (score desc, global
            # position asc) and compares with what the sharded search returned for the sampled queries
            st, l_ids, l_sc = oracle.flat_search_batch(rows, None, queries[:n_check], k, int(metric), nthreads=threads)
            assert st == 0
            l_ids = l_ids.astype(np.uint64) + np.uint64(rank * n_shard)
            gathered = [None] * world
            dist.all_gather_object(gathered, (l_ids, l_sc), group=host_group)
            if rank == 0:
                for qi in range(n_check):
                    cand = [(-float(sc[qi, j]), int(ids[qi, j])) for ids, sc in gathered for j in range(k)]
                    cand.sort()
                    want_i = [c[1] for c in cand[:k]]
                    want_s = np.array([-c[0] for c in cand[:k]], dtype=np.float64)
                    assert want_i == [int(x) for x in got_ids[qi]], f"sharded ids differ from the oracle (query {qi})"
                    assert np.array_equal(want_s.view(np.uint64), got_sc[qi].view(np.uint64)), f"sharded scores differ (query {qi})"
                oracle_check = {"queries": int(n_check), "ids": "identical", "scores": "bit-identical f64",
                                "how": "per-shard CPU oracle on every rank, merged on rank 0"}
        del rows

    # ---- HNSW section (replicas only: one full graph per GPU; measured on rank 0 at N = 1) ---------------
    hnsw = None
    if rank == 0 and world == 1 and args.hnsw_rows > 0:
        try:
            hnsw = hnsw_section(vl, args.hnsw_rows, args.hnsw_efc, local_rank)
        except Exception as e:  # noqa: BLE001 — the headline line must still be printed
            hnsw = {"error": repr(e)}
        try:   # the size the reference arm's CPU restatement is timed on (hnsw_reference_cpu): same data, same ef
            hnsw_small = hnsw_section(vl, 10_000, args.hnsw_efc, local_rank, nq=512)
            hnsw_small.pop("reference_restatement_recall", None)
            if isinstance(hnsw, dict):
                hnsw["same_size_as_reference_arm"] = hnsw_small
        except Exception as e:  # noqa: BLE001
            if isinstance(hnsw, dict):
                hnsw["same_size_as_reference_arm"] = {"error": repr(e)}

    if rank == 0:
        e2e = {"value": e2e_qps, "unit": "queries/s x 1M-row shards",
               "h2d_bytes_per_step": QUERIES_PER_STEP * DIM * 4,
               "d2h_bytes_per_step": QUERIES_PER_STEP * (k * 16 + 8),
               "callers": e2e_callers, "callers_impl": e2e_impl, "python_callers_value": e2e_python_callers,
               "single_caller_value": e2e_single, "python_single_caller_value": e2e_single_python,
               "api": "vl_index_search (host buffers in, host results out), one query per call; concurrent callers on a "
                      "handle are combined into batched launches by the handle (csrc/combiner.h)"}
        e2e.update(e2e_extra)
        line = {
            "metric": "flat_1m_384d_k10_qps", "value": value, "unit": "queries/s x 1M-row shards",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16 rows x f32 query, f32 accumulate; f64 rescore" if scan_elem_bytes == 2 else "f32; f64 rescore"),
            "data": "synthetic",
            "config": bench_config(args),
            "global_qps": qps_global,
            "parallelism": (f"row-sharded x{world}, " + ("single shard" if world == 1 else
                            "data plane: per-shard top-k pushed into peer HBM over NVLink by the finalize kernel + "
                            "stamp-waiting merge kernel (no collective call; NCCL only bootstraps and reduces the timings)"
                            if idx.exchange == "p2p" else "NCCL all-gather + merge kernel")),
            "exactness": "ids == oracle, f64 scores bit-identical (approximate scan over the "
                         + ("bf16 mirror of the rows" if scan_elem_bytes == 2 else "fp32 arena")
                         + " + f64 rescore of the over-selected candidates + optimality certificate)",
            "oracle_check": oracle_check,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches + (merges if (world > 1 and idx.exchange == "nccl") else 0)),
            "clocks": clocks,
            "extras": extras,
            "hnsw": hnsw,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
