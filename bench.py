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
QUERIES_PER_STEP = 64
E2E_CALLERS = min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4))
# ^ concurrent host threads calling vl_index_search in the e2e leg (N = 1): as many as the reference arm uses
#   (one query per host thread)


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


def device_synth_rows(vl, seed, first_row, n, device):
    """[n, DIM] f32 synthetic rows `first_row …` of stream `seed` from the PRODUCT's own counter-based generator
    (vl_index_fill_synthetic; bit-identical to the oracle's synth_rows, tests/test_flat_gpu.py
    test_device_generator_matches_oracle), so the measured arm does not need oracle/ for its inputs."""
    src = vl.FlatIndex(DIM, device=device)
    try:
        src.fill_synthetic(seed, n, first_row=first_row)
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


def synth_host_rows(oracle, seed, n, dim, threads):
    """Host copy of the synthetic store (counter-based → chunks generated in parallel)."""
    out = np.empty((n, dim), dtype=np.float32)
    chunk = (n + threads - 1) // threads

    def work(t):
        lo, hi = t * chunk, min(n, (t + 1) * chunk)
        if lo < hi:
            out[lo:hi] = oracle.synth_rows(seed, lo, hi - lo, dim)
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
    sweep = {}
    for ef in (0, 16, 32, 64, 128, 256):
        h.search_batch(queries, k, metric, ef)      # warm-up at the timed batch size (same kernel variant)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            gi, gs, gc = h.search_batch(queries, k, metric, ef)
        dt = (time.perf_counter() - t0) / reps
        hit = sum(len(set(map(int, gi[i, :gc[i]])) & set(map(int, truth[i]))) for i in range(nq))
        sweep[str(ef)] = {"recall_at_10": hit / (nq * k), "qps_e2e": nq / dt,
                          "visited_per_query": h.stats()["hnsw_visited"] / nq, "beam": 8 * (ef or k)}
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
    for name in ("hnsw_reference_recall_n20000_c1024_M16.json", "hnsw_reference_recall_n20000_c0_M16.json",
                 "hnsw_reference_recall_n50000_c1024_M16.json"):
        pth = os.path.join(ROOT, "tests", "golden", name)
        if os.path.exists(pth):
            d = json.load(open(pth))
            ref[name] = {e: round(v["recall_at_10"], 4) for e, v in d["sweep"].items()}
            if "qps_4threads" in next(iter(d["sweep"].values())):
                # timed when the golden file was generated (build container's host cores, 4 threads) — not this box
                ref[name]["qps_4threads_when_generated"] = {e: round(v["qps_4threads"]) for e, v in d["sweep"].items()}
                ref[name]["visited_per_query"] = {e: round(v["visited_per_query"]) for e, v in d["sweep"].items()}
    h.close()
    return {"rows": n, "dim": DIM, "data": f"synthetic {clusters}-centre mixture, unit norm", "M": 16, "M0": 32,
            "ef_construction": efc, "k": k, "batch": nq, "build_seconds": build_s, "builder": build_info,
            "host_threads": cpu_threads(),
            "sweep": sweep, "single_query_latency_us_ef_k": single_us,
            "concurrent_single_query_callers": {"callers": E2E_CALLERS, "qps_e2e_ef_k": conc_qps, "callers_impl": conc_impl,
                                                "python_callers_qps": conc_py},
            "reference_restatement_recall": ref,
            "note": "reference recall = oracle restatement of crate hnsw 0.11 + u64-quantised functors at "
                    "efC=400 on smaller N (CPU build is single-threaded); parity at equal parameters is "
                    "asserted in tests/test_hnsw_gpu.py"}


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
    hnsw_ref = None
    if args.hnsw_rows > 0:
        hnsw_ref = hnsw_reference_cpu(oracle, threads)
    line = {
        "impl": "reference", "metric": "flat_1m_384d_k10_qps", "value": qps,
        "unit": "queries/s x 1M-row shards", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"flat {n}x{DIM} f32 per GPU shard, {args.metric}, k={args.k}, B=1 "
                               f"({QUERIES_PER_STEP} single-query searches per step)",
                   "reference_arm": f"CPU: f32 rows widened to f64, {nq} queries per step, one per host thread "
                                    "(bounded sample of the same workload)",
                   "note": "reference is Rust (no toolchain here): C++ oracle restatement of "
                           "flat.rs:98-119 + lib.rs:425-572, -O2 -ffp-contract=off"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{nq * args.steps} queries over the full {n}-row store",
                         "with_16B_text_clone_per_row_qps": qps_clone, "single_thread_qps": qps_1t},
        "e2e": {"value": qps, "unit": "queries/s x 1M-row shards", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time the other metrics / batch modes")
    ap.add_argument("--hnsw-rows", type=int, default=1_000_000, help="HNSW section size (0 = skip; rank 0, N=1 only)")
    ap.add_argument("--hnsw-efc", type=int, default=400, help="ef_construction (reference default: 400)")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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

    def step_device():
        # a stream of independent single-query searches; with N > 1 the per-query exchange (one packed
        # all-gather + merge) runs on a side stream and overlaps the next query's scan
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

    # ---- correctness gate before any number counts (rank-local flags + sampled oracle) ----------
    o_ids, o_sc, o_cnt, flg = idx.search_device(d_queries[:4], k, metric)
    torch.cuda.synchronize()
    assert int(flg.max()) == 0, "optimality certificate failed on the bench workload"
    got_ids = o_ids.cpu().numpy().copy()

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

    # ---- end to end through the host API (host buffers in, host results out) -----------------------
    # One caller at a time (a lone client), and — on a single shard — E2E_CALLERS concurrent callers on the same
    # handle, which is how the reference serves searches (tokio workers under a read lock, client.rs:398) and how
    # the reference arm is timed (one query per host thread): the host-side part of one search overlaps the scan
    # of another.  The sharded exchange is ordered per handle, so N > 1 keeps a single caller per rank.
    def one_search(qi):
        if world == 1:
            idx.local.search_batch(queries[qi:qi + 1], k, metric)
        else:
            idx.search(queries[qi:qi + 1], k, metric)

    def step_e2e():
        for qi in range(QUERIES_PER_STEP):
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
        return e2e_steps * QUERIES_PER_STEP / dt * world

    e2e_steps = max(2, args.steps // 4)
    e2e_single = time_e2e(step_e2e)
    e2e_qps, e2e_callers = e2e_single, 1
    e2e_python_callers, e2e_impl, e2e_single_python = None, "python threads (ctypes)", None
    if world == 1:
        # E2E_CALLERS threads issue the steps' single-query searches back to back (a shared cursor, no barrier
        # between steps); every call is still one query in host memory → one result in host memory
        import itertools
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=E2E_CALLERS)

        def run_concurrent(n_steps):
            total = n_steps * QUERIES_PER_STEP
            cursor = itertools.count()

            def work(_):
                while True:
                    i = next(cursor)
                    if i >= total:
                        return
                    one_search(i % QUERIES_PER_STEP)
            list(pool.map(work, range(E2E_CALLERS)))
        run_concurrent(2)
        conc_steps = max(e2e_steps, args.steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_concurrent(conc_steps)
        e2e_qps = conc_steps * QUERIES_PER_STEP / (time.perf_counter() - t0)
        e2e_callers = E2E_CALLERS
        pool.shutdown()
        # the same calls from plain host threads (no interpreter lock between them): the headline e2e figure
        e2e_python_callers = e2e_qps
        want_ids, want_sc, _ = idx.local.search_batch(queries[:QUERIES_PER_STEP], k, metric)
        native_callers(idx.local, queries[:QUERIES_PER_STEP], k, metric, 0, E2E_CALLERS, 4 * QUERIES_PER_STEP)
        nat = native_callers(idx.local, queries[:QUERIES_PER_STEP], k, metric, 0, E2E_CALLERS,
                             conc_steps * QUERIES_PER_STEP)
        if nat is not None:
            assert np.array_equal(nat[1], want_ids) and np.array_equal(nat[2].view(np.uint64), want_sc.view(np.uint64))
            e2e_qps, e2e_impl = nat[0], "native host threads (scripts/native_callers.cpp)"
            # a lone native caller: the C-ABI latency without the Python wrapper's allocations
            e2e_single_python = e2e_single
            e2e_single = native_callers(idx.local, queries[:QUERIES_PER_STEP], k, metric, 0, 1, 8 * QUERIES_PER_STEP)[0]

    # ---- extras: other metrics, batched (B=1024) tensor-core / CUDA-core pipelines ---------------------
    extras = {}
    if args.extras or world == 1:   # N = 1: always (a second of GPU time); N > 1: only on request
        for name, mid in METRIC_NAMES.items():
            m2 = vl.SimilarityMetric(mid)

            def f():
                for qi in range(QUERIES_PER_STEP):
                    idx.search_device(d_queries[qi:qi + 1], k, m2)
            f()
            t_ms = timed(f, 5)
            extras[f"flat_b1_{name}_qps"] = 5 * QUERIES_PER_STEP / (t_ms * 1e-3)
        # the same single-query search with scans pinned to the fp32 arena (VL_MODE_FP32): the 4 B/element headline
        idx.local.set_mode(vl.Mode.Fp32)
        f32_name, _, f32_bytes, f32_ms, f32_n, f32_tr = scan_roofline(metric)
        t_ms = timed(step_device, 5)
        extras["flat_b1_cosine_fp32_scan" if args.metric == "cosine" else f"flat_b1_{args.metric}_fp32_scan"] = {
            "qps": 5 * QUERIES_PER_STEP / (t_ms * 1e-3) * world, "kernel": f32_name, "kernel_ms": f32_ms,
            "achieved_gbs": f32_bytes / (f32_ms * 1e-3) / 1e9, "frac": f32_bytes / (f32_ms * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_launch": f32_bytes, "traffic": f32_tr}
        idx.local.set_mode(vl.Mode.Auto)
        B = 1024
        bq = device_synth_rows(vl, 43, 1000, B, local_rank)
        d_bq = torch.from_numpy(bq).to(dev)
        tc_peak = None
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            tc_peak = json.load(open(pk)).get("bf16_tflops_sustained")
        for name, mid in METRIC_NAMES.items():
            m2 = vl.SimilarityMetric(mid)
            reps = 2 if name == "manhattan" else 5

            res = {}

            def g():
                res["r"] = idx.search_device(d_bq, k, m2)
            g()
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

    # ---- CPU baseline + sampled oracle check (rank 0, N = 1 only) ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle  # the checker / CPU baseline: the only leg of this arm that executes anything under oracle/
        oracle.build()
        threads = cpu_threads()
        rows = synth_host_rows(oracle, 42, n_shard, DIM, threads)
        nq_cpu = max(3 * threads, 4)   # ≈ 20-25 s of CPU work at 1M rows (0.5 s per query per thread)
        qps_cpu, dt_cpu, cpu_ids = cpu_flat_qps(oracle, rows, queries[:nq_cpu], k, int(metric), threads)
        assert np.array_equal(cpu_ids[:4].astype(np.int64), got_ids[:4]), "GPU ids differ from the oracle"
        qps_1t, dt_1t, _ = cpu_flat_qps(oracle, rows, queries[:2], k, int(metric), 1)
        cpu = {"value": qps_cpu, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": f"{nq_cpu} queries over the full {n_shard}-row store in {dt_cpu:.1f}s "
                         f"(one query per thread); single thread: {qps_1t:.2f} q/s",
               "single_thread_qps": qps_1t}
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
        line = {
            "metric": "flat_1m_384d_k10_qps", "value": value, "unit": "queries/s x 1M-row shards",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16 rows x f32 query, f32 accumulate; f64 rescore" if scan_elem_bytes == 2 else "f32; f64 rescore"),
            "data": "synthetic",
            "config": {"workload": f"flat {n_shard}x{DIM} f32 per GPU shard, {args.metric}, k={k}, B=1 "
                                   f"({QUERIES_PER_STEP} single-query searches per step)",
                       "rows_total": n_total, "parallelism": (f"row-sharded x{world}, " + ("single shard" if world == 1 else
                                       "per-shard top-k pushed into peer HBM over NVLink by the finalize kernel + "
                                       "stamp-waiting merge kernel (no collective call)" if idx.exchange == "p2p"
                                       else "NCCL all-gather + merge kernel")),
                       "l2": "inputs larger than L2 (1.536 GB store vs 126 MB)",
                       "pipelining": "programmatic dependent launch between consecutive searches", "exactness":
                       "ids == oracle, f64 scores bit-identical (approximate scan over the "
                       + ("bf16 mirror of the rows" if scan_elem_bytes == 2 else "fp32 arena")
                       + " + f64 rescore of the over-selected candidates + optimality certificate)",
                       "scanned_copy": ("bf16 mirror, 2 B/element (SURVEY §8d); the fp32-arena scan is reported in "
                                        "extras.flat_b1_cosine_fp32_scan") if scan_elem_bytes == 2 else "fp32 arena"},
            "global_qps": qps_global,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": scan_kernel, "kernel_ms": scan_ms_avg, "launches_timed": scan_n,
                         "algorithmic_bytes_per_launch": algo_bytes, "scanned_copy_bytes_per_element": scan_elem_bytes,
                         "frac_of_nominal_8000": achieved / 8000.0},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_qps, "unit": "queries/s x 1M-row shards",
                    "h2d_bytes_per_step": QUERIES_PER_STEP * DIM * 4,
                    "d2h_bytes_per_step": QUERIES_PER_STEP * (k * 16 + 8),
                    "callers": e2e_callers, "callers_impl": e2e_impl, "python_callers_value": e2e_python_callers,
                    "single_caller_value": e2e_single, "python_single_caller_value": e2e_single_python,
                    "bf16_retries": idx.local.stats()["bf16_retries"],
                    "api": "vl_index_search (host buffers in, host results out), one query per call; concurrent callers on a "
                           "handle are combined into batched launches by the handle (csrc/api.cu flat_search)"},
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
