"""One process drives G GPUs through a shard group (vl_group_search): 16 native callers, one query per call.
Env: G (GPUs), N (rows per shard), CALLERS; VL_COMBINE_SPIN_US / VL_GROUP_SPIN_US / VL_DISABLE_ZERO_COPY for A/B."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, vectorlite_b200 as vl
G = int(os.environ.get("G", 2)); n = int(os.environ.get("N", 1_000_000)); dim, k = 384, 10
L = vl.lib()
shards = []
for g in range(G):
    sh = vl.FlatIndex(dim, device=g); sh.fill_synthetic(42, n, first_row=g * n, first_id=g * n); shards.append(sh)
qi = vl.FlatIndex(dim, device=0); qi.fill_synthetic(43, 1024); q = np.ascontiguousarray(qi.export()[1], dtype=np.float32); qi.close()
arr = (C.c_void_p * G)(*[sh.handle for sh in shards]); grp = C.c_void_p()
assert L.vl_group_create(arr, G, C.byref(grp)) == 0
m = vl.SimilarityMetric.Cosine
bench.native_callers_group(L, grp, q, k, m, 16, 4096)
out = {}
for c in [int(x) for x in os.environ.get("CALLERS", "1,16").split(",")]:
    best = 0.0
    for _ in range(2):
        best = max(best, bench.native_callers_group(L, grp, q, k, m, c, 2048 if c == 1 else 16384)[0])
    out[str(c)] = round(best)
try:
    cpu_max = open("/sys/fs/cgroup/cpu.max").read().strip()
except Exception:
    cpu_max = None
print(json.dumps({"gpus": G, "global_qps_by_callers": out, "x_shards_16": out.get("16", 0) * G,
                  "spin_us": [os.environ.get("VL_COMBINE_SPIN_US", "default"), os.environ.get("VL_GROUP_SPIN_US", "default")],
                  "zero_copy": not os.environ.get("VL_DISABLE_ZERO_COPY"), "affinity_cpus": len(os.sched_getaffinity(0)),
                  "os_cpu_count": os.cpu_count(), "cgroup_cpu_max": cpu_max}))
