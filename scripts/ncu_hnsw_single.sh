#!/bin/bash
# ncu --set full capture of the single-query (search mode, not construction) HNSW kernel; run under gpurun.
set -e
python scripts/hnsw_single_probe.py > gpurun_out/hnsw_single_plain.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:hnsw_search_kernel<\(int\)0, \(int\)3, \(bool\)0' --launch-skip 4 -c 2 \
    -o gpurun_out/prof_hnsw_single -f python scripts/hnsw_single_probe.py > gpurun_out/hnsw_single_ncu.log 2>&1
