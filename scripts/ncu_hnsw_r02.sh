#!/bin/bash
# Round-2 ncu captures of the HNSW search kernel (run under gpurun, one GPU): throughput mode (4096 queries, beam 80,
# one warp per query) with the bf16-mirror gathers and with fp32 gathers (before / after), and the lone-query kernel.
# Every program runs once WITHOUT ncu first.
set -x
N=1000000 EF=10 python scripts/hnsw_probe.py > gpurun_out/r02_hnsw_probe_plain.log 2>&1 || exit 1
N=1000000 EF=10 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:hnsw_search_kernel<\(int\)0, \(int\)3, \(bool\)0, \(int\)1, \(bool\)1' --launch-skip 1 -c 1 \
    -o gpurun_out/r02_prof_hnsw_tp_bf16 -f python scripts/hnsw_probe.py > gpurun_out/r02_ncu_hnsw_tp_bf16.log 2>&1
VL_HNSW_FP32_GATHER=1 N=1000000 EF=10 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:hnsw_search_kernel<\(int\)0, \(int\)3, \(bool\)0, \(int\)1, \(bool\)0' --launch-skip 1 -c 1 \
    -o gpurun_out/r02_prof_hnsw_tp_fp32 -f python scripts/hnsw_probe.py > gpurun_out/r02_ncu_hnsw_tp_fp32.log 2>&1
python scripts/hnsw_single_probe.py > gpurun_out/r02_hnsw_single_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:hnsw_search_kernel<\(int\)0, \(int\)3, \(bool\)0, \(int\)16, \(bool\)1' --launch-skip 4 -c 1 \
    -o gpurun_out/r02_prof_hnsw_single -f python scripts/hnsw_single_probe.py > gpurun_out/r02_ncu_hnsw_single.log 2>&1
# the reports are large (the full set over thousands of CTAs): keep the raw-page CSV (what profiles/summarize.py reads)
for r in gpurun_out/r02_prof_hnsw_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.csv 2>/dev/null; rm -f $r; done
ls -la gpurun_out/r02_prof_hnsw*
