#!/bin/bash
# Batched (B = 1024) flat pipeline, cosine, 1M x 384: timing without a profiler, then the ncu launch list and one
# ncu --set full capture of the tensor-core stage kernels (the long launch is stage 2).  Run under gpurun.
set -e
REPS=5 python scripts/batch_time.py > gpurun_out/batch_time_plain.json 2> gpurun_out/batch_time_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_batch_tc.csv \
    env REPS=2 python scripts/batch_time.py > gpurun_out/ncu_batch_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_scan_tc_kernel --launch-skip 9 -c 3 \
    -o gpurun_out/prof_batch_tc -f env REPS=2 python scripts/batch_time.py > gpurun_out/ncu_batch_tc.log 2>&1
