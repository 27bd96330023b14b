#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU).  Every program first runs WITHOUT ncu; numbers printed under
# ncu are never bench values.
#   1. bench.py launch list (gpu__time_duration.sum, first 600 launches of the default command at 2 steps)
#   2. B = 1024 batched pipeline: launch list + ncu --set full of the main tensor-core stage
#   3. B = 1024 manhattan (CUDA-core tile kernel): ncu --set full of batch_scan_cc_kernel
set -x
python bench.py --steps 2 --warmup 3 --hnsw-rows 0 --config5-rows 0 --no-cpu-baseline > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --hnsw-rows 0 --config5-rows 0 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
REPS=5 python scripts/batch_time.py > gpurun_out/r02_batch_time_plain.json 2> gpurun_out/r02_batch_time_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_batch_tc.csv \
    env REPS=2 python scripts/batch_time.py > gpurun_out/r02_ncu_batch_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_scan_tc_kernel --launch-skip 11 -c 1 \
    -o gpurun_out/r02_prof_batch_tc -f env REPS=2 python scripts/batch_time.py > gpurun_out/r02_ncu_batch_tc.log 2>&1
REPS=2 METRICS=2 python scripts/batch_time.py > gpurun_out/r02_batch_time_l1_plain.json 2> gpurun_out/r02_batch_time_l1_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:batch_scan_cc_kernel --launch-skip 5 -c 1 \
    -o gpurun_out/r02_prof_batch_cc -f env REPS=1 METRICS=2 python scripts/batch_time.py > gpurun_out/r02_ncu_batch_cc.log 2>&1
ls -la gpurun_out/r02_*
