#!/bin/bash
# B = 1024 batched pipeline timing (scripts/batch_time.py) for a list of stage schedules: "VL_BATCH_RATIO=r" (fixed
# geometric ratio, round-1/2 behaviour) or "VL_BATCH_STAGES=m" (balanced schedule with m filtered stages).  Run under gpurun.
for rep in 1 2; do
for arm in ${ARMS:-VL_BATCH_RATIO=16 VL_BATCH_STAGES=2 VL_BATCH_STAGES=3}; do
  env $arm REPS=${REPS:-30} METRICS=${METRICS:-0,1,3} python scripts/batch_time.py 2>&1 | tail -n 1 | sed "s/^/$arm /"
done
done
