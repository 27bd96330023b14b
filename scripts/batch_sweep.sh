#!/bin/bash
# B = 1024 batched pipeline timing (scripts/batch_time.py) for a list of VL_BATCH_RATIO values.  Run under gpurun.
for r in ${RATIOS:-0 8 12}; do
  VL_BATCH_RATIO=$r REPS=20 METRICS=${METRICS:-0,1,3} python scripts/batch_time.py 2>&1 | tail -n 1 | sed "s/^/ratio=$r /"
done
