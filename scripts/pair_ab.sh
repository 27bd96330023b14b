#!/bin/bash
# A/B of the batched tensor stage: CTA pairs (cta_group::2, the default for even query-block counts) vs single CTAs
# (VL_TC_PAIR=0), interleaved so that clock / power drift hits every arm alike; then one ncu --set full capture of
# the default kernel.  Run under gpurun.
for i in 1 2 3; do
  for arm in "VL_TC_PAIR=0" "VL_TC_PAIR=1"; do
    env $arm METRICS=${METRICS:-0} REPS=${REPS:-50} timeout 120 python scripts/batch_time.py 2>&1 | tail -n 1 | sed "s/^/$arm /"
  done
done
if [ "${NCU:-1}" = "1" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_scan_tc_kernel --launch-skip 9 -c 3 \
      -o gpurun_out/r02_prof_batch_tc_pair -f env REPS=2 python scripts/batch_time.py > gpurun_out/r02_ncu_batch_tc_pair.log 2>&1
  echo "ncu rc=$?"
fi
