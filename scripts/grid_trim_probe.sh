#!/bin/bash
# value (device-resident single-query cosine searches per second) for several scan-grid trims; run under gpurun
for t in 0 1 2 4 8; do
  VL_SCAN_GRID_TRIM=$t python bench.py --hnsw-rows 0 --no-cpu-baseline --steps 20 --warmup 3 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('trim', $t, 'value', round(d['value'],1), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), 'frac', round(d['roofline']['frac'],4), 'fp32', round(d['extras']['flat_b1_cosine_fp32_scan']['qps'],1))
"
done
