#!/bin/bash
# dev tool: bits per level of the bucket reduction, shared-bucket mode (2^15 buckets) at n = 2^18
for G0 in 1 2 3; do for WG in 3 4 5; do for WM in 256 1024 4096 16384; do
  echo -n "GB0=$G0 WARP_GB=$WG WARP_MIN=$WM  "
  MSM_B200_REDUCE_GB0=$G0 MSM_B200_REDUCE_WARP_GB=$WG MSM_B200_REDUCE_WARP_MIN=$WM python tools/perf_sweep.py --sizes 18 --reps 5 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('total', d['total_ms'], 'red', d['reduce'], 'launches', d['launches'])
"
done; done; done
