#!/bin/bash
# dev tool: bucket-reduction knobs for MANY buckets: 2^21 shared buckets (2^24 points, c = 22) and the classic
# layout at 2^18 points (8 x 2^15 buckets)
fmt() { grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['lg'], 'c', d['c'], 'shared', d['shared'], 'total', d['total_ms'], 'red', d['reduce'])
"; }
for Q in 0 1; do for G in 2 3 4; do for WG in 3 5; do
  echo -n "2^24 Q0=$Q GB0=$G WARP_GB=$WG: "; MSM_B200_REDUCE_Q0=$Q MSM_B200_REDUCE_GB0=$G MSM_B200_REDUCE_WARP_GB=$WG python tools/perf_sweep.py --sizes 24 --reps 2 2>&1 | fmt
done; done; done
for Q in 0 1; do for G in 2 3 4; do for WM in 4096 65536; do
  echo -n "2^18 classic Q0=$Q GB0=$G WARP_MIN=$WM: "; MSM_B200_TABLES=0 MSM_B200_REDUCE_Q0=$Q MSM_B200_REDUCE_GB0=$G MSM_B200_REDUCE_WARP_MIN=$WM python tools/perf_sweep.py --sizes 18 --reps 5 2>&1 | fmt
done; done; done
