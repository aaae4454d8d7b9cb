#!/bin/bash
# dev tool: shared-bucket tables on/off and the tree-tail knobs over a few sizes (BLS12-377)
fmt() { grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['lg'], 'c', d['c'], 'shared', d['shared'], 'rounds', d['rounds'], 'total', d['total_ms'], 'dig', d['digits'], 'sort', d['sort'], 'acc', d['acc'], 'hot', d['hot'], 'red', d['reduce'], 'launches', d['launches'])
"; }
echo "== tables off"; MSM_B200_TABLES=0 python tools/perf_sweep.py --sizes 16,18,20 --reps 5 2>&1 | fmt
echo "== tables on (default knobs)"; python tools/perf_sweep.py --sizes 16,18,20 --reps 5 2>&1 | fmt
for E in 8 16 32 64; do for R in 3e5 1e6 3e6; do
  echo "== tables on ELEMS=$E ROUND=$R"
  MSM_B200_FINISH_ELEMS=$E MSM_B200_FINISH_ROUND=$R python tools/perf_sweep.py --sizes 16,18,20 --reps 5 2>&1 | fmt
done; done
for G in 2 3 4; do
  echo "== tables on REDUCE_GB0=$G"; MSM_B200_REDUCE_GB0=$G python tools/perf_sweep.py --sizes 18 --reps 5 2>&1 | fmt
done
echo "== tables up to 2^22"; MSM_B200_TABLE_MAX_LOG2N=22 python tools/perf_sweep.py --sizes 21,22 --reps 3 2>&1 | fmt
echo "== no tables 2^21, 2^22"; MSM_B200_TABLES=0 python tools/perf_sweep.py --sizes 21,22 --reps 3 2>&1 | fmt
