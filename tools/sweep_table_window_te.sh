#!/bin/bash
# dev tool: window size of the twisted-Edwards tables per problem size
fmt() { grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['lg'], 'c', d['c'], 'K', d['K'], 'shared', d['shared'], 'total', d['total_ms'], 'dig', d['digits'], 'sort', d['sort'], 'acc', d['acc'], 'red', d['reduce'])
"; }
for C in 16 18 20; do
  echo "== ed table window $C"; MSM_B200_TABLE_WINDOW=$C python tools/perf_sweep.py --curve ed-on-bls12-377 --sizes 19,20,21,22,23 --reps 3 2>&1 | fmt
done
echo "== ed no tables"; MSM_B200_TABLES=0 python tools/perf_sweep.py --curve ed-on-bls12-377 --sizes 22,23 --reps 3 2>&1 | fmt
echo "== pallas default policy"; python tools/perf_sweep.py --curve pallas --sizes 18,20,21,22,24 --reps 3 2>&1 | fmt
echo "== pallas no tables"; MSM_B200_TABLES=0 python tools/perf_sweep.py --curve pallas --sizes 18,20,21,22,24 --reps 3 2>&1 | fmt
