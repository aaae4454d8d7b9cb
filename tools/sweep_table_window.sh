#!/bin/bash
# dev tool: window size of the tables (shared buckets) per problem size
fmt() { grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['lg'], 'c', d['c'], 'K', d['K'], 'shared', d['shared'], 'rounds', d['rounds'], 'total', d['total_ms'], 'dig', d['digits'], 'sort', d['sort'], 'acc', d['acc'], 'hot', d['hot'], 'red', d['reduce'])
"; }
CURVE=${1:-bls12-377}
for C in 16 19 22; do
  echo "== $CURVE table window $C"; MSM_B200_TABLE_WINDOW=$C python tools/perf_sweep.py --curve $CURVE --sizes 19,20,21,22,23,24 --reps 3 2>&1 | fmt
done
echo "== $CURVE no tables"; MSM_B200_TABLES=0 python tools/perf_sweep.py --curve $CURVE --sizes 19,20,21,22,23,24 --reps 3 2>&1 | fmt
