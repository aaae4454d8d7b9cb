#!/bin/bash
# dev tool: sweeps the cost-model knobs of the tree tail (engine.cuh run_affine_glv) over a few sizes
for E in 16 32; do for R in 1e6 2.5e6 1e7; do
  echo "== ELEMS=$E ROUND=$R"
  MSM_B200_FINISH_ELEMS=$E MSM_B200_FINISH_ROUND=$R python tools/perf_sweep.py --sizes 16,18,20,22 --reps 5 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['lg'], d['c'], 'rounds', d['rounds'], 'total', d['total_ms'], 'acc', d['acc'], 'hot', d['hot'], 'red', d['reduce'])
"
done; done
