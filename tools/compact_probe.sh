#!/bin/bash
# Probe: run-to-convergence with / without active-set compaction at several batch sizes.
for b in 16384 131072 524288; do
  for f in "" "--compact"; do
    python bench.py --batch $b --steps 5 --warmup 3 --e2e-chunks 1 $f 2>/dev/null |
      python -c "import json,sys; d=json.load(sys.stdin)['converge']; print($b, '$f', 'ms', d['ms'], 'M traj-it/s', round(d['value']/1e6,2))"
  done
done
