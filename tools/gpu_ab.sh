#!/bin/bash
# A/B of a kernel-path switch on one box: tools/gpu_ab.sh ENVVAR [workload ...]   (experiments; prints ms per step)
VAR=$1; shift
for wl in "$@"; do
  for v in 0 1 0 1; do
    env $VAR=$v python bench.py --workload $wl --steps 20 --warmup 5 --skip-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl $VAR=$v', round(d['ms_per_step'],4), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
