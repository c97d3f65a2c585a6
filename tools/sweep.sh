#!/bin/bash
# Batch-size sweep on one GPU (BASELINE config 5): fixed-iteration throughput, classic mode.
#   bash tools/sweep.sh > gpurun_out/sweep.txt
for w in quadrotor_classic cartpole_classic; do
  for b in 1024 4096 16384 65536 262144 1048576; do
    python bench.py --workload $w --batch $b --steps 10 --warmup 3 --skip-extras 2> /tmp/sweep_err.txt |
      python -c "import json,sys; d=json.load(sys.stdin); print('$w', $b, 'ms/step', round(d['ms_per_step'],3), 'M traj-it/s', round(d['value']/1e6,2))" ||
      tail -2 /tmp/sweep_err.txt
  done
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
