#!/bin/bash
# The numbers quoted in DESIGN.md / README.md: the three bench lines and the launch lists of the transformer workloads.
for wl in quadrotor_classic quadrotor_tf cartpole_tf; do
  python bench.py --workload $wl > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_transformer.csv python bench.py --workload quadrotor_tf --steps 2 --warmup 1 --skip-extras > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_cartpole_tf.csv python bench.py --workload cartpole_tf --steps 2 --warmup 1 --skip-extras > /dev/null 2>&1
python - <<EOP
import json
for f in ["quadrotor_classic","quadrotor_tf","cartpole_tf"]:
    try:
        d=json.loads(open("gpurun_out/final_"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["kernel"][:20], d["roofline"]["frac"], d["roofline"].get("achieved"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        print("   ", {k:v["ms"] for k,v in d["kernels"].items()}, d.get("converge",{}).get("ms"))
    except Exception as e: print(f, "ERR", e)
EOP
