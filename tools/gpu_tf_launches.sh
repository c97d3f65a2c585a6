#!/bin/bash
# per-kernel mean launch time of a transformer-mode step (ncu launch list, cold cache, serialised): tools/gpu_tf_launches.sh workload tag
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$2.csv python bench.py --workload $1 --steps 2 --warmup 1 --skip-extras > /dev/null 2>&1
python - <<EOP
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_$2.csv', errors='ignore')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: agg[d['Kernel Name'][:60]].append(float(d['Metric Value'].replace(',','')))
        except: pass
print("$1 $2")
for k,vv in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:12]:
    print(f"  {k:60s} n={len(vv):3d} mean={sum(vv)/len(vv)/1e3:8.1f}us")
EOP
