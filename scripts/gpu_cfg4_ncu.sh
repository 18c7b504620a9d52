#!/bin/bash
# launch list of the cfg4 step (H = 512 cluster recurrence): per-kernel share of the step
mkdir -p gpurun_out
NSTEPS=1 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg4.csv python scripts/bench_cfg4.py > gpurun_out/ncu_cfg4.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_cfg4.csv') if l.startswith('"')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
agg=collections.OrderedDict(); tot=0
for row in rows:
    name=row['Kernel Name'][:70]; val=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    val = val/1000 if u=='ns' else (val*1000 if u=='ms' else val)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=val; tot+=val
print("launches", len(rows), "total us", round(tot))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]:
    print(f"{t:9.1f} us n={c:4d} avg={t/c:8.1f} {100*t/tot:5.1f}%  {k}")
PY
