#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no_cpu_baseline"
$CMD > gpurun_out/plain_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches2.csv $CMD > gpurun_out/ncu_launches2.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches2.csv') if l.startswith('"')]
agg=collections.OrderedDict(); tot=0
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'][:60]; val=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    val = val/1000 if u=='ns' else (val*1000 if u=='ms' else val)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=val; tot+=val
print("total us", round(tot))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"{t:9.1f} us n={n:3d} avg={t/n:8.1f} {100*t/tot:5.1f}%  {k}")
PY
