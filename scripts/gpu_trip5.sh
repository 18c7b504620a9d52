#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary5.txt; }
rm -f gpurun_out/summary5.txt
run t5_all python -m pytest tests -q -m gpu --timeout 300
run prof5 python scripts/prof_lstm_steps.py
run bench5 python bench.py --steps 20 --warmup 5
run bench5_fp32 python bench.py --steps 3 --warmup 3 --precision fp32 --no_cpu_baseline
cat gpurun_out/summary5.txt; tail -n 5 gpurun_out/t5_all.log; grep -A5 "forward\|backward" gpurun_out/prof5.log | grep -v "^--"
python - <<'PY'
import json
for n in ("bench5","bench5_fp32"):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
        print("   roofline", d["roofline"]["frac"], "filter", d["roofline_filter"]["frac"], "loss", d.get("roofline_loss",{}).get("frac"), "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
