#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary7.txt; }
rm -f gpurun_out/summary7.txt
run t7_all python -m pytest tests -q -m gpu --timeout 300
run prof7 python scripts/prof_lstm_steps.py
CSN_LOSS_NO_STAGE=1 run loss_nostage python scripts/loss_bench.py
for kb in 70 100 140 200; do CSN_LOSS_STAGE_KB=$kb run loss_stage$kb python scripts/loss_bench.py; done
run bench7 python bench.py --steps 20 --warmup 5 --no_cpu_baseline
cat gpurun_out/summary7.txt; tail -n 4 gpurun_out/t7_all.log; grep "median" gpurun_out/prof7.log; for f in loss_nostage loss_stage70 loss_stage100 loss_stage140 loss_stage200; do echo "== $f"; grep -v "^env" gpurun_out/$f.log | tail -n 4; done
python - <<'PY'
import json
for n in ("bench7",):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, d["n_gpus"], round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", round(d["e2e"]["value"]))
        if "roofline" in d: print("   roofline", d["roofline"]["frac"], "filter", d["roofline_filter"]["frac"], "loss", d.get("roofline_loss",{}).get("frac"))
    except Exception as e: print(n, "ERR", e)
PY
