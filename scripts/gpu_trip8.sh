#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary8.txt; }
rm -f gpurun_out/summary8.txt
run t8_lstm python -m pytest tests/test_gpu_lstm.py -q -m gpu --timeout 300 -x
run t8_rest python -m pytest tests/test_gpu_step.py tests/test_gpu_gemm.py tests/test_gpu_loss.py -q -m gpu --timeout 300
run bench8 python bench.py --steps 20 --warmup 5
run cfg4 python scripts/bench_cfg4.py
cat gpurun_out/summary8.txt; tail -n 15 gpurun_out/t8_lstm.log | cut -c1-200; tail -n 3 gpurun_out/t8_rest.log; cat gpurun_out/cfg4.log | tail -n 8
python - <<'PY'
import json
for n in ("bench8",):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, d["n_gpus"], round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", round(d["e2e"]["value"]))
        print("   roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "filter", d["roofline_filter"]["frac"], "loss", d.get("roofline_loss",{}).get("frac"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
