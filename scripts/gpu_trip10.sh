#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t10 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_step.py tests/test_gpu_gemm.py -q -m gpu --timeout 300
tail -n 3 gpurun_out/t10.log
run prof10 python scripts/prof_lstm_steps.py; grep median gpurun_out/prof10.log
run bench10 python bench.py --steps 20 --warmup 5 --no_cpu_baseline
python - <<'PY'
import json
for n in ("bench10",):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, d["n_gpus"], round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", round(d["e2e"]["value"]))
    except Exception as e: print(n, "ERR", e)
PY
