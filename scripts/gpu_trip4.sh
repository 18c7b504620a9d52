#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary4.txt; }
rm -f gpurun_out/summary4.txt
run umma_bench python scripts/umma_bench.py
run t4_lstm python -m pytest tests/test_gpu_lstm.py tests/test_gpu_step.py -q -m gpu --timeout 200
for ks in 1 2 4; do CSN_LSTM_KS_F=$ks run prof4_ksf$ks python scripts/prof_lstm_steps.py; done
for ksb in 1 4 8; do CSN_LSTM_KS_B=$ksb run bench4_ksb$ksb python bench.py --steps 10 --warmup 3 --no_cpu_baseline; done
cat gpurun_out/summary4.txt; cat gpurun_out/umma_bench.log; for k in 1 2 4; do tail -n 2 gpurun_out/prof4_ksf$k.log; done
python - <<'PY'
import json
for n in ("bench4_ksb1","bench4_ksb4","bench4_ksb8"):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["value"]))
    except Exception as e: print(n, "ERR", e)
PY
