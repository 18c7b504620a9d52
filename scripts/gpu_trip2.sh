#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary2.txt; }
rm -f gpurun_out/summary2.txt
run t2_umma python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120
run t2_lstm python -m pytest tests/test_gpu_lstm.py -q -m gpu --timeout 200
run t2_step python -m pytest tests/test_gpu_step.py tests/test_gpu_filter.py tests/test_gpu_loss.py -q -m gpu --timeout 200
run prof_tmem python scripts/prof_lstm_steps.py
CSN_LSTM_W_SMEM=1 run prof_smem python scripts/prof_lstm_steps.py
run bench_tmem python bench.py --steps 10 --warmup 3 --no_cpu_baseline
CSN_LSTM_W_SMEM=1 run bench_smem python bench.py --steps 10 --warmup 3 --no_cpu_baseline
cat gpurun_out/summary2.txt; cat gpurun_out/prof_tmem.log gpurun_out/prof_smem.log
python - <<'PY'
import json
for n in ("bench_tmem","bench_smem"):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stages_ms"].items()})
    except Exception as e: print(n, "ERR", e)
PY
