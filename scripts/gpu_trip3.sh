#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary3.txt; }
rm -f gpurun_out/summary3.txt
run t3_lstm python -m pytest tests/test_gpu_lstm.py tests/test_gpu_gemm.py -q -m gpu --timeout 200
run t3_step python -m pytest tests/test_gpu_step.py -q -m gpu --timeout 200
for ks in 1 2 4; do CSN_LSTM_KS_F=$ks run prof_ksf$ks python scripts/prof_lstm_steps.py; done
for ksb in 1 4 8; do CSN_LSTM_KS_B=$ksb run bench_ksb$ksb python bench.py --steps 10 --warmup 3 --no_cpu_baseline; done
CSN_LSTM_KS_F=4 run bench_ksf4 python bench.py --steps 10 --warmup 3 --no_cpu_baseline
CSN_LSTM_KS_F=4 CSN_LSTM_KS_B=4 run t3_lstm_ks4 python -m pytest tests/test_gpu_lstm.py -q -m gpu -k bf16 --timeout 200
cat gpurun_out/summary3.txt; tail -3 gpurun_out/prof_ksf1.log gpurun_out/prof_ksf2.log gpurun_out/prof_ksf4.log
python - <<'PY'
import json
for n in ("bench_ksb1","bench_ksb4","bench_ksb8","bench_ksf4"):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["value"]))
    except Exception as e: print(n, "ERR", e)
PY
