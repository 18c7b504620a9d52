#!/bin/bash
# 2-GPU box: DP bench (NCCL) + single-GPU checks
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary6.txt; }
rm -f gpurun_out/summary6.txt
run t6_lstm python -m pytest tests/test_gpu_lstm.py tests/test_gpu_step.py -q -m gpu --timeout 300
run prof6 python scripts/prof_lstm_steps.py
run bench6_n1 python bench.py --steps 20 --warmup 5 --no_cpu_baseline
run bench6_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --no_cpu_baseline
run bench6_ref python bench.py --impl reference --steps 3 --warmup 1
cat gpurun_out/summary6.txt; tail -n 3 gpurun_out/t6_lstm.log; grep "median" gpurun_out/prof6.log; tail -n 5 gpurun_out/bench6_n2.log | cut -c1-600
python - <<'PY'
import json
for n in ("bench6_n1","bench6_n2","bench6_ref"):
    try:
        l=[x for x in open(f"gpurun_out/{n}.log") if x.startswith("{")][-1]; d=json.loads(l)
        print(n, d["n_gpus"], round(d["value"]), "trials/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", round(d["e2e"]["value"]))
        if "roofline" in d: print("   roofline", d["roofline"]["frac"], "filter", d["roofline_filter"]["frac"], "loss", d.get("roofline_loss",{}).get("frac"))
    except Exception as e: print(n, "ERR", e)
PY
