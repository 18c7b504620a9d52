#!/bin/bash
# persistent GEMM variant: parity, then the shapes it was written for with and without it
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py -x -q -k "gemm or 512 or 256" 2>&1 | tail -3
python - <<'PY'
import os, torch
from cerebralsignalnetworks_b200 import ops
def t(M,N,K,ta=False,tb=True,n=20,dt=torch.float32):
    a=torch.randn((K,M) if ta else (M,K),device="cuda").bfloat16(); b=torch.randn((N,K) if tb else (K,N),device="cuda").bfloat16()
    out=torch.empty(M,N,device="cuda",dtype=dt)
    for _ in range(3): ops.gemm_bf16(a,b,ta,tb,out=out) if dt==torch.float32 else ops.gemm_bf16(a,b,ta,tb,out_dtype=dt)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): ops.gemm_bf16(a,b,ta,tb,out=out) if dt==torch.float32 else ops.gemm_bf16(a,b,ta,tb,out_dtype=dt)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/n
    print(f"M={M} N={N} K={K} {dt}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s  out {M*N*out.element_size()/ms/1e6:.0f} GB/s  persistent={'off' if os.environ.get('CSN_GEMM_NO_PERSISTENT')=='1' else 'on'}")
t(56320,2048,128); t(56320,2048,512); t(384,65536,256); t(7040,2048,128)
PY
CSN_GEMM_NO_PERSISTENT=1 python - <<'PY'
import os, torch
from cerebralsignalnetworks_b200 import ops
def t(M,N,K,n=20):
    a=torch.randn(M,K,device="cuda").bfloat16(); b=torch.randn(N,K,device="cuda").bfloat16(); out=torch.empty(M,N,device="cuda")
    for _ in range(3): ops.gemm_bf16(a,b,False,True,out=out)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): ops.gemm_bf16(a,b,False,True,out=out)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/n
    print(f"M={M} N={N} K={K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s  persistent=off")
t(56320,2048,128); t(56320,2048,512); t(384,65536,256); t(7040,2048,128)
PY
timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1
CSN_GEMM_NO_PERSISTENT=1 timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1
