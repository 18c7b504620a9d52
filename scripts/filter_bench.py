"""Band-pass kernel bandwidth: cfg2 shape [256,128,440] and larger batches; algorithmic bytes = 8*C*T per trial."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200 import ops
sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 4)
def run(B, C, T, layout, dtype, reps=20):
    xs = [torch.randn(B, C, T, device="cuda") for _ in range(max(2, int(300e6 / (B * C * T * 4)) + 1))]
    y = ops.sosfilt(xs[0], sos, out_layout=layout, out_dtype=dtype)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        ops.sosfilt(xs[r % len(xs)], sos, out_layout=layout, out_dtype=dtype, out=y)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    alg = 8.0 * B * C * T
    moved = B * C * T * (4 + (4 if dtype == torch.float32 else 2))
    return us, alg / us / 1e3, moved / us / 1e3
print("R =", os.environ.get("CSN_FILTER_R", "auto"))
for B in (256, 1024, 4096):
    for layout, dtype in (("TBC", torch.bfloat16), ("BCT", torch.float32)):
        us, alg_gbs, moved_gbs = run(B, 128, 440, layout, dtype)
        print(f"B={B:5d} {layout} {str(dtype)[6:]:9s}: {us:8.1f} us  algorithmic {alg_gbs:6.0f} GB/s ({alg_gbs/6548.8*100:4.1f}%)  moved {moved_gbs:6.0f} GB/s")
