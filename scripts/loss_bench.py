"""DINO loss kernel bandwidth at the cfg3 shape (6 student + 2 teacher views, 64 trials, K = 65536) and at cfg2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib, ops
dev = "cuda"
def run(Vs, Vt, B, K, mode, reps=12):
    g = torch.Generator(device=dev).manual_seed(7)
    sets = [(torch.randn(Vs, B, K, device=dev, generator=g), torch.randn(Vt, B, K, device=dev, generator=g)) for _ in range(3)]
    rows = B if mode == _lib.DINO_MULTICROP_REF else 1
    cen = torch.zeros(rows * K, device=dev)
    bc = torch.zeros(rows * K, device=dev)
    s3 = [(s if Vs > 1 else s[0], t if Vt > 1 or mode != _lib.DINO_SINGLE else t[0]) for s, t in sets]
    for s_, t_ in s3:
        ops.dino_loss_fwd_bwd(s_, t_, cen, 0.1, 0.04, mode, batch_center=bc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        s_, t_ = s3[r % 3]
        ops.dino_loss_fwd_bwd(s_, t_, cen, 0.1, 0.04, mode, batch_center=bc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = (Vs + Vt + Vs) * 4.0 * K * B + (2 * 4.0 * B * K if rows > 1 else 0)
    return ms, nbytes / ms / 1e6
print("env", {k: v for k, v in os.environ.items() if k.startswith("CSN_")})
ms, gbs = run(6, 2, 64, 65536, _lib.DINO_MULTICROP_REF)
print(f"cfg3 multicrop K=65536 B=64: {ms*1e3:.1f} us  {gbs:.0f} GB/s  ({gbs/6548.8*100:.1f}% of 6548.8)")
ms, gbs = run(6, 2, 256, 65536, _lib.DINO_MULTICROP_REF, reps=6)
print(f"multicrop K=65536 B=256: {ms*1e3:.1f} us  {gbs:.0f} GB/s  ({gbs/6548.8*100:.1f}%)")
ms, gbs = run(1, 1, 256, 384, _lib.DINO_SINGLE)
print(f"cfg2 single K=384 B=256: {ms*1e3:.1f} us  {gbs:.0f} GB/s")
ms, gbs = run(1, 1, 4096, 8192, _lib.DINO_SINGLE)
print(f"single K=8192 B=4096: {ms*1e3:.1f} us  {gbs:.0f} GB/s  ({gbs/6548.8*100:.1f}%)")
