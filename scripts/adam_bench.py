"""Fused optimiser kernel (one rank) against the three-launch path: device time per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import dp, ops
n, K = 181632, 384
dev = torch.device("cuda")
p = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
ex = dp.LocalExchange(n + K, dev); ex.buf.normal_()
center = torch.zeros(K, device=dev); step = torch.zeros(1, dtype=torch.int32, device=dev); consts = torch.zeros(2, device=dev)
def fused():
    ops.dp_adam_step_peer(p, m, v, ex.grad_ptrs, ex.flag_ptrs, 1, 0, center, 0.9, 1.0 / 256, step, ex.ticket, 1e-3)
def three():
    ops.adam_step_graph(p, ex.buf[:n], m, v, step, consts, 1e-3)
    ops.center_ema(center, ex.buf[n:], 0.9, 1.0 / 256)
for name, fn in (("fused", fused), ("three launches", three)):
    for _ in range(5): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / 20:.2f} us per call (graph of 20)")
