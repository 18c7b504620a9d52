"""CUDA-event time of one LSTM layer forward / backward at the cfg2 shape (env toggles: CSN_LSTM_NO_SERVERS, CSN_LSTM_NO_CONSUMERS, CSN_LSTM_NO_SIDE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import ops
T, B, I, H = int(os.environ.get("PT", "440")), int(os.environ.get("PB", "256")), int(os.environ.get("PI", "128")), int(os.environ.get("PH", "128"))
torch.manual_seed(0)
k = 1.0 / H ** 0.5
w = [((torch.rand(4 * H, I) * 2 - 1) * k).cuda(), ((torch.rand(4 * H, H) * 2 - 1) * k).cuda(),
     ((torch.rand(4 * H) * 2 - 1) * k).cuda(), ((torch.rand(4 * H) * 2 - 1) * k).cuda()]
xs = [torch.randn(T, B, I).cuda().bfloat16() for _ in range(4)]
grads = tuple(torch.empty_like(t) for t in w)
dh = torch.randn(B, H).cuda()
def timed(fn, n=10):
    ts, out = [], None
    for i in range(n):
        out = None  # (the caching allocator reuses the previous call's reserve / workspace blocks)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2000000)  # the host enqueues the whole call behind a spin: device time only
        a.record(); out = fn(i); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, out
for _ in range(2):
    h, r, ws = ops.lstm_layer_fwd(xs[0], *w, torch.bfloat16, True)
    ops.lstm_layer_bwd(xs[0], w[0], w[1], h, r, ws, None, dh, grads, False, torch.bfloat16)
torch.cuda.synchronize()
tf, (h, r, ws) = timed(lambda i: ops.lstm_layer_fwd(xs[i % 4], *w, torch.bfloat16, True))
tb, _ = timed(lambda i: ops.lstm_layer_bwd(xs[3], w[0], w[1], h, r, ws, None, dh, grads, False, torch.bfloat16))
tags = [k for k in ("CSN_LSTM_NO_SERVERS", "CSN_LSTM_NO_CONSUMERS", "CSN_LSTM_NO_SIDE") if os.environ.get(k) == "1"]
print(f"T={T} B={B} I={I} H={H} {tags or 'default'}: fwd {tf:.1f} us  bwd {tb:.1f} us")
