"""Bring-up instrumentation: clock64 stamps of the first recurrence steps of CTA 0 (cfg2 shapes), forward and backward."""
import ctypes
import os
import statistics
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib, ops

T, B, I, H = 440, int(os.environ.get("PB", "256")), 128, 128
torch.manual_seed(0)
k = 1.0 / H ** 0.5
w = [((torch.rand(4 * H, I) * 2 - 1) * k).cuda(), ((torch.rand(4 * H, H) * 2 - 1) * k).cuda(),
     ((torch.rand(4 * H) * 2 - 1) * k).cuda(), ((torch.rand(4 * H) * 2 - 1) * k).cuda()]
x = torch.randn(T, B, I).cuda().bfloat16()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
grads = tuple(torch.empty_like(t) for t in w)
dh = torch.randn(B, H).cuda()
TRAIN = os.environ.get("PTRAIN", "1") == "1"
def run():
    h, r, ws = ops.lstm_layer_fwd(x, *w, torch.bfloat16, TRAIN)
    if TRAIN:
        ops.lstm_layer_bwd(x, w[0], w[1], h, r, ws, None, dh, grads, False, torch.bfloat16)
for _ in range(2):
    run()
_lib.call("csn_dbg_lstm_profile_buffer", ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("csn_dbg_lstm_profile_buffer", None)
ph = buf[1024:1040].cpu()
print("forward kernel phases (CTA 0): alloc %d cyc, weight staging %d cyc, time loop %d cyc; wall %d ns => %.0f MHz" % (
    int(ph[2] - ph[1]), int(ph[3] - ph[2]), int(ph[4] - ph[3]), int(ph[5] - ph[0]), 1e3 * float(ph[4] - ph[1]) / max(1.0, float(ph[5] - ph[0]))))
print("   cycles since kernel start at t=1,8,64,200,400:", [int(ph[6 + i] - ph[1]) for i in range(5)], "loop end:", int(ph[4] - ph[1]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
h, r, ws = ops.lstm_layer_fwd(x, *w, torch.bfloat16, True)
torch.cuda.synchronize(); e0.record(); h, r, ws = ops.lstm_layer_fwd(x, *w, torch.bfloat16, True); e1.record(); torch.cuda.synchronize()
rw = buf[1200:1200 + 128].view(64, 2).cpu()
print("bwd ring wait (cycles) steps 3..39:", [int(rw[i, 1] - rw[i, 0]) for i in range(3, 40)])
print("lstm_layer_fwd (cast + pack + GEMM + recurrence) event time: %.1f us" % (1e3 * e0.elapsed_time(e1)))
for name, p in (("forward", buf[:512].view(64, 8).cpu()), ("backward", buf[512:1024].view(64, 8).cpu())):
    print(name, "B =", B)
    print(" t | period | accwait->ld | ld->math | fence+arrive | arrive->issuer wake | issue+commit | commit->epi wake")
    rows = []
    for t in range(3, 40):
        period = int(p[t + 1, 0] - p[t, 0])
        a = int(p[t, 1] - p[t, 0]); b = int(p[t, 2] - p[t, 1]); c = int(p[t, 3] - p[t, 2])
        d = int(p[t + 1, 4] - p[t, 3]); e = int(p[t + 1, 5] - p[t + 1, 4]); f = int(p[t + 1, 0] - p[t + 1, 5])
        wa = int(p[t + 1, 0] - p[t + 1, 6])   # how long thread 0 actually waited on the accumulator barrier
        st = int(p[t, 7] - p[t, 3]) if name == "backward" else 0  # off-path stores after the arrive
        pre = int(p[t + 1, 6] - p[t, 7]) if name == "backward" else int(p[t + 1, 6] - p[t, 3])  # loop top -> before the acc wait
        rows.append((period, a, b, c, d, e, f, wa, st, pre))
        if t < 6:
            print(f"{t:3d} | {period:6d} | {a:6d} | {b:6d} | {c:6d} | {d:6d} | {e:6d} | {f:6d}")
    if name == "forward":
        print("fwd: accwake -> tmem_ld issued:", [int(p[t, 7] - p[t, 0]) for t in range(3, 20)], " issued -> ld done:", [int(p[t, 1] - p[t, 7]) for t in range(3, 20)])
    print("periods t=3..39:", [r[0] for r in rows])
    print("wait-on-acc   :", [r[7] for r in rows])
    print("median", [int(statistics.median(r[i] for r in rows)) for i in range(10)], "(last three: time actually spent waiting on bar_acc, off-path stores, ring wait + dh-independent math before the acc wait)")
    print("late steps (t >= 20) median", [int(statistics.median(r[i] for r in rows[17:])) for i in range(10)])
