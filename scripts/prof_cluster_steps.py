"""Bring-up instrumentation of the cluster recurrence (H = 512): clock64 stamps of the first steps of CTA 0."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib, ops

T, B, I, H = 440, int(os.environ.get("PB", "128")), 128, int(os.environ.get("PH", "512"))
torch.manual_seed(0)
k = 1.0 / H ** 0.5
w = [((torch.rand(4 * H, I) * 2 - 1) * k).cuda(), ((torch.rand(4 * H, H) * 2 - 1) * k).cuda(),
     ((torch.rand(4 * H) * 2 - 1) * k).cuda(), ((torch.rand(4 * H) * 2 - 1) * k).cuda()]
x = torch.randn(T, B, I).cuda().bfloat16()
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.lstm_layer_fwd(x, *w, torch.bfloat16, True)
_lib.call("csn_dbg_lstm_profile_buffer", ctypes.c_void_p(buf.data_ptr()))
ops.lstm_layer_fwd(x, *w, torch.bfloat16, True)
torch.cuda.synchronize()
_lib.call("csn_dbg_lstm_profile_buffer", None)
p = buf[:1024].view(128, 8).cpu()
q = buf[1024:1536].view(128, 4).cpu()
names = ["period", "acc->ld done", "ld->arrive", "arrive->sender wake", "sender: copies issued", "arrive->issuer0 handoff seen",
         "handoff->grp0", "grp0->grp1", "grp1->commit", "commit->acc ready", "issuer1: handoff->grp2", "grp2->grp3", "grp3->commit"]
rows = []
for t in range(4, 100):
    rows.append((int(p[t + 1, 0] - p[t, 0]), int(p[t, 1] - p[t, 0]), int(p[t, 2] - p[t, 1]),
                 int(p[t + 1, 3] - p[t, 2]), int(p[t + 1, 4] - p[t, 2]), int(p[t + 1, 3] - p[t, 2]),
                 int(p[t + 1, 5] - p[t + 1, 3]), int(p[t + 1, 6] - p[t + 1, 5]), int(p[t + 1, 7] - p[t + 1, 6]),
                 int(p[t + 1, 0] - p[t + 1, 7]), int(q[t + 1, 0] - q[t + 1, 3]), int(q[t + 1, 1] - q[t + 1, 0]), int(q[t + 1, 2] - q[t + 1, 1])))
for i, n in enumerate(names):
    print(f"{n:32s} median {int(statistics.median(r[i] for r in rows)):6d}   first: {[r[i] for r in rows[:8]]}")
