import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib
out = torch.zeros(64, dtype=torch.int64, device="cuda")
def run(M, N, n_acc, a_mode, reps=8):
    _lib.call("csn_dbg_umma_bench", ctypes.c_void_p(out.data_ptr()), M, N, n_acc, a_mode, reps, None)
    torch.cuda.synchronize()
    o = out[:2 * reps].view(reps, 2).cpu()
    return int(o[-1, 0]), int(o[-1, 1])
print("grid", os.environ.get("CSN_UMMA_BENCH_GRID", "1"))
for name, args in (("TS N16 1acc, B changes every MMA", (128, 16, 1, 2)), ("TS N16 4acc, B changes every MMA", (128, 16, 4, 2)),
                   ("TS N16 4acc, B shared by 4 MMAs", (128, 16, 4, 3)), ("SS N16", (128, 16, 1, 0))):
    i, t = run(*args)
    print(f"{name}: issue {i} complete {t} per-MMA {t/32:.1f}")
