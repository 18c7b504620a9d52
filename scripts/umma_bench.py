"""tcgen05.mma cost model probe: cycles to ISSUE 32 MMAs (K=16 each) and cycles until they COMPLETE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib
out = torch.zeros(64, dtype=torch.int64, device="cuda")
def run(M, N, n_acc, a_mode, reps=6):
    _lib.call("csn_dbg_umma_bench", ctypes.c_void_p(out.data_ptr()), M, N, n_acc, a_mode, reps, None)
    torch.cuda.synchronize()
    o = out[:2 * reps].view(reps, 2).cpu()
    return int(o[-1, 0]), int(o[-1, 1])
print("M N n_acc a_mode(2=TMEM,0=SMEM) -> issue, complete cycles for 32 MMAs; per-MMA")
for a_mode in (2, 0):
    for M in (128, 64):
        for N in (16, 64, 128, 256):
            for n_acc in (1, 4, 8):
                if M == 64 and n_acc != 1: continue
                i, t = run(M, N, n_acc, a_mode)
                print(M, N, n_acc, a_mode, "->", i, t, round(t / 32, 1))
