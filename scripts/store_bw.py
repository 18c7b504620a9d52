"""Per-SM global store bandwidth (csn_dbg_store_bw): how fast can the spare CTAs of a recurrence launch write Xp?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerebralsignalnetworks_b200 import _lib
per = 8 << 20
for n_cta in (1, 20, 148):
    dst = torch.empty(n_cta * per // 4, device="cuda")
    out = torch.zeros(n_cta, dtype=torch.int64, device="cuda")
    for mode, name in ((0, "STG contiguous"), (1, "STG 4 rows x 128 B"), (2, "TMA bulk 16 KB")):
        for _ in range(2):
            _lib.call("csn_dbg_store_bw", ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(out.data_ptr()), per, n_cta, mode, 4, None)
        torch.cuda.synchronize()
        cyc = out.float().mean().item()
        print(f"{n_cta:4d} CTAs  {name:20s}: {4 * per / cyc:6.1f} B/clk/SM  ({n_cta * 4 * per / cyc * 1.965:8.0f} GB/s total at 1965 MHz)")
