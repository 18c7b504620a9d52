"""Small odd-shaped invocations of every kernel family in one process (a crash / sticky CUDA error canary; the pool
does not allow compute-sanitizer, so bounds are also checked by guard-band tests in tests/test_gpu_filter.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200 import _lib, ops, retrieval
torch.manual_seed(0)
sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 4)
# filter: fast path (32 | series, T % 4 == 0, tail chunk), general path, zero phase
for (B, C, T) in ((4, 16, 100), (3, 5, 37), (2, 16, 440)):
    x = torch.randn(B, C, T, device="cuda")
    for layout, dt in (("TBC", torch.bfloat16), ("BCT", torch.float32), ("TBC", torch.float32)):
        ops.sosfilt(x, sos, out_layout=layout, out_dtype=dt)
    ops.sosfilt(x, sos, zero_phase=True, out_layout="BCT", out_dtype=torch.float32)
# one bf16 and one fp32 train step (fused projection, rowwarp loss, resident GEMMs, fused optimiser), graph replays
for dtype in (torch.bfloat16, torch.float32):
    model = csn.Model(16, 32, 2, 24, include_top=True, compute_dtype=dtype).cuda()
    crit = csn.DINOLoss(24, 1, 1.5, 0.22, 5, 10).cuda()
    step = csn.DistillTrainStep(model, crit, lr=1e-3, sos=sos)
    e = torch.randn(6, 16, 52, device="cuda"); f = torch.randn(6, 24, device="cuda")
    for _ in range(4):
        l = step.step(e, f, 0)
    print("step", dtype, float(l))
# large-hidden path (per-step GEMMs, PDL, split slabs)
model = csn.Model(16, 136, 1, 24, include_top=False, compute_dtype=torch.bfloat16).cuda()
out = model(torch.randn(5, 9, 16, device="cuda")); out.sum().backward(); print("H=136 ok", float(out.abs().mean()))
# cluster recurrence (H = 256 / 512): one and two trial groups per cluster, ragged last group, with / without d_hseq
for (T, B, I, H) in ((9, 5, 16, 256), (7, 130, 32, 512), (40, 20, 64, 512)):
    model = csn.Model(I, H, 2, 24, include_top=False, compute_dtype=torch.bfloat16).cuda()
    out = model(torch.randn(B, T, I, device="cuda")); out.sum().backward()
    print("cluster H=%d B=%d ok" % (H, B), float(out.abs().mean()))
# multi-crop loss (staged cluster kernel) at a small K, canonical and reference modes
for K in (2048, 4096):
    s = torch.randn(4, 3, K, device="cuda"); t = torch.randn(2, 3, K, device="cuda")
    ops.dino_loss_fwd_bwd(s, t, torch.zeros(3 * K, device="cuda"), 0.1, 0.04, _lib.DINO_MULTICROP_REF)
    ops.dino_loss_fwd_bwd(s, t, torch.zeros(K, device="cuda"), 0.1, 0.04, _lib.DINO_MULTICROP_CANONICAL)
# alt losses, retrieval, inference
fd = csn.FeatureDistributionLoss(10, 1.5, 0.22, 5)
sx = torch.randn(5, 100, device="cuda", requires_grad=True); px = torch.randn(5, 7, device="cuda", requires_grad=True)
fd(sx, torch.randn(5, 100, device="cuda"), 1, torch.randint(0, 7, (5,), device="cuda"), pred_label=px).backward()
csn.CosineSimilarityLoss()(sx, torch.randn(5, 100, device="cuda")).backward()
D, I = retrieval.topk_search(torch.randn(301, 37, device="cuda"), torch.randn(33, 37, device="cuda"), 5)
m63 = csn.Model(63, 32, 1, 16, include_top=False).cuda().eval()
print("encode", tuple(m63.encode_trials(torch.randn(3, 63, 64, device="cuda"), sos=sos).shape), "topk", tuple(I.shape))
torch.cuda.synchronize()
print("sanitize_small done")
