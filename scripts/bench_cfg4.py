"""cfg 4 of BASELINE.json on one GPU: stacked 2-layer LSTM hidden 512, ViT-B 768-d targets, 128 trials (= 1024 global / 8)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cerebralsignalnetworks_b200 as csn
B = int(os.environ.get("PB", "128"))
torch.manual_seed(43)
model = csn.Model(128, 512, 2, 768, include_top=False, compute_dtype=torch.bfloat16).cuda()
crit = csn.DINOLoss(768, 1, 1.5, 0.22, 50, 100).cuda()
step = csn.DistillTrainStep(model, crit, lr=1e-3, sos=csn.EEGFilters(1000.0).sos(5.0, 95.0, 4))
g = torch.Generator(device="cuda").manual_seed(1)
eeg = [torch.randn(B, 128, 440, device="cuda", generator=g) for _ in range(4)]
feats = [torch.randn(B, 768, device="cuda", generator=g) for _ in range(4)]
t0 = time.time()
for i in range(3):
    l = step.step(eeg[i % 4], feats[i % 4], 0)
torch.cuda.synchronize()
print("warmup+capture s", round(time.time() - t0, 2), "loss", float(l))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("NSTEPS", "10"))
e0.record()
for i in range(n):
    l = step.step(eeg[i % 4], feats[i % 4], 0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"cfg4 (L2 H512 K768) B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} trials/s  loss {float(l):.4f}")
