"""cfg3 (MultiCropDistillStep) on one GPU: step time, launches per step; run under ncu for the per-kernel time sum."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200 import _lib
from cerebralsignalnetworks_b200.schedules import cosine_scheduler
if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # torchrun: one rank per GPU, the step exchanges over peer memory
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
B, C, T, H, L, K = int(os.environ.get("PB", "64")), 96, 495, 128, 4, int(os.environ.get("PK", "65536"))
torch.manual_seed(43); np.random.seed(43)
def make():
    return csn.MultiCropWrapper(csn.Model(C, H, L, H, include_top=False, compute_dtype=torch.bfloat16),
                                csn.DINOHead(H, K, compute_dtype=torch.bfloat16)).cuda()
student, teacher = make(), make()
crit = csn.DINOLoss(K, 6, 0.04, 0.04, 30, 100).cuda()
n_it = 4096
step = csn.MultiCropDistillStep(student, teacher, crit, cosine_scheduler(5e-4 * B / 256, 1e-6, 1, n_it), cosine_scheduler(0.04, 0.4, 1, n_it),
                                cosine_scheduler(0.996, 1.0, 1, n_it), clip_grad=3.0, freeze_last_layer=0, batch_size=B,
                                concurrent=os.environ.get("CONCURRENT", "1") == "1", use_cuda_graph=os.environ.get("GRAPH", "1") == "1", cta_budget=tuple(int(v) for v in os.environ.get("CTA_BUDGET", "48,48,52,74").split(",")))
g = torch.Generator(device="cuda").manual_seed(1)
eeg = [torch.randn(B, T, C, device="cuda", generator=g) for _ in range(3)]
n = int(os.environ.get("NSTEPS", "10"))
for i in range(3):
    step.step(eeg[i % 3], epoch=0)
torch.cuda.synchronize()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(n):
    loss = step.step(eeg[i % 3], epoch=0)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
if int(os.environ.get("RANK", "0")) == 0:
    print(f"cfg3 B={B} K={K}: {ms:.3f} ms/step ({B / ms * 1e3:.0f} trials/s), host enqueue {1e3 * t_host / n:.3f} ms/step, "
      f"{(_lib.launch_count() - l0) // n} libcsn launches/step, loss {float(loss):.4f}")
