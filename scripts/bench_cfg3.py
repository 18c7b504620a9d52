"""cfg 3 of BASELINE.json on one GPU: the LstmDistillation.py step (lines 537-626) -- 2 global (300-sample) + 4 local
(200-sample) crops of [64, 495, 96] trials, 4-layer LSTM (hidden 128) student/teacher backbones in MultiCropWrapper,
DINOHead with a 65536-d output, reference multi-crop DINO loss, AdamW with the get_params_groups split, per-parameter
gradient clipping, EMA teacher.  Everything runs on libcsn_b200 kernels through the autograd bridges."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200.optim import EMATeacher, FusedAdam, get_params_groups

B, C, T, H, K = int(os.environ.get("PB", "64")), 96, 495, 128, int(os.environ.get("PK", "65536"))
dtype = torch.bfloat16
torch.manual_seed(43); np.random.seed(43)
def make():
    return csn.MultiCropWrapper(csn.Model(C, H, 4, H, include_top=False, compute_dtype=dtype),
                                csn.DINOHead(H, K, compute_dtype=dtype)).cuda()
student, teacher = make(), make()
teacher.load_state_dict(student.state_dict())
for p in teacher.parameters():
    p.requires_grad = False
crit = csn.DINOLoss(K, 6, 0.04, 0.04, 30, 100).cuda()
opt = FusedAdam(get_params_groups(student), lr=5e-4, weight_decay=0.04, decoupled=True)
ema = EMATeacher(teacher, opt, student)
g = torch.Generator(device="cuda").manual_seed(1)
eeg = torch.randn(B, T, C, device="cuda", generator=g)

def crops():
    out = []
    for n, length in ((2, 300), (4, 200)):
        for _ in range(n):
            s = np.random.randint(0, T)
            if s + length > T:
                s -= s + length - T
            out.append(eeg[:, s:s + length, :].contiguous())
    return out[:2], out[2:]

BATCH_VIEWS = os.environ.get("BATCH_VIEWS", "1") == "1"
# Independent recurrences side by side: the teacher pass, the student's global crops and its local crops are three
# serial chains with nothing in common; each gets its own stream and a share of the SMs (autograd replays every
# backward node on its forward stream, so the two student BPTT chains overlap as well).
CONCURRENT = os.environ.get("CONCURRENT", "1") == "1"
BUDGET = int(os.environ.get("CTA_BUDGET", "64"))
if CONCURRENT:
    from cerebralsignalnetworks_b200 import ops as _ops
    _ops.set_lstm_cta_budget(BUDGET)
    s_t, s_g, s_l = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def step(epoch=0):
    gv, lv = crops()
    if BATCH_VIEWS and CONCURRENT:
        cur = torch.cuda.current_stream()
        gcat, lcat = torch.cat(gv), torch.cat(lv)
        for st in (s_t, s_g, s_l):
            st.wait_stream(cur)
        with torch.cuda.stream(s_t), torch.no_grad():
            t_feat = teacher.backbone(gcat)
        with torch.cuda.stream(s_g):
            f_g = student.backbone(gcat)
        with torch.cuda.stream(s_l):
            f_l = student.backbone(lcat)
        for st in (s_t, s_g, s_l):
            cur.wait_stream(st)
        with torch.no_grad():
            t_out = teacher.head(t_feat).view(2, B, K)
        s_out = student.head(torch.cat([f_g, f_l])).view(6, B, K)
    elif BATCH_VIEWS:
        # Same arithmetic as the reference loop (:581-589, one view per call), but crops of equal length share one
        # backbone pass (the LSTM treats trials independently): 3 backbone passes instead of 8.  (The reference's
        # MultiCropWrapper groups by the LAST dimension, which is the channel count for [B,T,C] EEG, so it cannot
        # group by length itself.)
        with torch.no_grad():
            t_out = teacher.head(teacher.backbone(torch.cat(gv))).view(2, B, K)
        feats = torch.cat([student.backbone(torch.cat(gv)), student.backbone(torch.cat(lv))])
        s_out = student.head(feats).view(6, B, K)
    else:
        with torch.no_grad():
            t_out = torch.stack([teacher(v) for v in gv], dim=0)
        s_out = torch.stack([student(v) for v in gv + lv], dim=0)
    loss = crit(s_out, t_out, epoch)
    opt.zero_grad()
    loss.backward()
    opt.clip_gradients(3.0)
    opt.step()
    ema.update(0.996)
    return loss

t0 = time.time()
for _ in range(2):
    l = step()
torch.cuda.synchronize()
print("warm-up s", round(time.time() - t0, 2), "loss", float(l), "center shape", tuple(crit.center.shape))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"cfg3 (L4 H128, head K={K}, 2x300+4x200 crops) B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} trials/s  loss {float(l):.4f}")
print("peak memory GB", round(torch.cuda.max_memory_allocated() / 1e9, 2))
