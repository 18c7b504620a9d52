"""The data-parallel DINO step of LstmDistillation.py:537-626 as ONE product object (BASELINE.json config 3).

    crops (2 global x 300 samples, 4 local x 200, one start per crop for the whole batch, :551-569)
    -> teacher(global crops) [no grad], student(all crops)  : MultiCropWrapper(Model(96, 128, 4, 128), DINOHead(128, K))
    -> DINOLoss multi-crop (reference behaviour incl. the per-row centre, :118-159)
    -> backward
    -> DDP gradient mean + centre all-reduce (:445, :154-156)         : two-shot exchange over NVLink peer memory; the
                                                                          DINO head's gradients (21.8 M of 22.3 M at
                                                                          K = 65536) and the centre statistics leave
                                                                          while BPTT still runs
    -> clip_gradients(3.0) + cancel_gradients_last_layer + AdamW(param groups, cosine lr / wd) + EMA teacher (:607-619)
                                                                        : ONE fused sweep (csn_fused_optim_step)

Everything arithmetic is a libcsn_b200 kernel (the modules' autograd bridges, functional.py); torch owns tensors,
streams and the autograd graph.  The three backbone passes of a step (teacher on the global crops, student on the global
crops, student on the local crops) are independent serial chains: each runs on its own stream with a share of the SMs
(csn_lstm_set_cta_budget).  Mixed precision: bf16 operands / fp32 accumulation and master weights, so the reference's fp16
GradScaler (:478-480, :606-613) has nothing to do here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, dp, ops
from ._lib import DINO_MULTICROP_REF
from .optim import EMATeacher, FusedAdam


def _split_params(student):
    """The reference's get_params_groups (utils/utils.py:636-647: biases and 1-D tensors are not decayed), with the DINO
    head first (its gradients are final first) and `last_layer` in a group of its own (cancel_gradients_last_layer)."""
    groups = {k: [] for k in ("head_reg", "head_last", "head_noreg", "bb_reg", "bb_noreg")}
    for name, p in student.named_parameters():
        if not p.requires_grad:
            continue
        noreg = name.endswith(".bias") or p.dim() == 1
        if name.startswith("head."):
            key = "head_last" if "last_layer" in name else ("head_noreg" if noreg else "head_reg")
        else:
            key = "bb_noreg" if noreg else "bb_reg"
        groups[key].append(p)
    order = ["head_reg", "head_last", "head_noreg", "bb_reg", "bb_noreg"]
    out = []
    for k in order:
        g = {"params": groups[k], "name": k}
        if k.endswith("noreg"):
            g["weight_decay"] = 0.0
        out.append(g)
    return out


class MultiCropDistillStep:
    def __init__(self, student, teacher, loss, lr_schedule, wd_schedule, momentum_schedule, clip_grad=3.0,
                 freeze_last_layer=1, betas=(0.9, 0.999), eps=1e-8, n_global=2, n_local=4, global_len=300, local_len=200,
                 batch_size=None, concurrent=True, cta_budget=(48, 48, 52, 74), seed=None, use_cuda_graph=True):
        """student / teacher: MultiCropWrapper(Model, DINOHead) on the GPU with identical architectures (the teacher is
        overwritten with the student's weights, LstmDistillation.py:446, and frozen); loss: DINOLoss(out_dim, n_global +
        n_local, ...); the three schedules are per-ITERATION arrays as built by utils.cosine_scheduler (:483-496).
        batch_size: trials per rank and step (needed up front when world > 1: the per-row centre statistics are part of the
        exchanged buffer)."""
        _lib.require_gpu()
        self.student, self.teacher, self.loss = student, teacher, loss
        self.lr_schedule, self.wd_schedule, self.momentum_schedule = lr_schedule, wd_schedule, momentum_schedule
        self.clip_grad, self.freeze_last_layer = float(clip_grad or 0.0), int(freeze_last_layer)
        self.n_global, self.n_local, self.global_len, self.local_len = n_global, n_local, global_len, local_len
        self.world = dp.world_size()
        self.rng = np.random.RandomState(seed) if seed is not None else np.random
        dev = next(student.parameters()).device
        self.device = dev
        teacher.load_state_dict(student.state_dict())
        for p in teacher.parameters():
            p.requires_grad = False
        groups = _split_params(student)
        self.K = loss.center.shape[-1]
        n_param = sum((p.numel() + 3) // 4 * 4 for g in groups for p in g["params"])
        self.n_head = sum((p.numel() + 3) // 4 * 4 for g in groups[:3] for p in g["params"])
        self.n_center = 0
        if self.world > 1:
            if batch_size is None:
                raise _lib.CsnError("MultiCropDistillStep: pass batch_size (trials per rank) when world > 1")
            self.n_center = int(batch_size) * self.K
        self.xchg = dp.TwoShotExchange(n_param + self.n_center, dev)
        self.opt = FusedAdam(groups, lr=float(lr_schedule[0]), betas=betas, eps=eps, weight_decay=float(wd_schedule[0]),
                             decoupled=True, grad_buffer=self.xchg.buf)
        assert self.opt.n_flat == n_param
        self.ema = EMATeacher(teacher, self.opt, student)
        self.center_sums = self.xchg.buf[n_param:n_param + self.n_center] if self.n_center else None
        # head gradients complete -> their exchange starts on a side stream while BPTT runs
        self._head_params = [p for g in groups[:3] for p in g["params"]]
        self._head_pending = 0
        self._side = torch.cuda.Stream(device=dev)
        self._head_done = torch.cuda.Event()
        if self.world > 1:
            for p in self._head_params:
                p.register_post_accumulate_grad_hook(self._head_grad_ready)
        self.concurrent = bool(concurrent)
        # SM shares of the three forward chains (teacher, student global, student local) and of each backward chain: a
        # recurrence launch picks its batch tile AND its side roles (Xp-server / dW-consumer CTAs) from its share, so the
        # chains of a phase are co-resident instead of queueing for SMs.  Default (48, 48, 52 forward = 148 SMs; 74 + 74
        # backward), measured at the cfg3 shape: 2.22 ms per step against 2.30 with (64, 64, 64, 64) and 2.98 without side
        # roles under a budget (scripts/gpu_cfg3b.sh, gpu_cfg3c.sh).
        if isinstance(cta_budget, (tuple, list)):
            self.cta_budget = tuple(int(b) for b in cta_budget)
        else:
            self.cta_budget = (int(cta_budget),) * 4
        if self.concurrent:
            # the student's two backbone passes run (and are differentiated) on their own streams on purpose
            try:
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
        if self.concurrent:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        self.it = 0
        # CUDA-graph path: the ~150 launches of a step (and the autograd bookkeeping around them: the eager step is bound
        # by the HOST, 4.4 ms of Python per 4.4 ms step at cfg3) are captured once and replayed.  What varies between
        # steps lives on the device: the crops are copied into static buffers, lr / wd / EMA momentum / frozen groups
        # travel through the optimiser's device arrays, the exchange epochs are device counters.  The teacher temperature
        # is a launch constant: a new epoch value captures a new graph.
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graphs = {}
        self._pool = None
        self._static = None
        self._eager_steps = 0

    # ------------------------------------------------------------------------------------------------ hooks
    def _head_grad_ready(self, _param):
        self._head_pending -= 1
        if self._head_pending == 0:
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.xchg.all_reduce_(0, self.n_head, flag_set=0)
                # the per-row centre statistics (B x K floats: 16.8 MB at cfg3) were final when the loss kernel ended:
                # they travel now, under BPTT, instead of with the backbone gradients after it
                if self.n_center:
                    self.xchg.all_reduce_(self.opt.n_flat, self.n_center, flag_set=0)
                self._head_done.record(self._side)

    # ------------------------------------------------------------------------------------------------ pieces
    def make_crops(self, eeg_btc):
        """LstmDistillation.py:551-569: one random start per crop, shared by the whole batch; a crop that would run past
        the end is shifted back."""
        T = eeg_btc.shape[1]
        out = []
        for n, length in ((self.n_global, self.global_len), (self.n_local, self.local_len)):
            for _ in range(n):
                s = int(self.rng.randint(0, T))
                if s + length > T:
                    s -= s + length - T
                out.append(eeg_btc[:, s:s + length, :])
        return out[:self.n_global], out[self.n_global:]

    def _forward(self, gcat, lcat, B):
        student, teacher = self.student, self.teacher
        if self.concurrent:
            cur = torch.cuda.current_stream()
            s_t, s_g, s_l = self._streams
            for st in self._streams:
                st.wait_stream(cur)
            ops.set_lstm_cta_budget(self.cta_budget[0])
            with torch.cuda.stream(s_t), torch.no_grad():
                t_feat = teacher.backbone(gcat)
            ops.set_lstm_cta_budget(self.cta_budget[1])
            with torch.cuda.stream(s_g):
                f_g = student.backbone(gcat)
            ops.set_lstm_cta_budget(self.cta_budget[2])
            with torch.cuda.stream(s_l):
                f_l = student.backbone(lcat)
            ops.set_lstm_cta_budget(self.cta_budget[3])  # stays set through backward (the BPTT chains run side by side)
            for st in self._streams:
                cur.wait_stream(st)
        else:
            with torch.no_grad():
                t_feat = teacher.backbone(gcat)
            f_g, f_l = student.backbone(gcat), student.backbone(lcat)
        with torch.no_grad():
            t_out = teacher.head(t_feat).view(self.n_global, B, self.K)
        s_out = student.head(torch.cat([f_g, f_l])).view(self.n_global + self.n_local, B, self.K)
        return s_out, t_out

    # ------------------------------------------------------------------------------------------------ the step
    def _run(self, gcat, lcat, B, epoch):
        """Everything of one step that runs on the device (capturable): wait for the peers, forward, loss, backward,
        exchange, fused optimiser sweep, centre EMA.  gcat [n_global * B, global_len, C], lcat [n_local * B, local_len, C]."""
        opt, loss_mod = self.opt, self.loss
        self.xchg.wait_done(0)  # the peers are done with the previous step's buffer: gradients may be written again
        self.xchg.wait_done(1)
        opt.zero_grad()
        s_out, t_out = self._forward(gcat, lcat, B)
        # ---- loss (reference multi-crop semantics); the per-row centre statistics land in the exchanged buffer ----
        temp = float(loss_mod.teacher_temp_schedule[epoch])
        # The loss kernel emits dLoss/dstudent together with the loss: backward starts from the student logits with that
        # gradient (no autograd node for the loss, no pass over the 100 MB gradient to multiply it by d(loss)/d(loss) = 1)
        bc_dst = self.center_sums if self.n_center else None
        loss, d_student, bc = ops.dino_loss_fwd_bwd(s_out.detach().float().contiguous(), t_out.float().contiguous(),
                                                    loss_mod.center.contiguous(), loss_mod.student_temp, temp,
                                                    DINO_MULTICROP_REF, batch_center=bc_dst)
        stats = [bc]
        self._head_pending = len(self._head_params) if self.world > 1 else -1
        torch.autograd.backward(s_out, grad_tensors=d_student.view(s_out.shape).to(s_out.dtype))
        if self.concurrent:
            ops.set_lstm_cta_budget(0)
        # ---- exchange: the head and the centre statistics went out from the hook; the backbone gradients (2 MB) now ----
        if self.world > 1:
            self.xchg.all_reduce_(self.n_head, self.opt.n_flat - self.n_head, flag_set=1)
            torch.cuda.current_stream().wait_event(self._head_done)
        # ---- clip + AdamW + EMA teacher, one sweep (lr / wd / momentum are already on the device) ----
        opt.fused_step(clip=self.clip_grad, grad_scale=1.0 / self.world, ema=self.ema, upload=False)
        ops.center_ema(loss_mod.center, stats[0], loss_mod.center_momentum, 1.0 / (self.n_global * self.world))
        return loss

    def step(self, eeg_btc, epoch, it=None, crops=None):
        """eeg_btc: float32 [B, T, C] on the GPU (the DataLoader layout, utils/PerilsEEGDataset.py:569).  `it` indexes the
        per-iteration schedules (default: an internal counter).  Returns the loss (0-d device tensor, this rank's; on the
        CUDA-graph path it is the graph's static output: consume it before the next call)."""
        it = self.it if it is None else int(it)
        self.it = it + 1
        opt, loss_mod = self.opt, self.loss
        B, C = eeg_btc.shape[0], eeg_btc.shape[2]
        if self.n_center and B * self.K != self.n_center:
            raise _lib.CsnError("MultiCropDistillStep: batch size changed (%d trials, built for %d)" % (B, self.n_center // self.K))
        for i, g in enumerate(opt.param_groups):  # LstmDistillation.py:540-544
            g["lr"] = float(self.lr_schedule[it])
            if not g["name"].endswith("noreg"):
                g["weight_decay"] = float(self.wd_schedule[it])
        opt.set_group_active(1, epoch >= self.freeze_last_layer)  # cancel_gradients_last_layer (:610)
        opt.upload_hyper(float(self.momentum_schedule[it]))
        if loss_mod.center.numel() == self.K:  # the reference's centre becomes [1, B, K] at its first update
            loss_mod.center = loss_mod.center.reshape(1, 1, self.K).expand(1, B, self.K).contiguous()
        gv, lv = crops if crops is not None else self.make_crops(eeg_btc)
        graphed = self.use_cuda_graph and self._eager_steps >= 2
        if not graphed:
            self._eager_steps += 1
            return self._run(torch.cat(gv), torch.cat(lv), B, epoch)
        # crops into the graph's static inputs, then replay (capture on first use of this teacher temperature)
        if self._static is None or self._static[0].shape[0] != self.n_global * B:
            self._static = (torch.empty(self.n_global * B, self.global_len, C, device=self.device),
                            torch.empty(self.n_local * B, self.local_len, C, device=self.device))
            self._graphs.clear()
        gcat, lcat = self._static
        for k, v in enumerate(gv):
            gcat[k * B:(k + 1) * B].copy_(v)
        for k, v in enumerate(lv):
            lcat[k * B:(k + 1) * B].copy_(v)
        key = float(loss_mod.teacher_temp_schedule[epoch])
        hit = self._graphs.get(key)
        if hit is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            torch.cuda.current_stream().synchronize()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                out = self._run(gcat, lcat, B, epoch)
            hit = self._graphs[key] = (g, out)
        hit[0].replay()
        return hit[1]
