"""Fused optimiser / EMA-teacher helpers for the autograd (LstmDistillation-shaped) training loop.

`FusedAdam` keeps ALL parameters of a module in one flat fp32 buffer (module parameters become views), the
gradients in a second flat buffer (each `p.grad` is a view, so autograd accumulates straight into it) and the Adam
moments in two more; `step()` is one libcsn_b200 kernel launch per parameter group, `all_reduce_grads()` is one
NCCL call.  Param groups follow utils/utils.py:636-647 (`get_params_groups`): biases and 1-D tensors are not
decayed.  Semantics are torch.optim.Adam / AdamW (LSTMDistill.py:322, LstmDistillation.py:469-471).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def get_params_groups(model):
    """utils/utils.py:636-647."""
    regularized, not_regularized = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if name.endswith(".bias") or len(param.shape) == 1:
            not_regularized.append(param)
        else:
            regularized.append(param)
    return [{"params": regularized}, {"params": not_regularized, "weight_decay": 0.}]


class FusedAdam:
    def __init__(self, params_or_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 grad_buffer=None):
        """grad_buffer: optional flat fp32 tensor (at least as long as the padded parameter count) that becomes the
        gradient buffer -- e.g. a peer-mapped symmetric-memory buffer the data-parallel exchange reduces in place."""
        groups = list(params_or_groups)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.decoupled = decoupled
        self.param_groups = []
        total = 0
        layout = []
        for g in groups:
            ps = [p for p in g["params"] if p.requires_grad]
            start = total
            for p in ps:
                layout.append((p, total))
                total += (p.numel() + 3) // 4 * 4
            grp = dict(self.defaults)
            grp.update({k: v for k, v in g.items() if k != "params"})
            grp.update(params=ps, start=start, stop=total)
            self.param_groups.append(grp)
        if total == 0:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = layout[0][0].device
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        if grad_buffer is not None:
            if grad_buffer.numel() < total or grad_buffer.dtype != torch.float32 or grad_buffer.device != dev:
                raise ValueError("FusedAdam: grad_buffer must be a float32 tensor of >= %d elements on %s" % (total, dev))
            self.flat_g = grad_buffer[:total]
            self.flat_g.zero_()
        else:
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.n_flat = total
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in layout:
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
        self.step_count = 0
        # per-parameter segments for clip_gradients (utils/utils.py:132-141)
        import ctypes as _C
        offs = [o for _, o in layout] + [total]
        self._seg_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._n_seg = len(layout)
        self._n_chunks = sum((offs[i + 1] - offs[i] + 2047) // 2048 for i in range(len(layout)))
        # squared norms [n_seg] followed by the per-chunk partials they are folded from (fixed order, no atomics)
        self._sumsq_buf = torch.zeros(len(layout) + self._n_chunks, dtype=torch.float32, device=dev)
        self._sumsq = self._sumsq_buf[:len(layout)]
        self._layout = layout
        # ---- fused clip + AdamW + EMA sweep (csn_fused_optim_step) ----
        seg_group = []
        for gi, g in enumerate(self.param_groups):
            seg_group += [gi] * len(g["params"])
        self._seg_group = torch.tensor(seg_group, dtype=torch.int32, device=dev)
        self._seg_active = torch.ones(len(layout), dtype=torch.int32, device=dev)
        self._seg_step = torch.zeros(len(layout), dtype=torch.int32, device=dev)
        self._group_of_seg = seg_group
        n_hyper = 2 * len(self.param_groups) + 1
        self._hyper_dev = torch.zeros(n_hyper, dtype=torch.float32, device=dev)
        self._fused_ws = None

    def clip_gradients(self, clip):
        """Per-parameter L2 clipping (reference utils.clip_gradients) on the device; returns the squared norms tensor
        (no host sync; take .sqrt() if the values are wanted)."""
        import ctypes as _C
        from . import _lib
        _lib.call("csn_clip_grad_segments", _C.c_void_p(self.flat_g.data_ptr()), _C.c_void_p(self._seg_off.data_ptr()),
                  self._n_seg, self._n_chunks, _C.c_void_p(self._sumsq_buf.data_ptr()), float(clip),
                  _C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return self._sumsq

    def set_group_active(self, index, active):
        """cancel_gradients_last_layer (utils/utils.py:144-149) sets p.grad = None for the last layer during the first
        epochs, which makes torch's optimiser SKIP those parameters (no moment update, no weight decay, no step
        count).  Put such parameters in their own param group and switch the group off for those steps."""
        self.param_groups[index]["active"] = bool(active)

    def upload_hyper(self, ema_momentum=0.0):
        """Send every group's current lr / weight_decay, the EMA momentum and the groups' active flags to the device (tiny
        async copies from rotating pinned slots, ordered on the current stream).  fused_step() does this itself unless it is
        told the values are already there (a captured CUDA graph of the step: call this before every replay)."""
        n_groups = len(self.param_groups)
        if not hasattr(self, "_hyper_slots"):
            pin = self._hyper_dev.device.type == "cuda"
            self._hyper_slots = [torch.zeros(2 * n_groups + 1, dtype=torch.float32).pin_memory() if pin
                                 else torch.zeros(2 * n_groups + 1) for _ in range(16)]
            self._hyper_events = [None] * 16
            self._hyper_n = 0
        k = self._hyper_n % len(self._hyper_slots)
        self._hyper_n += 1
        if self._hyper_events[k] is not None:
            self._hyper_events[k].synchronize()  # the copy that last read this slot is done (16 steps ago)
        h = self._hyper_slots[k]
        for gi, g in enumerate(self.param_groups):
            h[2 * gi] = float(g["lr"])
            h[2 * gi + 1] = float(g["weight_decay"])
        h[2 * n_groups] = float(ema_momentum)
        self._hyper_dev.copy_(h, non_blocking=True)
        if self._hyper_dev.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_events[k] = ev
        active = [1 if self.param_groups[gi].get("active", True) else 0 for gi in self._group_of_seg]
        if active != getattr(self, "_active_cached", None):
            self._seg_active.copy_(torch.tensor(active, dtype=torch.int32))
            self._active_cached = active

    def fused_step(self, clip=0.0, grad_scale=1.0, ema=None, ema_momentum=0.0, want_norms=False, upload=True):
        """clip_gradients + optimizer.step() + the EMA teacher update of LstmDistillation.py:607-619 as ONE sweep over the
        flat buffers (csn_fused_optim_step): per-parameter clip of the rank-averaged gradient (grad_scale = 1 / world),
        AdamW / Adam with each group's CURRENT lr / weight_decay (set them in param_groups as the reference loop does,
        :540-544), groups switched off by set_group_active() skipped like p.grad = None, teacher = m teacher + (1 - m)
        student for an EMATeacher `ema`.  No host sync; the hyper-parameters travel in a small pinned buffer, so a captured
        CUDA graph of the step stays valid while the schedules move.  Returns the squared norms [n_param] if asked."""
        import ctypes as _C
        from . import _lib
        n_groups = len(self.param_groups)
        if upload:
            self.upload_hyper(ema_momentum)
        if self._fused_ws is None:
            n = _C.c_size_t()
            _lib.call("csn_fused_optim_workspace_bytes", self._n_seg, self._n_chunks, _C.byref(n))
            self._fused_ws = torch.empty(n.value, dtype=torch.uint8, device=self.flat_p.device)
        betas, eps = self.defaults["betas"], self.defaults["eps"]
        vp = _C.c_void_p
        _lib.call("csn_fused_optim_step", vp(self.flat_p.data_ptr()), vp(self.flat_g.data_ptr()), vp(self.exp_avg.data_ptr()),
                  vp(self.exp_avg_sq.data_ptr()), vp(ema.teacher_flat.data_ptr()) if ema is not None else None,
                  vp(self._seg_off.data_ptr()), vp(self._seg_group.data_ptr()), vp(self._seg_active.data_ptr()),
                  vp(self._seg_step.data_ptr()), self._n_seg, self._n_chunks, vp(self._hyper_dev.data_ptr()), n_groups,
                  float(betas[0]), float(betas[1]), float(eps), int(self.decoupled), float(clip), float(grad_scale),
                  vp(self._sumsq_buf.data_ptr()) if want_norms else None, vp(self._fused_ws.data_ptr()),
                  vp(torch.cuda.current_stream().cuda_stream))
        self.step_count += 1
        return self._sumsq if want_norms else None

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()  # gradients stay views of the flat buffer (never set to None)

    def all_reduce_grads(self):
        """DDP semantics: mean over ranks (the division is folded into the Adam kernel's grad_scale)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g)
            return 1.0 / dist.get_world_size()
        return 1.0

    def step(self, grad_scale=1.0):
        self.step_count += 1
        for g in self.param_groups:
            a, b = g["start"], g["stop"]
            if b > a and g.get("active", True):
                g["step"] = g.get("step", 0) + 1  # per-group step count, like torch's per-parameter state["step"]
                ops.adam_step(self.flat_p[a:b], self.flat_g[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], g["lr"],
                              g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.decoupled,
                              g["step"], grad_scale)


class EMATeacher:
    """teacher <- m teacher + (1-m) student over flat buffers (LstmDistillation.py:616-619).  `student_flat` is the
    FusedAdam flat parameter buffer of the student; the teacher's parameters are re-pointed at a flat copy."""

    def __init__(self, teacher_module, student_optimizer: FusedAdam, student_module):
        s_params = [p for g in student_optimizer.param_groups for p in g["params"]]
        s_index = {id(p): i for i, p in enumerate(s_params)}
        named_s = dict(student_module.named_parameters())
        self.student_flat = student_optimizer.flat_p
        self.teacher_flat = torch.zeros_like(self.student_flat)
        # same ordering / offsets as the student's flat buffer, matched by parameter name
        offsets = {}
        o = 0
        for g in student_optimizer.param_groups:
            o = g["start"]
            for p in g["params"]:
                offsets[id(p)] = o
                o += (p.numel() + 3) // 4 * 4
        for name, tp in teacher_module.named_parameters():
            sp = named_s.get(name)
            if sp is None or id(sp) not in offsets:
                continue  # e.g. frozen last_layer.weight_g: not in the optimiser, stays as initialised
            off = offsets[id(sp)]
            view = self.teacher_flat[off:off + tp.numel()].view(tp.shape)
            view.copy_(tp.data)
            tp.data = view
            tp.requires_grad_(False)

    def update(self, momentum):
        ops.ema_update_(self.teacher_flat, self.student_flat, momentum)
