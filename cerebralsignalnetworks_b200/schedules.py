"""Host-side schedule / bookkeeping helpers of the LstmDistillation.py loop: the per-iteration
learning-rate, weight-decay and teacher-momentum schedules (utils/utils.py:187-198, used at LstmDistillation.py:483-496
and applied per iteration at :540-544), the last-layer gradient freeze (utils/utils.py:144-149, LstmDistillation.py:602)
and checkpoint save / restart (utils/utils.py:152-185, LstmDistillation.py:634-646).  Pure host logic: the kernels take
the scheduled values as launch arguments (FusedAdam group `lr` / `weight_decay`, EMATeacher.update(momentum))."""
from __future__ import annotations

import os

import numpy as np
import torch


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """Linear warm-up from start_warmup_value to base_value over warmup_epochs, then half a cosine down to final_value;
    one entry per iteration, `epochs * niter_per_ep` in total."""
    total = int(epochs * niter_per_ep)
    n_warm = int(warmup_epochs * niter_per_ep)
    warm = np.linspace(start_warmup_value, base_value, n_warm) if warmup_epochs > 0 else np.array([])
    n_cos = total - n_warm
    phase = np.pi * np.arange(n_cos) / max(n_cos, 1)
    sched = np.concatenate((warm, final_value + 0.5 * (base_value - final_value) * (1.0 + np.cos(phase))))
    if len(sched) != total:
        raise ValueError("schedule length %d != epochs * niter_per_ep = %d" % (len(sched), total))
    return sched


def apply_schedules(optimizer, it, lr_schedule, wd_schedule):
    """LstmDistillation.py:540-544: every group gets the iteration's lr; only the first (regularised) group its wd."""
    for i, group in enumerate(optimizer.param_groups):
        group["lr"] = float(lr_schedule[it])
        if i == 0:
            group["weight_decay"] = float(wd_schedule[it])


def cancel_gradients_last_layer(epoch, model, freeze_last_layer):
    """Drop the gradients of every parameter whose name contains "last_layer" during the first epochs."""
    if epoch >= freeze_last_layer:
        return
    for n, p in model.named_parameters():
        if "last_layer" in n:
            p.grad = None


def save_checkpoint(path, **objects):
    """torch.save of {name: obj.state_dict() if it has one else obj} (the save_dict of LstmDistillation.py:634-644)."""
    blob = {k: (v.state_dict() if hasattr(v, "state_dict") else v) for k, v in objects.items()}
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)


def restart_from_checkpoint(ckp_path, run_variables=None, **kwargs):
    """Load every `name=object` found in the checkpoint (strict=False where supported) and refresh run_variables in
    place; silently returns when the file does not exist, as the reference does."""
    if not os.path.isfile(ckp_path):
        return False
    checkpoint = torch.load(ckp_path, map_location="cpu", weights_only=False)
    for key, value in kwargs.items():
        if key in checkpoint and value is not None:
            try:
                value.load_state_dict(checkpoint[key], strict=False)
            except TypeError:
                value.load_state_dict(checkpoint[key])
    if run_variables is not None:
        for name in run_variables:
            if name in checkpoint:
                run_variables[name] = checkpoint[name]
    return True
