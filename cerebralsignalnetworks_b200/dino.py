"""DINO projection head, DINO loss (teacher-temperature warm-up + centre EMA) and the multi-crop wrapper, with
the reference's constructor / forward signatures, running on libcsn_b200 kernels.

  DINOHead          LstmDistillation.py:65-99
  DINOLoss          single-view: LstmDistillFromDinoV2Train.py:45-105 ; multi-crop: LstmDistillation.py:101-159
  MultiCropWrapper  LstmDistillation.py:28-63 (= utils/utils.py:598-633)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, ops
from ._lib import ACT_GELU, ACT_NONE, DINO_MULTICROP_CANONICAL, DINO_MULTICROP_REF, DINO_SINGLE
from .functional import (ActFunction, BatchNormFunction, DINOLossFunction, L2NormFunction, LinearFunction,
                         WeightNormFunction)
from .lstm import Linear


class GELU(nn.Module):
    def forward(self, x):
        return ActFunction.apply(x, ACT_GELU)


class BatchNorm1d(nn.Module):
    """nn.BatchNorm1d(num_features) for [M, N] activations with the same parameters / buffers (state_dict compatible)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        if self.training:
            self.num_batches_tracked += 1
        return BatchNormFunction.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                                       self.training)


class WeightNormLinear(nn.Module):
    """nn.utils.weight_norm(nn.Linear(in, out, bias=False)) with the legacy parameter names weight_g / weight_v."""

    def __init__(self, in_features, out_features, compute_dtype=torch.float32):
        super().__init__()
        proto = nn.Linear(in_features, out_features, bias=False)
        v = proto.weight.detach().clone()
        self.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(v)
        self.compute_dtype = compute_dtype

    def forward(self, x):
        w = WeightNormFunction.apply(self.weight_v, self.weight_g)
        return LinearFunction.apply(x, w, None, ACT_NONE, self.compute_dtype)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256, compute_dtype=torch.float32):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = Linear(in_dim, bottleneck_dim, compute_dtype=compute_dtype)
        else:
            layers = [Linear(in_dim, hidden_dim, compute_dtype=compute_dtype)]
            if use_bn:
                layers.append(BatchNorm1d(hidden_dim))
            layers.append(GELU())
            for _ in range(nlayers - 2):
                layers.append(Linear(hidden_dim, hidden_dim, compute_dtype=compute_dtype))
                if use_bn:
                    layers.append(BatchNorm1d(hidden_dim))
                layers.append(GELU())
            layers.append(Linear(hidden_dim, bottleneck_dim, compute_dtype=compute_dtype))
            self.mlp = nn.Sequential(*layers)
        for m in self.modules():  # _init_weights, LstmDistillation.py:89-93
            if isinstance(m, Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.last_layer = WeightNormLinear(bottleneck_dim, out_dim, compute_dtype=compute_dtype)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False

    def forward(self, x):
        x = self.mlp(x.contiguous())
        x = L2NormFunction.apply(x)
        return self.last_layer(x)


class MultiCropWrapper(nn.Module):
    def __init__(self, backbone, head):
        super().__init__()
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head

    def forward(self, x):
        if not isinstance(x, list):
            x = [x]
        idx_crops = torch.cumsum(torch.unique_consecutive(
            torch.tensor([inp.shape[-1] for inp in x]), return_counts=True)[1], 0)
        start_idx, outputs = 0, []
        for end_idx in idx_crops:
            _out = self.backbone(torch.cat(x[start_idx:end_idx]))
            if isinstance(_out, tuple):
                _out = _out[0]
            outputs.append(_out)
            start_idx = end_idx
        output = outputs[0] if len(outputs) == 1 else torch.cat(outputs)
        return self.head(output)


class DINOLoss(nn.Module):
    """Same constructor and forward(student_output, teacher_output, epoch) as the reference.

    2-D inputs [B, K]        -> the single-view loss of LstmDistillFromDinoV2Train.py:62-93.
    3-D inputs [V, B, K]     -> the multi-crop loss of LstmDistillation.py:118-147, including its behaviour of
                                skipping student view 0 and keeping a per-row centre [1, B, K] after the first
                                update.  `canonical=True` selects upstream DINO semantics instead.
    The centre all-reduce uses torch.distributed when a process group is initialised (NCCL on the GPU box);
    unlike the reference it also works without one (world size 1)."""

    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9, canonical=False):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.canonical = canonical
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp,
        ))

    def forward(self, student_output, teacher_output, epoch):
        _lib.require_gpu()
        temp = float(self.teacher_temp_schedule[epoch])
        if student_output.dim() == 2:
            mode = DINO_SINGLE
        elif self.canonical:
            mode = DINO_MULTICROP_CANONICAL
        else:
            mode = DINO_MULTICROP_REF
            if student_output.shape[0] != self.ncrops:
                raise ValueError("student_output has %d views, ncrops=%d" % (student_output.shape[0], self.ncrops))
        stats = []
        loss = DINOLossFunction.apply(student_output.float(), teacher_output.detach().float(), self.center,
                                      self.student_temp, temp, mode, stats)
        self._update_center(stats[0], teacher_output, mode)
        return loss

    @torch.no_grad()
    def _update_center(self, batch_center, teacher_output, mode):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(batch_center)
        K = teacher_output.shape[-1]
        if mode == DINO_MULTICROP_REF:
            # torch.sum(dim=0) of the stacked [Vt, B, K] teacher, / (len(teacher_output) * world) -> [1, B, K]
            rows = teacher_output.shape[0]
            B = teacher_output.shape[1]
            if self.center.numel() == K:
                self.center = self.center.expand(1, B, K).contiguous() if self.center.dim() == 3 else \
                    self.center.reshape(1, 1, K).expand(1, B, K).contiguous()
            ops.center_ema(self.center, batch_center, self.center_momentum, 1.0 / (rows * world))
        else:
            rows = teacher_output.shape[0] if teacher_output.dim() == 2 else teacher_output.shape[0] * teacher_output.shape[1]
            ops.center_ema(self.center, batch_center, self.center_momentum, 1.0 / (rows * world))
