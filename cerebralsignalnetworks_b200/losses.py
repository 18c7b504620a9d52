"""The distillation losses LstmDistillFromDinoV2Train.py runs today, same constructors and
forward signatures, forward + backward in one libcsn_b200 kernel each (warp per batch row).

    criterion_feature_dist = FeatureDistributionLoss(nepochs=EPOCHS, warmup_teacher_temp=..., teacher_temp=...,
                                                     warmup_teacher_temp_epochs=...)          (Train.py:330-333)
    loss = criterion_feature_dist(lstm_output, image_features, EPOCH, label, pred_label=cls_pred)   (Train.py:371)
    criterion = CosineSimilarityLoss(); loss = criterion(lstm_output, image_features)         (Train.py:335,370)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .ops import _p, _stream, call


class HyperParams:
    """The class-level knobs of LstmDistillFromDinoV2Train.py:14-23 that the losses read (and, for T, write)."""
    learning_rate = 0.001
    T = 0.5
    soft_target_loss_weight = 0.25
    ce_loss_weight = 0.75
    warmup_teacher_temp = 1.5
    teacher_temp = 0.22
    warmup_teacher_temp_epochs = 50
    alpha = 0.5
    beta = 0.5


class _FeatureDistFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, pred, label, temperature, alpha, beta):
        B, K = student.shape
        loss = torch.empty((), dtype=torch.float32, device=student.device)
        d_student = torch.empty_like(student)
        d_pred = torch.empty_like(pred) if pred is not None else None
        call("csn_feature_dist_loss_fwd_bwd", _p(student), _p(teacher), _p(pred), _p(label), _p(loss), _p(d_student),
             _p(d_pred), B, K, 0 if pred is None else pred.shape[1], float(temperature), float(alpha), float(beta), 1.0,
             _p(ops.loss_workspace(B, student.device)), _stream())
        ctx.has_pred = pred is not None
        ctx.save_for_backward(d_student, d_pred) if ctx.has_pred else ctx.save_for_backward(d_student)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        saved = ctx.saved_tensors
        g = dloss.contiguous().to(torch.float32)
        ds = ops.scale_(saved[0], g)
        dp = ops.scale_(saved[1], g) if ctx.has_pred else None
        return ds, None, dp, None, None, None, None


class FeatureDistributionLoss(nn.Module):
    """alpha * CE(pred_label, label) + beta * F.cross_entropy(softmax(teacher / T), softmax(student / T)) with the
    teacher temperature warm-up schedule, exactly as LstmDistillFromDinoV2Train.py:107-140 writes it (the teacher
    probabilities act as logits, the student probabilities as the soft target)."""

    def __init__(self, nepochs, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs):
        super().__init__()
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp,
        ))

    def forward(self, student_outputs, teacher_outputs, epoch, label, pred_label=None):
        _lib.require_gpu()
        HyperParams.T = self.teacher_temp_schedule[epoch]  # the reference stores it on the class (:122)
        if pred_label is None:
            # F.cross_entropy(None, label) raises in the reference; the feature term alone is still useful here
            pred, lab = None, None
        else:
            pred = pred_label.float().contiguous()
            lab = label.to(device=pred.device, dtype=torch.int64).contiguous()
        return _FeatureDistFunction.apply(student_outputs.float().contiguous(), teacher_outputs.detach().float().contiguous(),
                                          pred, lab, float(HyperParams.T), HyperParams.alpha, HyperParams.beta)


class _CosineFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, eps):
        B, K = student.shape
        loss = torch.empty((), dtype=torch.float32, device=student.device)
        d_student = torch.empty_like(student)
        call("csn_cosine_loss_fwd_bwd", _p(student), _p(teacher), _p(loss), _p(d_student), B, K, float(eps), 1.0,
             _p(ops.loss_workspace(B, student.device)), _stream())
        ctx.save_for_backward(d_student)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (d_student,) = ctx.saved_tensors
        return ops.scale_(d_student, dloss.contiguous().to(torch.float32)), None, None


class CosineSimilarityLoss(nn.Module):
    """1 - nn.CosineSimilarity()(student, teacher).mean()  (LstmDistillFromDinoV2Train.py:36-43)."""

    def __init__(self):
        super().__init__()
        self.eps = 1e-8

    def forward(self, student_outputs, teacher_outputs):
        _lib.require_gpu()
        return _CosineFunction.apply(student_outputs.float().contiguous(), teacher_outputs.detach().float().contiguous(), self.eps)


class Parameters:
    """The class-level knobs of LSTMDistillRetreival.py:33-37 (`loss_fn_kd` reads alpha / temperature and the weights)."""
    ce_loss_weight = 0.50
    soft_target_loss_weight = 0.50
    alpha = 1
    temperature = 2


class _KDFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, label, temperature, c_kl, c_sl1, c_ce):
        B, K = student.shape
        loss = torch.empty((), dtype=torch.float32, device=student.device)
        d_student = torch.empty_like(student)
        call("csn_kd_loss_fwd_bwd", _p(student), _p(teacher), _p(label), _p(loss), _p(d_student), B, K, float(temperature),
             float(c_kl), float(c_sl1), float(c_ce), 1.0, _p(ops.loss_workspace(B, student.device)), _stream())
        ctx.save_for_backward(d_student)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (d_student,) = ctx.saved_tensors
        return ops.scale_(d_student, dloss.contiguous().to(torch.float32)), None, None, None, None, None, None


def _kd_inputs(student_logits, teacher_logits):
    _lib.require_gpu()
    s = student_logits.float().contiguous()
    t = teacher_logits.detach().float().contiguous()
    if s.dim() != 2 or s.shape != t.shape:
        raise ValueError("loss_fn_kd: student / teacher logits must both be [B, K] (got %s, %s)" % (tuple(s.shape), tuple(t.shape)))
    return s, t


def loss_fn_kd(student_logits, labels, teacher_logits, params=Parameters):
    """LSTMDistillRetreival.py:40-70: soft_target_loss_weight * T^2 / B * sum q (log q - log p) + ce_loss_weight *
    F.smooth_l1_loss(student, teacher), q = softmax(teacher / T), p = softmax(student / T).  `labels` is unused there too."""
    s, t = _kd_inputs(student_logits, teacher_logits)
    B, K = s.shape
    T = float(params.temperature)
    return _KDFunction.apply(s, t, None, T, params.soft_target_loss_weight * T * T / B, params.ce_loss_weight / (B * K), 0.0)


def loss_fn_kd_hinton(outputs, labels, teacher_outputs, params=Parameters):
    """LstmDistillFromDinoV2TrainSpampinato.py:107-121: nn.KLDivLoss()(log_softmax(outputs / T), softmax(teacher / T)) *
    (alpha T^2) + F.cross_entropy(outputs, labels) * (1 - alpha).  KLDivLoss's default reduction is the mean over ALL
    B*K elements, kept as written."""
    s, t = _kd_inputs(outputs, teacher_outputs)
    B, K = s.shape
    T, alpha = float(params.temperature), float(params.alpha)
    lab = labels.to(device=s.device, dtype=torch.int64).contiguous() if alpha != 1.0 else None
    return _KDFunction.apply(s, t, lab, T, alpha * T * T / (B * K), 0.0, (1.0 - alpha) / B)


def cosine_similarity_loss(v1, v2):
    """LSTMDistill.py:33-58 (what its loss_fn_kd returns, :60-98): -mean_b cos(v1_b, v2_b) with F.normalize's clamp
    (each norm at least 1e-12)."""
    _lib.require_gpu()
    return _CosineFunction.apply(v1.float().contiguous(), v2.detach().float().contiguous(), 1e-12) - 1.0
