"""Oracle: LSTM encoder, DINO head, DINO loss, train step on torch-CPU fp32.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every class cites the reference lines it restates.  DINOLoss/DINOHead/MultiCropWrapper are
pinned against the reference's own classes by tests/test_oracle_vs_reference.py and
tests/golden/*.npz; `Model` is a restatement of a file the reference does not ship
(parity unpinned for that class only).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Model(nn.Module):
    """Restated `models.lstm.Model` (ABSENT from the reference tree).

    Authority, in order: constructor keywords / return arity at
    LstmDistillFromDinoV2Train.py:323,326,365, LstmDistillFromDinoV2TrainSpampinato.py:368,
    LstmDistillation.py:427-428; body from LSTMDistillRetreival.py:85-110
    (nn.LSTM(batch_first=True) -> [:, -1, :] -> fc) and LSTMDistill.py:112-142 (class head on
    the features, ReLU on the returned features).  The `x.view(B, C, T)` of those analogues is
    NOT followed: the distill scripts size input_size to the channel count and feed [B, T, C]."""

    def __init__(self, input_size=128, lstm_size=128, lstm_layers=1, output_size=128, include_top=True,
                 n_classes=40):
        super().__init__()
        self.input_size = input_size
        self.lstm_size = lstm_size
        self.lstm_layers = lstm_layers
        self.output_size = output_size
        self.include_top = include_top
        self.lstm = nn.LSTM(input_size, lstm_size, num_layers=lstm_layers, batch_first=True)
        self.output = nn.Linear(lstm_size, output_size)
        if include_top:
            self.classifier = nn.Linear(output_size, n_classes)
        # MultiCropWrapper assigns these (LstmDistillation.py:40)
        self.fc = nn.Identity()
        self.head = nn.Identity()

    def forward(self, x):
        out, _ = self.lstm(x)  # zero initial state (LSTMDistillRetreival.py:98-100)
        feat = self.output(out[:, -1, :])
        if self.include_top:
            cls = self.classifier(feat)
            return F.relu(feat), cls
        return feat


class DINOHead(nn.Module):
    """Restatement of LstmDistillation.py:65-99 (both the default use_bn=False path and the BatchNorm1d layers of use_bn=True,
    :72-80)."""

    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim)]
            if use_bn:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers.append(nn.GELU())
            for _ in range(nlayers - 2):
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                if use_bn:
                    layers.append(nn.BatchNorm1d(hidden_dim))
                layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.constant_(m.bias, 0)
        self.last_layer = nn.utils.weight_norm(nn.Linear(bottleneck_dim, out_dim, bias=False))
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False

    def forward(self, x):
        x = self.mlp(x)
        x = F.normalize(x, dim=-1, p=2)
        return self.last_layer(x)


def teacher_temp_schedule(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs):
    """LstmDistillFromDinoV2Train.py:56-60 / LstmDistillation.py:112-116."""
    return np.concatenate((
        np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
        np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp,
    ))


class DINOLossSingleView(nn.Module):
    """Restatement of LstmDistillFromDinoV2Train.py:45-105 with the all-reduce made explicit:
    `world_sum` is a callable that sums a tensor over ranks in place (identity for 1 rank)."""

    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9, world_sum=None, world_size=1):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = teacher_temp_schedule(warmup_teacher_temp, teacher_temp,
                                                           warmup_teacher_temp_epochs, nepochs)
        self.world_sum = world_sum
        self.world_size = world_size

    def forward(self, student_output, teacher_output, epoch):
        student_out = student_output / self.student_temp
        temp = self.teacher_temp_schedule[epoch]
        teacher_out = F.softmax((teacher_output - self.center) / temp, dim=-1)
        loss = torch.sum(-teacher_out * F.log_softmax(student_out, dim=-1), dim=-1)
        total_loss = loss.mean()
        self.update_center(teacher_output)
        return total_loss

    @torch.no_grad()
    def update_center(self, teacher_output):
        batch_center = torch.sum(teacher_output, dim=0, keepdim=True)
        if self.world_sum is not None:
            self.world_sum(batch_center)
        batch_center = batch_center / (len(teacher_output) * self.world_size)
        self.center = self.center * self.center_momentum + batch_center * (1 - self.center_momentum)


class DINOLossMultiCrop(DINOLossSingleView):
    """Restatement of LstmDistillation.py:101-159, called with stacked 3-D tensors
    student [ncrops, B, K], teacher [2, B, K] (LstmDistillation.py:591-593).  Reproduces the
    reference quirks: teacher `.chunk(1)` keeps both global views in ONE chunk, so student view 0
    is skipped and views 1.. are each matched against both teacher views; and
    `update_center` sums over dim 0 of the 3-D teacher, so `center` becomes [1, B, K] after the
    first step (Q3)."""

    def forward(self, student_output, teacher_output, epoch):
        student_out = student_output / self.student_temp
        student_out = student_out.chunk(self.ncrops)
        temp = self.teacher_temp_schedule[epoch]
        teacher_out = F.softmax((teacher_output - self.center) / temp, dim=-1)
        teacher_out = teacher_out.detach().chunk(1)
        total_loss = 0
        n_loss_terms = 0
        for iq, q in enumerate(teacher_out):
            for v in range(len(student_out)):
                if v == iq:
                    continue
                loss = torch.sum(-q * F.log_softmax(student_out[v], dim=-1), dim=-1)
                total_loss += loss.mean()
                n_loss_terms += 1
        total_loss /= n_loss_terms
        self.update_center(teacher_output)
        return total_loss


class MultiCropWrapper(nn.Module):
    """Restatement of LstmDistillation.py:28-63 (= utils/utils.py:598-633)."""

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
        start_idx, output = 0, torch.empty(0).to(x[0].device)
        for end_idx in idx_crops:
            _out = self.backbone(torch.cat(x[start_idx:end_idx]))
            if isinstance(_out, tuple):
                _out = _out[0]
            output = torch.cat((output, _out))
            start_idx = end_idx
        return self.head(output)


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """Restatement of utils/utils.py:187-198."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


def synthetic_batch(batch, channels=128, samples=440, feat_dim=384, n_classes=40, seed=43):
    """Synthetic Spampinato-shaped batch: EEG [B, C, T] from the reference's
    generator formula, teacher features N(0,1) [B, feat_dim], labels uniform in [0, n_classes)."""
    from .filters import synthetic_eeg

    eeg = synthetic_eeg(batch, channels, samples, seed=seed)
    rng = np.random.default_rng(seed + 1)
    feats = rng.normal(0.0, 1.0, size=(batch, feat_dim)).astype(np.float32)
    labels = rng.integers(0, n_classes, size=(batch,)).astype(np.int64)
    return eeg, feats, labels


class DistillStepOracle:
    """The north-star composition on the CPU: scipy-equivalent sosfilt
    (order-4 Butterworth 5-95 Hz) -> restated Model -> DINOLoss single-view -> backward ->
    torch.optim.Adam(lr).  Step order follows LstmDistillFromDinoV2Train.py:358-375."""

    def __init__(self, input_size=128, lstm_size=128, lstm_layers=1, output_size=384, include_top=False,
                 nepochs=100, lr=1e-3, low_hz=5.0, high_hz=95.0, fs=1000.0, order=4, zero_phase=False,
                 warmup_teacher_temp=1.5, teacher_temp=0.22, warmup_teacher_temp_epochs=50, seed=43,
                 apply_filter=True):
        from .filters import design_bandpass_sos

        torch.manual_seed(seed)
        self.model = Model(input_size, lstm_size, lstm_layers, output_size, include_top)
        self.loss = DINOLossSingleView(output_size, 1, warmup_teacher_temp, teacher_temp,
                                       warmup_teacher_temp_epochs, nepochs)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr)
        self.sos = design_bandpass_sos(low_hz, high_hz, fs, order)
        self.zero_phase = zero_phase
        self.apply_filter = apply_filter

    def filter(self, eeg_bct: np.ndarray) -> np.ndarray:
        from scipy.signal import sosfilt, sosfiltfilt

        if not self.apply_filter:
            return eeg_bct
        f = sosfiltfilt if self.zero_phase else sosfilt
        return f(self.sos, eeg_bct.astype(np.float64), axis=-1).astype(np.float32)

    def step(self, eeg_bct: np.ndarray, teacher_feats: np.ndarray, epoch: int = 0):
        x = torch.from_numpy(np.ascontiguousarray(self.filter(eeg_bct))).transpose(1, 2).contiguous()  # [B, T, C]
        t = torch.from_numpy(teacher_feats)
        self.opt.zero_grad()
        out = self.model(x)
        if isinstance(out, tuple):
            out = out[0]
        loss = self.loss(out, t, epoch)
        loss.backward()
        self.opt.step()
        return float(loss)


# ---- the losses LstmDistillFromDinoV2Train.py runs today (pinned by tests/golden/alt_losses.npz, which the reference's
# ---- own classes generated) ------------------------------------------------------------------------------------------
def feature_distribution_loss(student, teacher, temperature, label, pred_label, alpha=0.5, beta=0.5):
    """LstmDistillFromDinoV2Train.py:118-140: alpha * CE(pred_label, label) + beta * F.cross_entropy(softmax(teacher/T),
    softmax(student/T)) -- teacher probabilities as the logits argument, student probabilities as the soft target."""
    import torch.nn.functional as F
    teacher_p = F.softmax(teacher / temperature, dim=-1)       # :124
    student_p = F.softmax(student / temperature, dim=-1)       # :125
    term1 = alpha * F.cross_entropy(pred_label, label)         # :128
    term2 = beta * F.cross_entropy(teacher_p, student_p)       # :129
    return term1 + term2


def cosine_similarity_loss(student, teacher):
    """LstmDistillFromDinoV2Train.py:36-43."""
    return 1 - nn.CosineSimilarity()(student, teacher).mean()


def kd_loss_retrieval(student_logits, teacher_logits, temperature, soft_target_loss_weight, ce_loss_weight):
    """LSTMDistillRetreival.py:40-70 (`loss_fn_kd`): T^2-scaled soft-target divergence summed over the batch and divided
    by B, plus the element-mean smooth-L1 between the raw logits."""
    T = temperature
    q = F.softmax(teacher_logits / T, dim=-1)
    logp = F.log_softmax(student_logits / T, dim=-1)
    soft = torch.sum(q * (q.log() - logp)) / logp.size(0) * (T ** 2)
    return soft_target_loss_weight * soft + ce_loss_weight * F.smooth_l1_loss(student_logits, teacher_logits)


def kd_loss_hinton(outputs, labels, teacher_outputs, temperature, alpha):
    """LstmDistillFromDinoV2TrainSpampinato.py:107-121 (`loss_fn_kd`): nn.KLDivLoss() (element mean) * alpha T^2 +
    cross-entropy on the hard labels * (1 - alpha)."""
    T = temperature
    kd = nn.KLDivLoss()(F.log_softmax(outputs / T, dim=1), F.softmax(teacher_outputs / T, dim=1)) * (alpha * T * T)
    return kd + F.cross_entropy(outputs, labels) * (1.0 - alpha)


def neg_cosine_loss(v1, v2):
    """LSTMDistill.py:33-58 (`cosine_similarity_loss`, returned by its loss_fn_kd :60-98)."""
    return -torch.mean(torch.sum(F.normalize(v1, p=2, dim=1) * F.normalize(v2, p=2, dim=1), dim=1))
