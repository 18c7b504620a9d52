"""GPU-resident EEG dataset.

Takes the .pth dictionary written by ConvertToPth.py:170-201 ({"dataset": [{"eeg": Tensor[C, T_raw], "image": i,
"label": k, "subject": s}, ...], "labels": [...], "images": [...], ...}) -- the file utils/PerilsEEGDataset.EEGDataset
loads (:62) -- uploads the EEG tensors ONCE as a [N, C, T_raw] fp32 tensor and serves batches with one gather kernel
(csn_gather_trials) instead of a per-item Python `__getitem__` + DataLoader collate + host-to-device copy.  Same item
semantics as EEGDataset.__getitem__ (:541-573): crop [time_low, time_high), optional (x - mean) / std with the
dataset-level scalars of :93-103 (mean of the per-item means, mean of the per-item unbiased stds)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import LAYOUT_BCT, LAYOUT_BTC
from .ops import _p, _stream, call


class DeviceEEGDataset:
    def __init__(self, loaded, time_low=20, time_high=480, apply_norm_with_stds_and_means=False, device="cuda",
                 filter_channels=(), apply_channel_wise_norm=False):
        _lib.require_gpu()
        if isinstance(loaded, str):
            loaded = torch.load(loaded, map_location="cpu", weights_only=False)
        items = loaded["dataset"]
        if len(items) == 0:
            raise _lib.CsnError("empty dataset")
        self.class_labels = list(loaded.get("labels", []))
        self.image_names = list(loaded.get("images", []))
        self.time_low, self.time_high = int(time_low), int(time_high)
        self.apply_norm_with_stds_and_means = bool(apply_norm_with_stds_and_means)
        # dataset-level scalars exactly as utils/PerilsEEGDataset.py:93-103 accumulates them (stored dtype, torch's
        # unbiased std, arithmetic mean over items)
        mean = sum(it["eeg"].mean() for it in items) / len(items)
        std = sum(it["eeg"].std() for it in items) / len(items)
        self.mean, self.std = float(mean), float(std)
        shapes = {tuple(it["eeg"].shape) for it in items}
        if len(shapes) != 1:
            raise _lib.CsnError("trials of different shape cannot share one resident tensor: %s" % sorted(shapes))
        self.device = torch.device(device)
        self.eeg = torch.stack([it["eeg"].float() for it in items]).contiguous().to(self.device)  # [N, C, T_raw]
        self.labels = torch.tensor([int(it["label"]) for it in items], dtype=torch.int64, device=self.device)
        self.image_index = torch.tensor([int(it["image"]) for it in items], dtype=torch.int64, device=self.device)
        self.N, self.C, self.T_raw = self.eeg.shape
        if not (0 <= self.time_low < self.time_high <= self.T_raw):
            raise _lib.CsnError("need 0 <= time_low < time_high <= %d" % self.T_raw)
        if len(filter_channels) > 0:
            self.select_channels(filter_channels, apply_channel_wise_norm)

    def select_channels(self, filter_channels, apply_channel_wise_norm=False):
        """The `filter_channels` branch of EEGDataset.__getitem__ (utils/PerilsEEGDataset.py:554-565), applied once to the
        resident tensor: keep the listed channels (in the listed order), crop [time_low, time_high), and with
        `apply_channel_wise_norm` z-score every (trial, channel) row over the cropped window (normlizeEEG :454-461,
        population standard deviation).  Afterwards the dataset serves [B, len(filter_channels), T] batches; the
        dataset-level mean / std stay those of the original trials, as in the reference."""
        ch = [int(c) for c in filter_channels]
        if not ch or min(ch) < 0 or max(ch) >= self.C:
            raise IndexError("filter_channels must be a non-empty list of channel indices in [0, %d)" % self.C)
        ch_dev = torch.tensor(ch, dtype=torch.int32, device=self.device)
        out = torch.empty((self.N, len(ch), self.samples), dtype=torch.float32, device=self.device)
        call("csn_select_crop_zscore", _p(self.eeg), _p(ch_dev), _p(out), self.N, self.C, self.T_raw, len(ch), self.time_low,
             self.time_high, 1 if apply_channel_wise_norm else 0, _stream())
        self.eeg, self.C, self.T_raw = out, len(ch), self.samples
        self.time_low, self.time_high = 0, self.T_raw
        self.filter_channels = ch
        return self

    @classmethod
    def from_tensor(cls, eeg_nct, labels=None, image_index=None, time_low=20, time_high=480,
                    apply_norm_with_stds_and_means=False, device="cuda"):
        """Same dataset from an already stacked [N, C, T_raw] array (any device) instead of the .pth item list."""
        _lib.require_gpu()
        self = cls.__new__(cls)
        self.class_labels, self.image_names = [], []
        self.time_low, self.time_high = int(time_low), int(time_high)
        self.apply_norm_with_stds_and_means = bool(apply_norm_with_stds_and_means)
        self.device = torch.device(device)
        self.eeg = torch.as_tensor(eeg_nct).to(device=self.device, dtype=torch.float32).contiguous()
        if self.eeg.dim() != 3 or self.eeg.shape[0] == 0:
            raise _lib.CsnError("from_tensor: need a non-empty [N, C, T_raw] array")
        self.N, self.C, self.T_raw = self.eeg.shape
        self.mean = float(self.eeg.mean(dim=(1, 2)).mean())   # mean of the per-item means / unbiased stds (:93-103)
        self.std = float(self.eeg.std(dim=(1, 2)).mean())
        n = self.N
        self.labels = (torch.zeros(n, dtype=torch.int64) if labels is None else torch.as_tensor(labels, dtype=torch.int64)).to(self.device)
        self.image_index = (torch.arange(n) if image_index is None else torch.as_tensor(image_index, dtype=torch.int64)).to(self.device)
        if not (0 <= self.time_low < self.time_high <= self.T_raw):
            raise _lib.CsnError("need 0 <= time_low < time_high <= %d" % self.T_raw)
        return self

    def __len__(self):
        return self.N

    @property
    def samples(self):
        return self.time_high - self.time_low

    def _gather(self, indices, layout):
        idx = torch.as_tensor(indices, dtype=torch.int64, device=self.device).contiguous()
        if idx.dim() != 1:
            raise _lib.CsnError("indices must be one-dimensional")
        if idx.numel() and (int(idx.max()) >= self.N or int(idx.min()) < -self.N):
            raise IndexError("trial index out of range for %d trials" % self.N)
        B, T = idx.numel(), self.samples
        shape = (B, self.C, T) if layout == LAYOUT_BCT else (B, T, self.C)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        mean, std = (self.mean, self.std) if self.apply_norm_with_stds_and_means else (0.0, 1.0)
        call("csn_gather_trials", _p(self.eeg), _p(idx), _p(out), self.N, self.C, self.T_raw, B, self.time_low,
             self.time_high, float(mean), float(std), layout, _stream())
        return out, idx

    def check_indices(self, indices):
        """Host-side validation of a batch of trial indices (no device synchronisation) -> contiguous CPU int64 [B]."""
        idx = torch.as_tensor(indices, dtype=torch.int64, device="cpu").contiguous()
        if idx.dim() != 1:
            raise _lib.CsnError("indices must be one-dimensional")
        if idx.numel() and (int(idx.max()) >= self.N or int(idx.min()) < -self.N):
            raise IndexError("trial index out of range for %d trials" % self.N)
        return idx

    def gather_into(self, out_bct, idx_dev):
        """The gather kernel alone: trials `idx_dev` (device int64 [B], already validated) -> `out_bct` ([B, C, T]
        float32, preallocated).  No allocation, no synchronisation: safe on any stream and under graph capture."""
        B = idx_dev.numel()
        if tuple(out_bct.shape) != (B, self.C, self.samples) or out_bct.dtype != torch.float32 or not out_bct.is_contiguous():
            raise _lib.CsnError("gather_into: out must be contiguous float32 [%d, %d, %d]" % (B, self.C, self.samples))
        mean, std = (self.mean, self.std) if self.apply_norm_with_stds_and_means else (0.0, 1.0)
        call("csn_gather_trials", _p(self.eeg), _p(idx_dev), _p(out_bct), self.N, self.C, self.T_raw, B, self.time_low,
             self.time_high, float(mean), float(std), LAYOUT_BCT, _stream())
        return out_bct

    def fused_filter_ok(self):
        """Whether csn_sosfilt_gather_f32 serves this dataset's shape (else gather_into + sosfilt)."""
        return self.C % 32 == 0 and self.T_raw % 4 == 0 and self.time_low % 4 == 0 and self.samples % 4 == 0

    def norm_scalars(self):
        return (self.mean, self.std) if self.apply_norm_with_stds_and_means else (0.0, 1.0)

    def batch(self, indices):
        """-> (eeg [B, C, T] float32 in the stored layout -- feed it to DistillTrainStep.step / Model.encode_trials --,
        labels [B] int64, image indices [B] int64), all on the GPU."""
        out, idx = self._gather(indices, LAYOUT_BCT)
        return out, self.labels[idx], self.image_index[idx]

    def batch_btc(self, indices):
        """Same batch in the DataLoader layout [B, T, C] (`eeg.t()[time_low:time_high]` per item), for Model.forward."""
        out, idx = self._gather(indices, LAYOUT_BTC)
        return out, self.labels[idx], self.image_index[idx]

    def epoch_batches(self, batch_size, shuffle=True, generator=None, drop_last=True, rank=0, world=1):
        """Index batches of one epoch; with world > 1 every rank takes its contiguous share of each global batch
        (DistributedSampler + drop_last of LstmDistillation.py:406-414)."""
        order = torch.randperm(self.N, generator=generator, device="cpu") if shuffle else torch.arange(self.N)
        per = batch_size * world
        n_full = self.N // per
        for k in range(n_full):
            g = order[k * per:(k + 1) * per]
            yield g[rank * batch_size:(rank + 1) * batch_size]
        if not drop_last and world == 1 and self.N % per:
            yield order[n_full * per:]
