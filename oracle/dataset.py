"""ORACLE (test infrastructure only).  CPU restatement of what utils/PerilsEEGDataset.EEGDataset does to the EEG
tensors of the .pth file (ConvertToPth.py:170-201): dataset-level mean / std (:93-103) and the per-item transform of
`__getitem__` on the default path (:541-573: float, transpose to [T_raw, C], crop [time_low, time_high), optional
(x - mean) / std).  Pinned by tests/golden/dataset.npz, which oracle/make_golden.py produced by running the reference's
own EEGDataset class on a small synthetic .pth file."""
from __future__ import annotations

import torch


def dataset_scalars(items):
    """:93-103 -- mean of the per-item means and mean of the per-item (unbiased) standard deviations."""
    mean = std = 0
    for it in items:
        mean += it["eeg"].mean()
        std += it["eeg"].std()
    return float(mean / len(items)), float(std / len(items))


def item_transform(eeg_ct, time_low, time_high, mean=None, std=None):
    """:541-573 on the default path -> [T, C] float32."""
    eeg = eeg_ct.float().t()                 # :542, :551
    eeg = eeg[time_low:time_high, :]          # :569
    if mean is not None:
        eeg = (eeg - mean) / std              # :572-573
    return eeg


def item_transform_channels(eeg_ct, time_low, time_high, filter_channels, channel_wise_norm=False, mean=None, std=None):
    """:554-565 (`filter_channels` branch) -> [len(filter_channels), T] float32 (the branch transposes back): selected
    channels of the cropped window, each optionally z-scored with numpy's population std (normlizeEEG :454-461)."""
    import numpy as np
    eeg = eeg_ct.float().t().cpu().numpy()
    out = np.zeros((time_high - time_low, len(filter_channels)), dtype=eeg.dtype)
    for k, ch in enumerate(filter_channels):
        out[:, k] = eeg[time_low:time_high, ch]
        if channel_wise_norm:
            col = out[:, k]
            out[:, k] = (col - col.mean()) / col.std()
    res = torch.from_numpy(out).t()
    if mean is not None:
        res = (res - mean) / std
    return res
