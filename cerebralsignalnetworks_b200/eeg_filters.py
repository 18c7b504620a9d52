"""`EEGFilters` with the reference constructor (utils/EEGFilters.py:4-44) plus the apply path the reference
never had: the reference class only DESIGNS Butterworth band-passes (orders 3/4/5, (b, a) form, local
variables); the only applied band-pass in its tree is Utilities.remove_noise (utils/Utilities.py:411-428,
filtfilt per (sample, channel) in a Python double loop).  Here the design stays on the host (scipy, SOS form --
the (b, a) designs at 0.1-60 Hz / 1 kHz are unstable for order >= 4) and the apply runs in the
libcsn_b200 band-pass kernel over the whole [B, C, T] batch."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def butter_bandpass_sos(low_hz, high_hz, fs, order=4):
    from scipy.signal import butter

    nyq = 0.5 * fs
    return np.asarray(butter(order, [low_hz / nyq, high_hz / nyq], btype="bandpass", output="sos"), dtype=np.float64)


class EEGFilters:
    def __init__(self, fs) -> None:
        # same attributes as the reference (utils/EEGFilters.py:9-16)
        self.low_cutoff = 0.1
        self.high_cutoff = 60.0
        self.fs = fs
        self.low_cutoff_norm = self.low_cutoff / (self.fs / 2)
        self.high_cutoff_norm = self.high_cutoff / (self.fs / 2)
        self._sos_cache = {}

    def sos(self, low=None, high=None, order=4):
        low = self.low_cutoff if low is None else low
        high = self.high_cutoff if high is None else high
        key = (float(low), float(high), int(order))
        if key not in self._sos_cache:
            self._sos_cache[key] = butter_bandpass_sos(low, high, self.fs, order)
        return self._sos_cache[key]

    def apply(self, x, low=None, high=None, order=4, zero_phase=False, out_layout="BCT", out_dtype=torch.float32):
        """Band-pass x [B, C, T] (float32, CUDA) along time.  zero_phase=True reproduces filtfilt
        (Utilities.remove_noise semantics); out_layout='TBC' + out_dtype=bfloat16 is the fused hand-off to the
        LSTM encoder."""
        return ops.sosfilt(x, self.sos(low, high, order), zero_phase=zero_phase, out_layout=out_layout,
                           out_dtype=out_dtype)

    def remove_noise(self, eeg_data, sampling_rate=None):
        """Utilities.remove_noise drop-in (utils/Utilities.py:411-428): eeg_data [S, T, C] -> zero-phase order-4
        Butterworth 1-50 Hz, returned as [S, T, C]."""
        fs = self.fs if sampling_rate is None else sampling_rate
        sos = butter_bandpass_sos(1.0, 50.0, fs, 4)
        x = torch.as_tensor(eeg_data, dtype=torch.float32, device="cuda").permute(0, 2, 1).contiguous()
        return ops.sosfilt(x, sos, zero_phase=True, out_layout="BTC")
