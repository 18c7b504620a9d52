"""Oracle: IIR band-pass along time.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference anchors
  * design:  utils/EEGFilters.py:26-39  (scipy.signal.butter(order, [lo, hi], btype='bandpass'))
  * apply:   utils/Utilities.py:411-428 (butter(4, [1, 50]/nyq, 'band') + filtfilt per (sample, channel))
The arithmetic lives in SciPy (un-vendored, no version pin in the reference;
scipy 1.18.1 in this image).  The apply path is restated here in float64 numpy
(direct-form-II-transposed biquad cascade, which is what scipy's `_sosfilt`
does) and pinned against scipy.signal.{sosfilt,sosfiltfilt,filtfilt} in
tests/test_oracle_filters.py.
"""
from __future__ import annotations

import numpy as np


def design_bandpass_sos(low_hz: float, high_hz: float, fs: float, order: int = 4) -> np.ndarray:
    """Butterworth band-pass as second-order sections, float64 [order, 6].

    Same normalisation as utils/EEGFilters.py:14-15 (cut-off / (fs/2)).  SOS form is
    mandatory: the (b, a) designs of EEGFilters are numerically unstable for order >= 4
    (SURVEY.md section 0 fact 3)."""
    from scipy.signal import butter

    nyq = 0.5 * fs
    return np.asarray(butter(order, [low_hz / nyq, high_hz / nyq], btype="bandpass", output="sos"), dtype=np.float64)


def sosfilt_np(sos: np.ndarray, x: np.ndarray, zi: np.ndarray | None = None):
    """Cascade of DF2T biquads along the last axis, float64.

    y[n]  = b0 x[n] + s1
    s1'   = b1 x[n] - a1 y[n] + s2
    s2'   = b2 x[n] - a2 y[n]
    zi: optional [n_sections, ..., 2] initial state; returns (y, zf) if given."""
    sos = np.asarray(sos, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    lead = x.shape[:-1]
    n_sec = sos.shape[0]
    zf = np.zeros((n_sec,) + lead + (2,), dtype=np.float64)
    for s in range(n_sec):
        b0, b1, b2, a0, a1, a2 = sos[s]
        b0, b1, b2, a1, a2 = b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0
        if zi is not None:
            s1 = np.array(zi[s][..., 0], dtype=np.float64, copy=True) * np.ones(lead)
            s2 = np.array(zi[s][..., 1], dtype=np.float64, copy=True) * np.ones(lead)
        else:
            s1 = np.zeros(lead)
            s2 = np.zeros(lead)
        for n in range(y.shape[-1]):
            xn = y[..., n].copy()
            yn = b0 * xn + s1
            s1 = b1 * xn - a1 * yn + s2
            s2 = b2 * xn - a2 * yn
            y[..., n] = yn
        zf[s][..., 0] = s1
        zf[s][..., 1] = s2
    if zi is not None:
        return y, zf
    return y


def sosfilt_zi_np(sos: np.ndarray) -> np.ndarray:
    """Steady-state step-response initial state per section (scipy.signal.sosfilt_zi)."""
    sos = np.asarray(sos, dtype=np.float64)
    n_sec = sos.shape[0]
    zi = np.zeros((n_sec, 2))
    scale = 1.0
    for s in range(n_sec):
        b = sos[s, :3] / sos[s, 3]
        a = sos[s, 3:] / sos[s, 3]
        # lfilter_zi for a biquad: solve (I - A^T) z = B with companion-form A
        # z0 = (b1 - a1 b0) + z1 - a1 * ... ; closed form below
        # state equations at steady state with unit input and output y = sum(b)/sum(a):
        yss = b.sum() / a.sum()
        z2 = b[2] - a[2] * yss
        z1 = b[1] - a[1] * yss + z2
        zi[s] = scale * np.array([z1, z2])
        scale *= yss
    return zi


def sosfiltfilt_np(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Zero-phase forward-backward filtering with odd extension (scipy.signal.sosfiltfilt /
    filtfilt default padtype='odd'), float64.  padlen = 3*ntaps with scipy's ntaps rule;
    equals filtfilt(b, a)'s 3*max(len(a), len(b)) = 27 for the order-4 band-pass used at
    utils/Utilities.py:421-427."""
    sos = np.asarray(sos, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n_sec = sos.shape[0]
    ntaps = 2 * n_sec + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    padlen = 3 * ntaps
    if x.shape[-1] <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    left = 2 * x[..., :1] - x[..., padlen:0:-1]
    right = 2 * x[..., -1:] - x[..., -2:-(padlen + 2):-1]
    ext = np.concatenate([left, x, right], axis=-1)
    zi = sosfilt_zi_np(sos)  # [n_sec, 2]
    lead = x.shape[:-1]
    zi_b = zi.reshape((n_sec,) + (1,) * len(lead) + (2,))
    x0 = ext[..., :1]
    y, _ = sosfilt_np(sos, ext, zi=zi_b * x0[None, ...])
    y0 = y[..., -1:]
    y, _ = sosfilt_np(sos, y[..., ::-1], zi=zi_b * y0[None, ...])
    y = y[..., ::-1]
    return y[..., padlen:-padlen]


def remove_noise_ref(eeg_data: np.ndarray, sampling_rate: float) -> np.ndarray:
    """The reference's only *applied* band-pass, verbatim semantics of
    utils/Utilities.py:411-428: order-4 Butterworth 1-50 Hz in (b, a) form, scipy filtfilt
    over time for each (sample, channel) of eeg_data[S, T, C]."""
    from scipy import signal

    nyquist_freq = 0.5 * sampling_rate
    b, a = signal.butter(4, [1.0 / nyquist_freq, 50.0 / nyquist_freq], btype="band")
    out = np.zeros_like(eeg_data)
    for s in range(eeg_data.shape[0]):
        for i in range(eeg_data.shape[-1]):
            out[s, :, i] = signal.filtfilt(b, a, eeg_data[s, :, i])
    return out


def synthetic_eeg(batch: int, channels: int, samples: int, seed: int = 43, fs: float = 1000.0,
                  frequency: float = 40.0, amplitude: float = 0.5) -> np.ndarray:
    """Synthetic trials [B, C, T] fp32 with the reference's own generator formula
    N(0,1) + 0.5 sin(2 pi 40 t / fs)  (utils/PerilsEEGDataset.py:140-147)."""
    rng = np.random.default_rng(seed)
    noise = rng.normal(0.0, 1.0, size=(batch, channels, samples))
    t = np.arange(samples) / fs
    return (noise + amplitude * np.sin(2 * np.pi * frequency * t)).astype(np.float32)
