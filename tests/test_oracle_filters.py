"""Oracle filter restatement vs scipy (the arithmetic the reference calls) and vs the golden vectors."""
import numpy as np
import pytest
from scipy import signal

from oracle.filters import (design_bandpass_sos, remove_noise_ref, sosfilt_np, sosfilt_zi_np, sosfiltfilt_np,
                            synthetic_eeg)


@pytest.mark.parametrize("order,lo,hi", [(4, 5.0, 95.0), (3, 0.1, 60.0), (5, 1.0, 50.0), (4, 14.0, 71.0)])
def test_sosfilt_matches_scipy(order, lo, hi):
    sos = design_bandpass_sos(lo, hi, 1000.0, order)
    x = np.random.default_rng(0).normal(size=(2, 3, 200))
    np.testing.assert_allclose(sosfilt_np(sos, x), signal.sosfilt(sos, x, axis=-1), rtol=1e-10, atol=1e-12)


def test_zi_matches_scipy():
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    np.testing.assert_allclose(sosfilt_zi_np(sos), signal.sosfilt_zi(sos), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("T", [28, 64, 440])
def test_sosfiltfilt_matches_scipy(T):
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    x = np.random.default_rng(1).normal(size=(2, 2, T))
    np.testing.assert_allclose(sosfiltfilt_np(sos, x), signal.sosfiltfilt(sos, x, axis=-1), rtol=1e-8, atol=1e-10)


def test_sosfiltfilt_too_short_raises():
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    with pytest.raises(ValueError):
        sosfiltfilt_np(sos, np.zeros((1, 27)))


def test_golden(golden):
    g = golden("filters.npz")
    x = g["x"].astype(np.float64)
    np.testing.assert_allclose(sosfilt_np(g["sos_5_95"], x), g["sosfilt_5_95"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sosfiltfilt_np(g["sos_5_95"], x), g["sosfiltfilt_5_95"], rtol=1e-7, atol=1e-9)
    # the reference's own remove_noise ((b, a) filtfilt, [S, T, C]) vs the SOS zero-phase oracle
    ref = g["remove_noise_1_50"]  # [S, T, C]
    ours = sosfiltfilt_np(g["sos_1_50"], x).transpose(0, 2, 1)
    # (b, a) form at 1-50 Hz / 1 kHz carries ~1.5e-4 of its own round-off (poles near z=1); SOS is the accurate one
    np.testing.assert_allclose(ours, ref, rtol=0, atol=5e-4)
    np.testing.assert_allclose(remove_noise_ref(x.transpose(0, 2, 1), 1000.0), ref, rtol=1e-9, atol=1e-11)


def test_linearity_and_shape():
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    a = synthetic_eeg(2, 4, 100, seed=1).astype(np.float64)
    b = synthetic_eeg(2, 4, 100, seed=2).astype(np.float64)
    np.testing.assert_allclose(sosfilt_np(sos, 2 * a - 3 * b), 2 * sosfilt_np(sos, a) - 3 * sosfilt_np(sos, b),
                               rtol=1e-9, atol=1e-10)
    assert synthetic_eeg(3, 5, 7).shape == (3, 5, 7) and synthetic_eeg(3, 5, 7).dtype == np.float32
