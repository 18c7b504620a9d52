"""Band-pass kernel vs the float64 oracle (scipy arithmetic) -- through the C-ABI.
Tolerance: the kernel keeps fp32 state; scipy is float64.  On N(0,1) input the fp32-coefficient cascade differs
from float64 by ~3e-5, so atol = 2e-4 for fp32 output and 2e-2 for bf16 output."""
import numpy as np
import pytest
import torch

from oracle.filters import design_bandpass_sos, sosfilt_np, sosfiltfilt_np, synthetic_eeg

pytestmark = pytest.mark.gpu


def _ops():
    from cerebralsignalnetworks_b200 import ops
    return ops


@pytest.mark.parametrize("B,C,T", [(3, 5, 440), (2, 128, 440), (1, 1, 1), (2, 3, 33), (4, 63, 200), (1, 96, 495), (7, 130, 64)])
@pytest.mark.parametrize("order", [3, 4, 5])
def test_sosfilt_bct(B, C, T, order):
    ops = _ops()
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, order)
    x = synthetic_eeg(B, C, T, seed=B + C + T)
    ref = sosfilt_np(sos, x.astype(np.float64))
    y = ops.sosfilt(torch.from_numpy(x).cuda(), sos).cpu().numpy()
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-4)


@pytest.mark.parametrize("layout", ["BTC", "TBC"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sosfilt_fused_layout_and_cast(layout, dtype):
    ops = _ops()
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    x = synthetic_eeg(5, 96, 300, seed=7)
    ref = sosfilt_np(sos, x.astype(np.float64))
    ref = ref.transpose(0, 2, 1) if layout == "BTC" else ref.transpose(2, 0, 1)
    y = ops.sosfilt(torch.from_numpy(x).cuda(), sos, out_layout=layout, out_dtype=dtype).float().cpu().numpy()
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("B,C,T,layout", [(3, 5, 440, "BCT"), (2, 40, 100, "BTC"), (2, 33, 28, "TBC"), (1, 63, 2000, "BCT")])
def test_zero_phase(B, C, T, layout):
    ops = _ops()
    sos = design_bandpass_sos(1.0, 50.0, 1000.0, 4)
    x = synthetic_eeg(B, C, T, seed=11)
    ref = sosfiltfilt_np(sos, x.astype(np.float64))
    ref = {"BCT": ref, "BTC": ref.transpose(0, 2, 1), "TBC": ref.transpose(2, 0, 1)}[layout]
    y = ops.sosfilt(torch.from_numpy(x).cuda(), sos, zero_phase=True, out_layout=layout).cpu().numpy()
    # low cut-off at 1 Hz: poles close to z=1, fp32 state costs a little more
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-3)


def test_golden_and_remove_noise(golden):
    import cerebralsignalnetworks_b200 as csn
    g = golden("filters.npz")
    x = torch.from_numpy(g["x"]).cuda()
    f = csn.EEGFilters(1000.0)
    np.testing.assert_allclose(f.apply(x, 5.0, 95.0, 4).cpu().numpy(), g["sosfilt_5_95"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(f.apply(x, 5.0, 95.0, 4, zero_phase=True).cpu().numpy(), g["sosfiltfilt_5_95"], rtol=0, atol=5e-4)
    # Utilities.remove_noise drop-in: [S, T, C] in, [S, T, C] out, reference's own output as the target
    out = f.remove_noise(g["x"].transpose(0, 2, 1)).cpu().numpy()
    np.testing.assert_allclose(out, g["remove_noise_1_50"], rtol=0, atol=2e-3)


def test_errors_and_empty():
    from cerebralsignalnetworks_b200 import _lib
    ops = _ops()
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    assert ops.sosfilt(torch.zeros(0, 4, 16, device="cuda"), sos).shape == (0, 4, 16)
    with pytest.raises(_lib.CsnError):
        ops.sosfilt(torch.zeros(1, 1, 20, device="cuda"), sos, zero_phase=True)  # T <= padlen, as scipy raises
    with pytest.raises(_lib.CsnError):
        ops.sosfilt(torch.zeros(1, 1, 20), sos)  # CPU tensor: no fallback


def test_full_size_properties():
    """cfg2 size [256,128,440]: linearity + time invariance + agreement with the oracle on a row sample."""
    ops = _ops()
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(256, 128, 440, device="cuda", generator=g)
    b = torch.randn(256, 128, 440, device="cuda", generator=g)
    ya, yb = ops.sosfilt(a, sos), ops.sosfilt(b, sos)
    yc = ops.sosfilt((2 * a - 3 * b).contiguous(), sos)
    assert torch.allclose(yc, 2 * ya - 3 * yb, atol=2e-4)
    shifted = torch.zeros_like(a); shifted[..., 5:] = a[..., :-5]
    ys = ops.sosfilt(shifted, sos)
    assert torch.allclose(ys[..., 5:], ya[..., :-5], atol=1e-5)
    rows = [(0, 0), (17, 3), (255, 127), (100, 64)]
    for bb, cc in rows:
        ref = sosfilt_np(sos, a[bb, cc].double().cpu().numpy())
        np.testing.assert_allclose(ya[bb, cc].cpu().numpy(), ref, atol=2e-4)
    y_tbc = ops.sosfilt(a, sos, out_layout="TBC", out_dtype=torch.bfloat16)
    assert y_tbc.shape == (440, 256, 128)
    assert torch.allclose(y_tbc.float(), ya.permute(2, 0, 1), atol=3e-2)


@pytest.mark.parametrize("B,C,T", [(2, 16, 36), (4, 16, 100), (1, 32, 440), (3, 5, 37), (2, 16, 4)])
@pytest.mark.parametrize("layout,dtype", [("TBC", torch.bfloat16), ("TBC", torch.float32), ("BCT", torch.float32), ("BTC", torch.float32)])
def test_filter_writes_stay_inside_the_output(B, C, T, layout, dtype):
    """Guard bands around the input and the output (the pool has no compute-sanitizer): the streaming kernels read in
    16-byte pieces and write whole tiles, so a tail bug would show up as a clobbered guard or a NaN pulled in."""
    from cerebralsignalnetworks_b200 import ops
    from oracle.filters import design_bandpass_sos, sosfilt_np
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    n = B * C * T
    pad = 256
    xin = torch.full((n + 2 * pad,), float("nan"), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    x = xin[pad:pad + n].view(B, C, T)
    x.copy_(torch.randn(B, C, T, device="cuda", generator=g))
    yout = torch.full((n + 2 * pad,), 12288.0, device="cuda", dtype=dtype)
    shape = {"BCT": (B, C, T), "BTC": (B, T, C), "TBC": (T, B, C)}[layout]
    y = ops.sosfilt(x, sos, out_layout=layout, out_dtype=dtype, out=yout[pad:pad + n].view(shape))
    torch.cuda.synchronize()
    assert (yout[:pad].float() == 12288.0).all() and (yout[pad + n:].float() == 12288.0).all()
    assert torch.isfinite(y.float()).all()
    ref = sosfilt_np(sos, x.cpu().numpy())
    perm = {"BCT": (0, 1, 2), "BTC": (0, 2, 1), "TBC": (2, 0, 1)}[layout]
    tol = 2e-4 if dtype == torch.float32 else 3e-2
    np.testing.assert_allclose(y.float().cpu().numpy(), ref.transpose(perm), atol=tol * max(1.0, np.abs(ref).max()))
