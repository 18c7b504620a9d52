"""LSTM encoder layer through the C-ABI vs torch.nn.LSTM on the CPU (the arithmetic the restated
models.lstm.Model stands on).
  fp32 mode : rtol 1e-4 / atol 1e-5 on h_seq, 1e-3 relative on gradients.
  bf16 mode : bf16 operands, fp32 accumulation and cell state; tolerance calibrated against the bf16 round-off
              floor: |h - h_ref| <= 3e-2 abs, gradient cosine >= 0.995 and relative L2 error <= 6e-2."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_layer(x_tbi, w, d_hseq=None, d_hlast=None):
    T, B, I = x_tbi.shape
    H = w[1].shape[1]
    lstm = torch.nn.LSTM(I, H, 1)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(w[0]); lstm.weight_hh_l0.copy_(w[1]); lstm.bias_ih_l0.copy_(w[2]); lstm.bias_hh_l0.copy_(w[3])
    x = x_tbi.clone().requires_grad_(True)
    out, _ = lstm(x)
    obj = 0
    if d_hseq is not None:
        obj = obj + (out * d_hseq).sum()
    if d_hlast is not None:
        obj = obj + (out[-1] * d_hlast).sum()
    obj.backward()
    return out.detach(), [lstm.weight_ih_l0.grad, lstm.weight_hh_l0.grad, lstm.bias_ih_l0.grad, lstm.bias_hh_l0.grad], x.grad


def _weights(I, H, seed):
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / np.sqrt(H)
    return [(torch.rand(4 * H, I, generator=g) * 2 - 1) * k, (torch.rand(4 * H, H, generator=g) * 2 - 1) * k,
            (torch.rand(4 * H, generator=g) * 2 - 1) * k, (torch.rand(4 * H, generator=g) * 2 - 1) * k]


def _run(x, w, dtype, d_hseq, d_hlast, need_dx):
    from cerebralsignalnetworks_b200 import ops
    wc = [t.cuda() for t in w]
    xc = x.cuda().to(dtype).contiguous()
    h_seq, reserve, ws = ops.lstm_layer_fwd(xc, *wc, dtype, True)
    grads = tuple(torch.full_like(t, float("nan")) for t in wc)
    dx = ops.lstm_layer_bwd(xc, wc[0], wc[1], h_seq, reserve, ws,
                            None if d_hseq is None else d_hseq.cuda().contiguous(),
                            None if d_hlast is None else d_hlast.cuda().contiguous(), grads, need_dx, dtype)
    torch.cuda.synchronize()
    return h_seq.float().cpu(), [g.cpu() for g in grads], None if dx is None else dx.cpu()


@pytest.mark.parametrize("T,B,I,H", [(5, 3, 8, 16), (40, 9, 24, 32), (12, 2, 5, 7), (20, 17, 128, 128), (7, 4, 16, 160)])
@pytest.mark.parametrize("mode", ["last", "seq", "both"])
def test_f32_layer(T, B, I, H, mode):
    g = torch.Generator().manual_seed(T * B + I)
    x = torch.randn(T, B, I, generator=g)
    w = _weights(I, H, 1)
    d_hseq = torch.randn(T, B, H, generator=g) if mode in ("seq", "both") else None
    d_hlast = torch.randn(B, H, generator=g) if mode in ("last", "both") else None
    ref_h, ref_g, ref_dx = _ref_layer(x, w, d_hseq, d_hlast)
    h, gr, dx = _run(x, w, torch.float32, d_hseq, d_hlast, True)
    np.testing.assert_allclose(h.numpy(), ref_h.numpy(), rtol=1e-4, atol=1e-5)
    for a, b, n in zip(gr, ref_g, ["dw_ih", "dw_hh", "db_ih", "db_hh"]):
        scale = b.abs().max().item() + 1e-8
        assert (a - b).abs().max().item() <= 1e-3 * scale, n
    assert (dx - ref_dx).abs().max().item() <= 1e-3 * (ref_dx.abs().max().item() + 1e-8)


def _cos(a, b):
    return float((a.flatten() @ b.flatten()) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("T,B,I,H", [(6, 4, 16, 16), (50, 5, 32, 64), (30, 37, 128, 128), (24, 3, 96, 96), (440, 16, 128, 128),
                                     (16, 300, 64, 128),
                                     # THE benchmark shape (cfg2: 128 recurrence CTAs of 2 trials + 20 Xp-server / dW-consumer CTAs)
                                     (440, 256, 128, 128),
                                     # side-role edges: rows not a multiple of the 128-row tile / 64-row chunk, I < 64, H = 16 / 48
                                     (33, 7, 40, 48), (19, 129, 128, 16), (64, 64, 96, 112),
                                     # edges of the fused-projection kernel: T below one block of timesteps, single trial,
                                     # NV = 8 tiles (two timesteps per block), smallest sizes, and I > 128 (hoisted projection)
                                     (1, 2, 8, 8), (2, 1, 16, 32), (3, 5, 8, 128), (9, 700, 32, 32), (17, 640, 16, 64), (5, 4, 136, 64),
                                     # large-hidden path (per-step tcgen05 GEMM with the fused cell epilogue), cfg 4 sizes
                                     (12, 5, 64, 256), (20, 130, 128, 512), (9, 3, 24, 136),
                                     # cluster recurrence (H = 256 / 512, lstm_cluster.cu): a single timestep, one trial, a ragged
                                     # last group, two trial groups per cluster at H = 256 (more than 15 groups of 16 trials),
                                     # a wide input (I = 512, the second layer of cfg 4)
                                     (1, 4, 32, 512), (3, 1, 16, 256), (37, 33, 64, 512), (6, 250, 32, 256), (10, 16, 512, 512)])
@pytest.mark.parametrize("mode", ["last", "both"])
def test_bf16_tensor_core_layer(T, B, I, H, mode):
    g = torch.Generator().manual_seed(T + B + I + H)
    x = torch.randn(T, B, I, generator=g)
    w = _weights(I, H, 2)
    d_hseq = torch.randn(T, B, H, generator=g) * 0.1 if mode == "both" else None
    d_hlast = torch.randn(B, H, generator=g)
    # reference sees the same bf16-rounded input the kernel sees
    ref_h, ref_g, ref_dx = _ref_layer(x.bfloat16().float(), w, d_hseq, d_hlast)
    h, gr, dx = _run(x, w, torch.bfloat16, d_hseq, d_hlast, True)
    assert torch.isfinite(h).all()
    assert (h - ref_h).abs().max().item() <= 3e-2, (h - ref_h).abs().max().item()
    for a, b, n in zip(gr, ref_g, ["dw_ih", "dw_hh", "db_ih", "db_hh"]):
        assert torch.isfinite(a).all(), n
        if b.norm().item() == 0.0:  # T = 1: no recurrent term, dW_hh is exactly zero
            assert a.abs().max().item() == 0.0, n
            continue
        rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        assert _cos(a, b) >= 0.995 and rel <= 6e-2, (n, _cos(a, b), rel)
    rel = ((dx - ref_dx).norm() / (ref_dx.norm() + 1e-30)).item()
    assert _cos(dx, ref_dx) >= 0.995 and rel <= 6e-2, ("dx", rel)


def test_bf16_unsupported_shape_is_an_error_not_a_fallback():
    from cerebralsignalnetworks_b200 import ops, _lib
    with pytest.raises(_lib.CsnError):
        ops.lstm_layer_bytes(10, 4, 63, 128, torch.bfloat16)  # TMA needs 16-byte rows: I % 8 != 0 is rejected
    with pytest.raises(_lib.CsnError):
        ops.lstm_layer_bytes(10, 4, 64, 100, torch.bfloat16)


@pytest.mark.parametrize("T,B,I,H", [(13, 6, 64, 128), (2000, 4, 64, 128), (7, 3, 136, 96), (11, 9, 32, 256), (600, 20, 64, 512)])
def test_bf16_inference_only_forward(T, B, I, H):
    """training = 0 (no BPTT reserve): the path Model.encode_trials takes, including the 2000-sample trials of config 5."""
    from cerebralsignalnetworks_b200 import ops
    g = torch.Generator().manual_seed(T + H)
    x = torch.randn(T, B, I, generator=g) * 0.5
    w = _weights(I, H, 3)
    ref_h, _, _ = _ref_layer(x.bfloat16().float(), w, None, torch.zeros(B, H))
    h, _, _ = ops.lstm_layer_fwd(x.cuda().bfloat16().contiguous(), *[t.cuda() for t in w], torch.bfloat16, False)
    torch.cuda.synchronize()
    # bf16 round-off accumulates slowly along a long sequence (the cell state is fp32): same 3e-2 bound holds
    assert (h.float().cpu() - ref_h).abs().max().item() <= 3e-2


def test_zero_weights_give_zero_state():
    from cerebralsignalnetworks_b200 import ops
    for dtype in (torch.float32, torch.bfloat16):
        w = [torch.zeros(64, 16).cuda(), torch.zeros(64, 16).cuda(), torch.zeros(64).cuda(), torch.zeros(64).cuda()]
        x = torch.randn(9, 3, 16).cuda().to(dtype)
        h, _, _ = ops.lstm_layer_fwd(x, *w, dtype, True)
        assert torch.all(h.float() == 0)
