"""Fused projection head + single-view DINO loss (csn_head_dino_fwd_bwd) against the unfused kernels it replaces
(csn_gemm_f32 -> csn_dino_loss_fwd_bwd -> csn_gemm_f32) and against the torch-CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _unfused(h32, w, b, act, teacher, center, ts, tt):
    from cerebralsignalnetworks_b200 import _lib, ops
    from cerebralsignalnetworks_b200.functional import linear_bwd, linear_fwd
    emb, pre = linear_fwd(h32, w, b, act)
    bc = torch.zeros(w.shape[0], device="cuda")
    loss, d_emb, _ = ops.dino_loss_fwd_bwd(emb, teacher, center, ts, tt, _lib.DINO_SINGLE, batch_center=bc)
    d_h, dw, db = linear_bwd(h32, w, pre, d_emb, act, need_dx=True)
    d_pre = d_emb if act == _lib.ACT_NONE else ops.act_bwd(pre, d_emb.contiguous(), act)
    return loss, d_h, d_pre, bc, dw, db


@pytest.mark.parametrize("B,I,K", [(256, 128, 384), (7, 32, 24), (33, 64, 100), (5, 128, 96), (600, 128, 384)])
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_head_matches_unfused_kernels(B, I, K, act, dtype):
    from cerebralsignalnetworks_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 7 + K + act)
    h = (torch.randn(B, I, device="cuda", generator=g) * 0.5).to(dtype)
    w = torch.randn(K, I, device="cuda", generator=g) * 0.2
    b = torch.randn(K, device="cuda", generator=g) * 0.1
    teacher = torch.randn(B, K, device="cuda", generator=g)
    center = torch.randn(K, device="cuda", generator=g) * 0.1
    assert ops.head_dino_supported(B, I, K)
    bc = torch.zeros(K, device="cuda")
    loss, d_h, d_pre = ops.head_dino_fwd_bwd(h, w, b, act, teacher, center, 0.1, 0.07, bc)
    r_loss, r_dh, r_dpre, r_bc, _, _ = _unfused(h.float().contiguous(), w, b, act, teacher, center, 0.1, 0.07)
    np.testing.assert_allclose(loss.item(), r_loss.item(), rtol=2e-5)
    # gradients: tolerance relative to the largest element of the batch (p - q differences cancel for sharp rows)
    for got, want in ((d_h, r_dh), (d_pre, r_dpre)):
        scale = want.abs().max().item()
        assert (got - want).abs().max().item() <= 2e-4 * scale + 1e-12
    # centre statistics: both paths are the SAME fixed-order column sum now (bit-equal), and within the fp32 bound of
    # a B-term sum, B * u * max|t| (u = 2^-24), of the float64 column sum
    assert torch.equal(bc, r_bc)
    exact = teacher.double().sum(0).cpu().numpy()
    bound = B * 2.0 ** -24 * teacher.abs().max().item()
    assert np.abs(bc.cpu().numpy().astype(np.float64) - exact).max() <= bound


def test_fused_head_matches_oracle_autograd():
    from cerebralsignalnetworks_b200 import ops
    from oracle import distill as od
    torch.manual_seed(11)
    B, I, K = 12, 64, 48
    h = torch.randn(B, I) * 0.5
    lin = torch.nn.Linear(I, K)
    teacher = torch.randn(B, K)
    crit = od.DINOLossSingle(K, 1, 1.5, 0.22, 5, 10) if hasattr(od, "DINOLossSingle") else None
    ts, tt = 0.1, 0.22
    center = torch.randn(K) * 0.05
    hh = h.clone().requires_grad_(True)
    emb = lin(hh)
    q = torch.softmax((teacher - center) / tt, dim=-1)
    ref = torch.sum(-q * torch.log_softmax(emb / ts, dim=-1), dim=-1).mean()
    ref.backward()
    bc = torch.zeros(K, device="cuda")
    loss, d_h, d_pre = ops.head_dino_fwd_bwd(h.cuda(), lin.weight.detach().cuda().contiguous(), lin.bias.detach().cuda(), 0,
                                             teacher.cuda(), center.cuda(), ts, tt, bc)
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=2e-5)
    np.testing.assert_allclose(d_h.cpu().numpy(), hh.grad.numpy(), rtol=2e-3, atol=2e-4 * hh.grad.abs().max().item())
    np.testing.assert_allclose(bc.cpu().numpy(), teacher.double().sum(0).numpy(), rtol=0,
                               atol=B * 2.0 ** -24 * teacher.abs().max().item())
    # dW from d_pre: what the train step computes on its side stream
    dw = ops.gemm_f32(d_pre, h.cuda(), True, False)
    np.testing.assert_allclose(dw.cpu().numpy(), lin.weight.grad.numpy(), rtol=2e-3, atol=2e-4 * lin.weight.grad.abs().max().item())


def test_unserved_head_shapes_are_refused():
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200 import ops
    assert not ops.head_dino_supported(8, 512, 768)   # W does not fit in shared memory
    assert not ops.head_dino_supported(8, 100, 64)    # encoder width not a multiple of 32
    assert not ops.head_dino_supported(8, 2048, 8)    # one thread per input column: at most 1024
    z = torch.zeros
    with pytest.raises(csn.CsnError):
        ops.head_dino_fwd_bwd(z(8, 512, device="cuda"), z(768, 512, device="cuda"), z(768, device="cuda"), 0,
                              z(8, 768, device="cuda"), z(768, device="cuda"), 0.1, 0.07, z(768, device="cuda"))
