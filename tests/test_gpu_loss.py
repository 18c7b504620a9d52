"""DINO loss kernel (fwd + bwd + centre statistics in one pass) vs the oracle and the reference-generated goldens."""
import numpy as np
import pytest
import torch

from oracle import distill as od

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.array(a))


def test_single_view_golden(golden):
    import cerebralsignalnetworks_b200 as csn
    g = golden("dino_loss_single.npz")
    crit = csn.DINOLoss(48, 4, 1.5, 0.22, 5, 12).cuda()
    for step in range(3):
        s = _t(g[f"student{step}"]).cuda().requires_grad_(True)
        loss = crit(s, _t(g[f"teacher{step}"]).cuda(), int(g[f"epoch{step}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"loss{step}"], rtol=2e-5)
        np.testing.assert_allclose(s.grad.cpu().numpy(), g[f"grad{step}"], rtol=1e-3, atol=2e-6)
        np.testing.assert_allclose(crit.center.cpu().numpy(), g[f"center_after{step}"], rtol=1e-5, atol=1e-6)


def test_multicrop_reference_semantics_golden(golden):
    import cerebralsignalnetworks_b200 as csn
    g = golden("dino_loss_multicrop.npz")
    crit = csn.DINOLoss(40, 6, 0.04, 0.07, 4, 10).cuda()
    for step in range(2):
        s = _t(g[f"student{step}"]).cuda().requires_grad_(True)
        loss = crit(s, _t(g[f"teacher{step}"]).cuda(), int(g[f"epoch{step}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"loss{step}"], rtol=5e-5)
        np.testing.assert_allclose(s.grad.cpu().numpy(), g[f"grad{step}"], rtol=2e-3, atol=2e-6)
        assert tuple(crit.center.shape) == g[f"center_after{step}"].shape
        np.testing.assert_allclose(crit.center.cpu().numpy(), g[f"center_after{step}"], rtol=1e-5, atol=1e-6)
    assert torch.all(s.grad[0] == 0)


@pytest.mark.parametrize("B,K", [(1, 4), (5, 37), (16, 384), (256, 384), (64, 768), (3, 2048), (2, 8192), (2, 65536), (3, 4100)])
def test_single_view_vs_oracle(B, K):
    from cerebralsignalnetworks_b200 import ops, _lib
    g = torch.Generator().manual_seed(B * 1000 + K)
    s = torch.randn(B, K, generator=g) * 2
    t = torch.randn(B, K, generator=g)
    c = torch.randn(1, K, generator=g) * 0.1
    crit = od.DINOLossSingleView(K, 1, 1.5, 0.22, 5, 12)
    crit.center = c.clone()
    s_ref = s.clone().requires_grad_(True)
    l_ref = crit(s_ref, t, 2)
    l_ref.backward()
    tau = float(crit.teacher_temp_schedule[2])
    loss, ds, bc = ops.dino_loss_fwd_bwd(s.cuda(), t.cuda(), c.cuda().reshape(-1), 0.1, tau, _lib.DINO_SINGLE)
    np.testing.assert_allclose(loss.item(), l_ref.item(), rtol=5e-5)
    np.testing.assert_allclose(ds.cpu().numpy(), s_ref.grad.numpy(), rtol=2e-3, atol=2e-5 / B)
    np.testing.assert_allclose(bc.cpu().numpy(), t.sum(0).numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("Vs,B,K", [(6, 4, 256), (6, 2, 65536), (3, 5, 100), (2, 1, 32)])
def test_multicrop_vs_oracle(Vs, B, K):
    from cerebralsignalnetworks_b200 import ops, _lib
    g = torch.Generator().manual_seed(Vs + B + K)
    s = torch.randn(Vs, B, K, generator=g)
    t = torch.randn(2, B, K, generator=g)
    crit = od.DINOLossMultiCrop(K, Vs, 0.04, 0.07, 3, 8)
    crit.center = torch.randn(1, B, K, generator=g) * 0.05  # post-first-step [1,B,K] centre
    c0 = crit.center.clone()
    s_ref = s.clone().requires_grad_(True)
    l_ref = crit(s_ref, t, 5)
    l_ref.backward()
    loss, ds, bc = ops.dino_loss_fwd_bwd(s.cuda(), t.cuda(), c0.cuda().reshape(-1), 0.1, 0.07, _lib.DINO_MULTICROP_REF)
    np.testing.assert_allclose(loss.item(), l_ref.item(), rtol=1e-4)
    np.testing.assert_allclose(ds.cpu().numpy(), s_ref.grad.numpy(), rtol=3e-3, atol=1e-7)
    np.testing.assert_allclose(bc.view(B, K).cpu().numpy(), t.sum(0).numpy(), rtol=1e-5, atol=1e-5)


def test_canonical_dino_semantics():
    """Upstream DINO (dino/main_dino.py:428-481): teacher chunk(2), skip v == iq."""
    import torch.nn.functional as F
    from cerebralsignalnetworks_b200 import ops, _lib
    g = torch.Generator().manual_seed(5)
    Vs, B, K = 4, 3, 64
    s = torch.randn(Vs, B, K, generator=g).requires_grad_(True)
    t = torch.randn(2, B, K, generator=g)
    c = torch.randn(1, K, generator=g) * 0.1
    q = F.softmax((t - c) / 0.05, dim=-1)
    total, n = 0, 0
    for iq in range(2):
        for v in range(Vs):
            if v == iq:
                continue
            total = total + torch.sum(-q[iq] * F.log_softmax(s[v] / 0.1, dim=-1), dim=-1).mean()
            n += 1
    total = total / n
    total.backward()
    loss, ds, bc = ops.dino_loss_fwd_bwd(s.detach().cuda(), t.cuda(), c.cuda().reshape(-1), 0.1, 0.05,
                                         _lib.DINO_MULTICROP_CANONICAL)
    np.testing.assert_allclose(loss.item(), total.item(), rtol=1e-4)
    np.testing.assert_allclose(ds.cpu().numpy(), s.grad.numpy(), rtol=3e-3, atol=1e-7)
    np.testing.assert_allclose(bc.cpu().numpy(), t.sum((0, 1)).numpy(), rtol=1e-5, atol=1e-5)


def test_shift_invariance_and_grad_rows_sum_to_zero():
    from cerebralsignalnetworks_b200 import ops, _lib
    g = torch.Generator(device="cuda").manual_seed(1)
    s = torch.randn(64, 768, device="cuda", generator=g)
    t = torch.randn(64, 768, device="cuda", generator=g)
    c = torch.zeros(768, device="cuda")
    l1, d1, _ = ops.dino_loss_fwd_bwd(s, t, c, 0.1, 0.22, _lib.DINO_SINGLE)
    l2, d2, _ = ops.dino_loss_fwd_bwd(s + 3.0, t, c, 0.1, 0.22, _lib.DINO_SINGLE)
    assert abs(l1.item() - l2.item()) < 1e-3 * abs(l1.item())
    assert torch.allclose(d1, d2, atol=1e-6)
    assert d1.sum(dim=-1).abs().max().item() < 1e-5


def test_centre_ema_fixed_point():
    from cerebralsignalnetworks_b200 import ops
    c = torch.full((1000,), 2.0, device="cuda")
    bc = torch.full((1000,), 2.0 * 8, device="cuda")
    ops.center_ema(c, bc, 0.9, 1.0 / 8)
    assert torch.allclose(c, torch.full_like(c, 2.0))
