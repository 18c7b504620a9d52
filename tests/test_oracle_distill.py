"""Oracle restatements vs the golden vectors generated from the reference's OWN classes
(oracle/make_golden.py), and -- when /root/reference is present -- vs the live reference classes."""
import numpy as np
import pytest
import torch

from oracle import distill as od
from oracle.ref_import import import_reference, reference_available


def _t(a):
    return torch.from_numpy(np.array(a))


def test_dino_single_golden(golden):
    g = golden("dino_loss_single.npz")
    crit = od.DINOLossSingleView(48, 4, 1.5, 0.22, 5, 12)
    np.testing.assert_allclose(crit.teacher_temp_schedule, g["schedule"], rtol=0, atol=0)
    for step in range(3):
        s = _t(g[f"student{step}"]).requires_grad_(True)
        np.testing.assert_allclose(crit.center.numpy(), g[f"center_before{step}"], rtol=1e-6, atol=1e-7)
        loss = crit(s, _t(g[f"teacher{step}"]), int(g[f"epoch{step}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"loss{step}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), g[f"grad{step}"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(crit.center.numpy(), g[f"center_after{step}"], rtol=1e-6, atol=1e-7)


def test_dino_multicrop_golden(golden):
    g = golden("dino_loss_multicrop.npz")
    crit = od.DINOLossMultiCrop(40, 6, 0.04, 0.07, 4, 10)
    for step in range(2):
        s = _t(g[f"student{step}"]).requires_grad_(True)
        loss = crit(s, _t(g[f"teacher{step}"]), int(g[f"epoch{step}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"loss{step}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), g[f"grad{step}"], rtol=1e-5, atol=1e-8)
        assert crit.center.shape == g[f"center_after{step}"].shape  # [1, B, K] quirk (Q3)
        np.testing.assert_allclose(crit.center.numpy(), g[f"center_after{step}"], rtol=1e-6, atol=1e-7)
    assert np.all(g["grad0"][0] == 0)  # student view 0 receives no gradient (Q4)


def test_dino_head_golden(golden):
    g = golden("dino_head.npz")
    head = od.DINOHead(16, 24, nlayers=3, hidden_dim=32, bottleneck_dim=8)
    with torch.no_grad():
        for n, p in head.named_parameters():
            p.copy_(_t(g["param." + n]))
    x = _t(g["x"]).requires_grad_(True)
    y = head(x)
    y.backward(_t(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["gx"], rtol=1e-4, atol=1e-6)
    for n, p in head.named_parameters():
        if p.grad is not None:
            np.testing.assert_allclose(p.grad.numpy(), g["grad." + n], rtol=1e-4, atol=1e-6)


def test_cosine_scheduler_golden(golden):
    g = golden("utils.npz")
    np.testing.assert_allclose(od.cosine_scheduler(0.0005, 1e-6, 10, 7, warmup_epochs=2), g["cos_a"], rtol=1e-12)
    np.testing.assert_allclose(od.cosine_scheduler(0.04, 0.4, 5, 3), g["cos_b"], rtol=1e-12)
    np.testing.assert_allclose(od.cosine_scheduler(0.996, 1.0, 4, 5), g["cos_c"], rtol=1e-12)


@pytest.mark.parametrize("tag,cfg", [("l1", (16, 32, 1, 24, False)), ("l2top", (12, 32, 2, 20, True))])
def test_distill_step_golden(golden, tag, cfg):
    """Restated Model + oracle DINOLoss + Adam reproduces the step recorded with the REFERENCE DINOLoss."""
    g = golden(f"distill_step_{tag}.npz")
    C, H, L, D, top = cfg
    model = od.Model(C, H, L, D, include_top=top)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(_t(g["w0." + n]))
    crit = od.DINOLossSingleView(D, 1, 1.5, 0.22, 5, 10)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = _t(g["filtered"]).transpose(1, 2).contiguous()
    y = model(x)
    emb = y[0] if isinstance(y, tuple) else y
    loss = crit(emb, _t(g["feats"]), 1)
    loss.backward()
    opt.step()
    np.testing.assert_allclose(emb.detach().numpy(), g["emb"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-6)
    for n, p in model.named_parameters():
        if p.grad is not None:
            np.testing.assert_allclose(p.grad.numpy(), g["g." + n], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(p.detach().numpy(), g["w1." + n], rtol=1e-5, atol=1e-7)


def test_synthetic_batch_shapes():
    eeg, feats, labels = od.synthetic_batch(5, 128, 440, 384)
    assert eeg.shape == (5, 128, 440) and feats.shape == (5, 384) and labels.shape == (5,)
    assert labels.min() >= 0 and labels.max() < 40


# ---- live checks against the reference's own classes (build container only) ----
needs_ref = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def pg():
    import os
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29592")
        dist.init_process_group("gloo", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


@needs_ref
def test_live_reference_dino_single(pg):
    ref = import_reference("LstmDistillFromDinoV2Train")
    a = ref.DINOLoss(64, 4, 1.5, 0.22, 5, 20)
    b = od.DINOLossSingleView(64, 4, 1.5, 0.22, 5, 20)
    g = torch.Generator().manual_seed(0)
    for epoch in (0, 2, 9):
        s = torch.randn(7, 64, generator=g)
        t = torch.randn(7, 64, generator=g)
        s1, s2 = s.clone().requires_grad_(True), s.clone().requires_grad_(True)
        l1, l2 = a(s1, t, epoch), b(s2, t, epoch)
        l1.backward(); l2.backward()
        assert torch.allclose(l1, l2, rtol=1e-6) and torch.allclose(s1.grad, s2.grad, rtol=1e-5, atol=1e-8)
        assert torch.allclose(a.center, b.center, rtol=1e-6, atol=1e-8)


@needs_ref
def test_live_reference_dino_multicrop_and_head(pg):
    ref = import_reference("LstmDistillation")
    a = ref.DINOLoss(32, 6, 0.04, 0.04, 3, 8)
    b = od.DINOLossMultiCrop(32, 6, 0.04, 0.04, 3, 8)
    g = torch.Generator().manual_seed(1)
    for epoch in (0, 5):
        s = torch.randn(6, 3, 32, generator=g)
        t = torch.randn(2, 3, 32, generator=g)
        s1, s2 = s.clone().requires_grad_(True), s.clone().requires_grad_(True)
        l1, l2 = a(s1, t, epoch), b(s2, t, epoch)
        l1.backward(); l2.backward()
        assert torch.allclose(l1, l2, rtol=1e-6) and torch.allclose(s1.grad, s2.grad, rtol=1e-5, atol=1e-8)
        assert a.center.shape == b.center.shape and torch.allclose(a.center, b.center, rtol=1e-6, atol=1e-8)
    torch.manual_seed(2)
    h1 = ref.DINOHead(16, 24, hidden_dim=32, bottleneck_dim=8)
    h2 = od.DINOHead(16, 24, hidden_dim=32, bottleneck_dim=8)
    h2.load_state_dict(h1.state_dict())
    x = torch.randn(5, 16)
    assert torch.allclose(h1(x), h2(x), rtol=1e-6, atol=1e-7)
    # the BatchNorm variant of the head (use_bn=True, :72-80): same layout, same training-mode outputs and running statistics
    h1 = ref.DINOHead(16, 24, use_bn=True, hidden_dim=32, bottleneck_dim=8)
    h2 = od.DINOHead(16, 24, use_bn=True, hidden_dim=32, bottleneck_dim=8)
    assert list(h1.state_dict().keys()) == list(h2.state_dict().keys())
    h2.load_state_dict(h1.state_dict())
    xb = torch.randn(9, 16)
    assert torch.allclose(h1(xb), h2(xb), rtol=1e-6, atol=1e-7)
    assert torch.allclose(h1.mlp[1].running_var, h2.mlp[1].running_var)
    # MultiCropWrapper grouping
    class Back(torch.nn.Module):
        def forward(self, x):
            return x.flatten(1)[:, :3]
    w1 = ref.MultiCropWrapper(Back(), torch.nn.Identity())
    w2 = od.MultiCropWrapper(Back(), torch.nn.Identity())
    xs = [torch.randn(2, 4, 16), torch.randn(2, 4, 16), torch.randn(2, 4, 8)]
    assert torch.equal(w1(xs), w2(xs))


def test_oracle_model_matches_the_reference_analogue_classes_on_square_inputs():
    """models.lstm.Model itself is absent from the reference; the two LSTMModel classes it derives from are not
    (LSTMDistillRetreival.py:85-110, LSTMDistill.py:112-142).  On square inputs their leading `x.view(B, C, T)` is the
    identity, and the restated Model must reproduce their outputs and weight gradients (fixture generated by
    oracle/make_golden.py from the reference's own classes)."""
    import os
    import numpy as np
    import torch
    from oracle.distill import Model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_analogues.npz"), allow_pickle=True)

    def load(prefix, model):
        sd = {}
        for k in g.files:
            if k.startswith(prefix + "_w_"):
                name = k[len(prefix) + 3:].replace("fc.", "output.").replace("class_pred.", "classifier.")
                sd[name] = torch.from_numpy(g[k])
        model.load_state_dict(sd)

    x = torch.from_numpy(g["ret_x"])
    m = Model(x.shape[2], int(g["ret_hidden"]), int(g["ret_layers"]), g["ret_y"].shape[1], include_top=False)
    load("ret", m)
    y = m(x)
    np.testing.assert_allclose(y.detach().numpy(), g["ret_y"], rtol=1e-5, atol=1e-6)
    y.pow(2).sum().backward()
    for name, p in m.named_parameters():
        want = g["ret_g_" + name.replace("output.", "fc.")]
        np.testing.assert_allclose(p.grad.numpy(), want, rtol=1e-4, atol=1e-6, err_msg=name)

    x2 = torch.from_numpy(g["dis_x"])
    m2 = Model(x2.shape[2], int(g["dis_hidden"]), int(g["dis_layers"]), g["dis_feat_last"].shape[1], include_top=True,
               n_classes=g["dis_cls_last"].shape[1])
    load("dis", m2)
    feat, cls = m2(x2)
    np.testing.assert_allclose(feat.detach().numpy(), g["dis_feat_last"], rtol=1e-5, atol=1e-6)   # ReLU'd features
    np.testing.assert_allclose(cls.detach().numpy(), g["dis_cls_last"], rtol=1e-5, atol=1e-6)     # class head on pre-ReLU features
