"""Modules under autograd and the fused DistillTrainStep vs the CPU oracle / reference-generated goldens."""
import numpy as np
import pytest
import torch

from oracle import distill as od

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.array(a))


def _cos(a, b):
    return float((a.flatten() @ b.flatten()) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("tag,cfg", [("l1", (16, 32, 1, 24, False)), ("l2top", (12, 32, 2, 20, True))])
def test_model_autograd_fp32_golden(golden, tag, cfg):
    """Drop-in Model + DINOLoss under torch autograd, fp32 mode, against the golden step (reference DINOLoss)."""
    import cerebralsignalnetworks_b200 as csn
    g = golden(f"distill_step_{tag}.npz")
    C, H, L, D, top = cfg
    model = csn.Model(C, H, L, D, include_top=top, compute_dtype=torch.float32).cuda()
    model.load_state_dict({k[3:]: _t(g[k]) for k in g.files if k.startswith("w0.")})
    crit = csn.DINOLoss(D, 1, 1.5, 0.22, 5, 10).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = _t(g["filtered"]).transpose(1, 2).contiguous().cuda()
    y = model(x)
    emb = y[0] if isinstance(y, tuple) else y
    np.testing.assert_allclose(emb.detach().cpu().numpy(), g["emb"], rtol=1e-4, atol=1e-5)
    if top:
        np.testing.assert_allclose(y[1].detach().cpu().numpy(), g["cls"], rtol=1e-4, atol=1e-5)
    loss = crit(emb, _t(g["feats"]).cuda(), 1)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
    for n, p in model.named_parameters():
        ref = g["g." + n]
        if p.grad is None:
            assert np.all(ref == 0), n
            continue
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 2e-3 * scale, n
    np.testing.assert_allclose(crit.center.cpu().numpy(), g["center_after"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag,cfg", [("l1", (16, 32, 1, 24, False)), ("l2top", (12, 32, 2, 20, True))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_train_step_golden(golden, tag, cfg, dtype):
    """DistillTrainStep (filter -> encoder -> loss -> BPTT -> Adam) from RAW trials vs the golden post-step weights."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos
    g = golden(f"distill_step_{tag}.npz")
    C, H, L, D, top = cfg
    if dtype == torch.bfloat16 and (C % 8 or H % 8):
        pytest.skip("bf16 path needs sizes that are multiples of 8")
    model = csn.Model(C, H, L, D, include_top=top, compute_dtype=dtype).cuda()
    model.load_state_dict({k[3:]: _t(g[k]) for k in g.files if k.startswith("w0.")})
    crit = csn.DINOLoss(D, 1, 1.5, 0.22, 5, 10).cuda()
    step = csn.DistillTrainStep(model, crit, lr=1e-3, sos=design_bandpass_sos(5.0, 95.0, 1000.0, 4))
    loss = step.step(_t(g["eeg"]).cuda(), _t(g["feats"]).cuda(), epoch=1)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert abs(loss.item() - float(g["loss"])) <= tol * abs(float(g["loss"]))
    np.testing.assert_allclose(crit.center.cpu().numpy(), g["center_after"], rtol=1e-5, atol=1e-6)
    for n, p in model.named_parameters():
        grad = step.grad_of(p).cpu()
        ref = _t(g["g." + n])
        if ref.abs().max() == 0:
            assert grad.abs().max() == 0, n
            continue
        if dtype == torch.float32:
            assert (grad - ref).abs().max().item() <= 2e-3 * ref.abs().max().item(), n
            # Adam's first step moves every weight by ~lr*sign(g); compare the post-step weights
            np.testing.assert_allclose(p.detach().cpu().numpy(), g["w1." + n], rtol=1e-4, atol=2e-5)
        else:
            assert _cos(grad, ref) >= 0.99, (n, _cos(grad, ref))


def test_adam_kernel_matches_torch():
    from cerebralsignalnetworks_b200 import ops
    torch.manual_seed(0)
    for decoupled, wd in ((False, 0.0), (False, 0.01), (True, 0.05)):
        p0 = torch.randn(10007)
        ref_p = p0.clone().requires_grad_(True)
        opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref_p], lr=3e-3, weight_decay=wd)
        pad = (10007 + 3) // 4 * 4
        p = torch.zeros(pad).cuda(); p[:10007] = p0
        m = torch.zeros(pad).cuda(); v = torch.zeros(pad).cuda()
        for s in range(1, 6):
            gr = torch.randn(10007)
            ref_p.grad = gr.clone()
            opt.step()
            gpad = torch.zeros(pad).cuda(); gpad[:10007] = gr * 4.0
            ops.adam_step(p, gpad, m, v, 3e-3, weight_decay=wd, decoupled=decoupled, step=s, grad_scale=0.25)
        np.testing.assert_allclose(p[:10007].cpu().numpy(), ref_p.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_dino_head_golden(golden):
    import cerebralsignalnetworks_b200 as csn
    g = golden("dino_head.npz")
    head = csn.DINOHead(16, 24, nlayers=3, hidden_dim=32, bottleneck_dim=8).cuda()
    head.load_state_dict({k[len("param."):]: _t(g[k]) for k in g.files if k.startswith("param.")})
    x = _t(g["x"]).cuda().requires_grad_(True)
    y = head(x)
    y.backward(_t(g["gy"]).cuda())
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["gx"], rtol=1e-3, atol=1e-5)
    for n, p in head.named_parameters():
        if p.requires_grad:
            np.testing.assert_allclose(p.grad.cpu().numpy(), g["grad." + n], rtol=1e-3, atol=1e-5, err_msg=n)


def test_dino_head_with_batchnorm_golden(golden):
    """DINOHead(use_bn=True) against vectors the reference's own class produced (oracle/make_golden.py::golden_dino_head_bn,
    LstmDistillation.py:65-99 in training mode): output, input / parameter gradients, running statistics after the step."""
    import cerebralsignalnetworks_b200 as csn
    g = golden("dino_head_bn.npz")
    head = csn.DINOHead(16, 24, use_bn=True, nlayers=3, hidden_dim=32, bottleneck_dim=8).cuda()
    head.load_state_dict({k[len("state0."):]: _t(g[k]) for k in g.files if k.startswith("state0.")})
    head.train()
    x = _t(g["x"]).cuda().requires_grad_(True)
    y = head(x)
    y.backward(_t(g["gy"]).cuda())
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["y"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["gx"], rtol=2e-3, atol=2e-5)
    for n, p in head.named_parameters():
        if p.requires_grad:
            np.testing.assert_allclose(p.grad.cpu().numpy(), g["grad." + n], rtol=2e-3, atol=2e-5, err_msg=n)
    for n, b in head.named_buffers():
        np.testing.assert_allclose(b.cpu().numpy(), g["buf1." + n], rtol=1e-5, atol=1e-6, err_msg=n)


@pytest.mark.parametrize("M,N", [(384, 2048), (7, 33), (2, 1), (600, 100)])
def test_batchnorm_matches_torch(M, N):
    """csn_batchnorm_fwd / _bwd against torch.nn.BatchNorm1d (what the reference's DINOHead(use_bn=True) calls,
    LstmDistillation.py:72-80): training statistics, running statistics, gradients, then evaluation mode."""
    from cerebralsignalnetworks_b200.dino import BatchNorm1d
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=g) * 2.0 + 0.5
    dy = torch.randn(M, N, generator=g)
    ref = torch.nn.BatchNorm1d(N).double()
    ours = BatchNorm1d(N).cuda()
    with torch.no_grad():
        w, b = torch.randn(N, generator=g), torch.randn(N, generator=g)
        ref.weight.copy_(w); ref.bias.copy_(b); ours.weight.copy_(w); ours.bias.copy_(b)
    for _ in range(2):  # two training steps: the running statistics move twice
        xr = x.double().requires_grad_(True)
        yr = ref(xr); yr.backward(dy.double())
        xo = x.cuda().requires_grad_(True)
        ours.zero_grad(); ref_w_grad = ref.weight.grad.clone(); ref_b_grad = ref.bias.grad.clone(); ref.zero_grad()
        yo = ours(xo); yo.backward(dy.cuda())
        np.testing.assert_allclose(yo.detach().cpu().numpy(), yr.detach().float().numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(xo.grad.cpu().numpy(), xr.grad.float().numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ours.weight.grad.cpu().numpy(), ref_w_grad.float().numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(ours.bias.grad.cpu().numpy(), ref_b_grad.float().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ours.running_mean.cpu().numpy(), ref.running_mean.float().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ours.running_var.cpu().numpy(), ref.running_var.float().numpy(), rtol=1e-5, atol=1e-6)
    assert int(ours.num_batches_tracked) == 2
    ref.eval(); ours.eval()
    xe = x.cuda().requires_grad_(True)
    ye = ours(xe); ye.backward(dy.cuda())
    xr = x.double().requires_grad_(True)
    yr = ref(xr); yr.backward(dy.double())
    np.testing.assert_allclose(ye.detach().cpu().numpy(), yr.detach().float().numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(xe.grad.cpu().numpy(), xr.grad.float().numpy(), rtol=1e-4, atol=2e-5)


def test_dino_head_with_batchnorm_matches_a_torch_head():
    """DINOHead(use_bn=True): the layer layout of LstmDistillation.py:65-88 with nn.BatchNorm1d after the hidden Linears,
    against the same stack built from torch modules with the same weights (fp32 mode)."""
    import cerebralsignalnetworks_b200 as csn
    torch.manual_seed(5)
    head = csn.DINOHead(24, 96, use_bn=True, nlayers=3, hidden_dim=64, bottleneck_dim=32).cuda()
    names = [type(m).__name__ for m in head.mlp]
    assert names == ["Linear", "BatchNorm1d", "GELU", "Linear", "BatchNorm1d", "GELU", "Linear"]
    tl = [torch.nn.Linear(24, 64), torch.nn.BatchNorm1d(64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.BatchNorm1d(64),
          torch.nn.GELU(), torch.nn.Linear(64, 32)]
    ref = torch.nn.Sequential(*tl).double()
    with torch.no_grad():
        for i in (0, 3, 6):
            ref[i].weight.copy_(head.mlp[i].weight.cpu()); ref[i].bias.copy_(head.mlp[i].bias.cpu())
    x = torch.randn(40, 24)
    out = head(x.cuda())
    z = torch.nn.functional.normalize(ref(x.double()), dim=-1, p=2)
    v, gw = head.last_layer.weight_v.detach().cpu().double(), head.last_layer.weight_g.detach().cpu().double()
    want = z @ (gw * v / v.norm(dim=1, keepdim=True)).t()
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.detach().float().numpy(), rtol=2e-4, atol=2e-5)
    out.square().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in head.named_parameters() if p.requires_grad)
    sd = head.state_dict()
    assert "mlp.1.running_mean" in sd and "mlp.1.num_batches_tracked" in sd


def test_multicrop_student_step_runs_like_the_reference_loop():
    """LstmDistillation.py:577-593 shape flow: 2 global + 4 local crops through MultiCropWrapper(Model, DINOHead),
    stacked [6,B,K] vs [2,B,K], reference multi-crop loss; compared with the CPU oracle on identical weights."""
    import cerebralsignalnetworks_b200 as csn
    torch.manual_seed(5)
    B, C, H, K = 3, 16, 32, 64
    ours_b = csn.Model(C, H, 2, H, include_top=False, compute_dtype=torch.float32)
    ours_h = csn.DINOHead(H, K, hidden_dim=48, bottleneck_dim=16)
    ref_b = od.Model(C, H, 2, H, include_top=False)
    ref_h = od.DINOHead(H, K, hidden_dim=48, bottleneck_dim=16)
    ref_b.load_state_dict(ours_b.state_dict()); ref_h.load_state_dict(ours_h.state_dict())
    student = csn.MultiCropWrapper(ours_b, ours_h).cuda()
    ref_student = od.MultiCropWrapper(ref_b, ref_h)
    g = torch.Generator().manual_seed(6)
    views = [torch.randn(B, 30, C, generator=g) for _ in range(2)] + [torch.randn(B, 20, C, generator=g) for _ in range(4)]
    teacher_out = torch.randn(2, B, K, generator=g)
    crit, ref_crit = csn.DINOLoss(K, 6, 0.04, 0.04, 3, 8).cuda(), od.DINOLossMultiCrop(K, 6, 0.04, 0.04, 3, 8)
    so = torch.stack([student(v.cuda()) for v in views], dim=0)
    ro = torch.stack([ref_student(v) for v in views], dim=0)
    np.testing.assert_allclose(so.detach().cpu().numpy(), ro.detach().numpy(), rtol=1e-3, atol=1e-5)
    l, rl = crit(so, teacher_out.cuda(), 0), ref_crit(ro, teacher_out, 0)
    l.backward(); rl.backward()
    np.testing.assert_allclose(l.item(), rl.item(), rtol=1e-4)
    for (n, p), (_, q) in zip(student.named_parameters(), ref_student.named_parameters()):
        if q.grad is None:
            continue
        scale = q.grad.abs().max().item() + 1e-12
        assert (p.grad.cpu() - q.grad).abs().max().item() <= 5e-3 * scale, n


def test_cuda_graph_step_matches_eager_steps():
    """The replayed CUDA graph (device-side Adam step counter, static input buffers) must track eager stepping."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    runs = {}
    for graphed in (False, True):
        torch.manual_seed(11)
        model = csn.Model(16, 32, 1, 24, include_top=False, compute_dtype=torch.float32).cuda()
        crit = csn.DINOLoss(24, 1, 1.5, 0.22, 5, 10).cuda()
        step = csn.DistillTrainStep(model, crit, lr=1e-2, sos=sos, use_cuda_graph=graphed)
        g = torch.Generator(device="cuda").manual_seed(3)
        losses = []
        for i in range(6):
            eeg = torch.randn(4, 16, 48, device="cuda", generator=g)
            feats = torch.randn(4, 24, device="cuda", generator=g)
            losses.append(float(step.step(eeg, feats, epoch=0 if i < 4 else 1)))  # epoch change -> re-capture
        runs[graphed] = (losses, {n: p.detach().clone() for n, p in model.named_parameters()}, crit.center.clone())
    np.testing.assert_allclose(runs[True][0], runs[False][0], rtol=1e-5)
    for n in runs[False][1]:
        assert torch.allclose(runs[True][1][n], runs[False][1][n], rtol=1e-5, atol=1e-7), n
    assert torch.allclose(runs[True][2], runs[False][2], rtol=1e-6, atol=1e-8)


def test_resident_input_graphs_match_eager_steps():
    """register_inputs(): rotating resident buffers each replay their own captured graph (trials read in place, no
    staging copy) and must track eager stepping over the same sequence of batches, including an epoch change."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    g = torch.Generator(device="cuda").manual_seed(5)
    bufs = [(torch.randn(4, 16, 48, device="cuda", generator=g), torch.randn(4, 24, device="cuda", generator=g)) for _ in range(3)]
    runs = {}
    for resident in (False, True):
        torch.manual_seed(11)
        model = csn.Model(16, 32, 1, 24, include_top=False, compute_dtype=torch.float32).cuda()
        crit = csn.DINOLoss(24, 1, 1.5, 0.22, 5, 10).cuda()
        step = csn.DistillTrainStep(model, crit, lr=1e-2, sos=sos, use_cuda_graph=resident)
        if resident:
            for e, f in bufs:
                step.register_inputs(e, f)
        losses = [float(step.step(*bufs[i % 3], epoch=0 if i < 7 else 1)) for i in range(10)]
        if resident:
            assert len(step._resident_graphs) == 6  # 3 buffers x 2 teacher temperatures
        runs[resident] = (losses, {n: p.detach().clone() for n, p in model.named_parameters()}, crit.center.clone())
    np.testing.assert_allclose(runs[True][0], runs[False][0], rtol=1e-5)
    for n in runs[False][1]:
        assert torch.allclose(runs[True][1][n], runs[False][1][n], rtol=1e-5, atol=1e-7), n
    assert torch.allclose(runs[True][2], runs[False][2], rtol=1e-6, atol=1e-8)


def test_fused_adam_groups_clip_and_ema_match_torch(golden):
    """FusedAdam (AdamW, get_params_groups split, per-parameter clipping, frozen group) + EMATeacher vs torch."""
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200.optim import EMATeacher, FusedAdam, get_params_groups
    torch.manual_seed(21)
    def make():
        torch.manual_seed(21)
        return torch.nn.Sequential(torch.nn.Linear(13, 17), torch.nn.Linear(17, 5)).cuda()
    ours, ref, teacher, ref_teacher = make(), make(), make(), make()
    opt = FusedAdam(get_params_groups(ours), lr=2e-3, weight_decay=0.04, decoupled=True)
    ropt = torch.optim.AdamW(get_params_groups(ref), lr=2e-3, weight_decay=0.04)
    ema = EMATeacher(teacher, opt, ours)
    g = torch.Generator(device="cuda").manual_seed(5)
    for it in range(4):
        x = torch.randn(9, 13, device="cuda", generator=g)
        opt.zero_grad(); ropt.zero_grad()
        (ours(x) ** 2).sum().backward()
        (ref(x) ** 2).sum().backward()
        # reference clip_gradients semantics (utils/utils.py:132-141)
        for p in ref.parameters():
            n = p.grad.norm(2)
            coef = 0.7 / (n + 1e-6)
            if coef < 1:
                p.grad.mul_(coef)
        opt.clip_gradients(0.7)
        opt.step(); ropt.step()
        ema.update(0.9)
        with torch.no_grad():
            for q, k in zip(ref.parameters(), ref_teacher.parameters()):
                k.mul_(0.9).add_(0.1 * q)
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    for a, b in zip(teacher.parameters(), ref_teacher.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # reference clip vectors generated from utils.clip_gradients itself
    gold = golden("utils.npz")
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3)).cuda()
    o2 = FusedAdam(list(lin.parameters()), lr=1e-3)
    for i, p in enumerate(lin.parameters()):
        p.grad.copy_(torch.from_numpy(gold[f"clip_gin{i}"]).cuda())
    sq = o2.clip_gradients(1.5)
    np.testing.assert_allclose(sq.sqrt().cpu().numpy(), gold["clip_norms"], rtol=1e-5)
    for i, p in enumerate(lin.parameters()):
        np.testing.assert_allclose(p.grad.cpu().numpy(), gold[f"clip_gout{i}"], rtol=1e-5, atol=1e-7)


def test_fused_optimizer_sweep_matches_torch_clip_adamw_ema():
    """csn_fused_optim_step: per-parameter clip + AdamW groups with moving lr / wd + a group skipped for the first steps
    (cancel_gradients_last_layer) + EMA teacher, in one sweep, against the same sequence written with torch."""
    from cerebralsignalnetworks_b200.optim import EMATeacher, FusedAdam, get_params_groups
    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(37, 61), torch.nn.Linear(61, 2500), torch.nn.Linear(2500, 3)).cuda()
    ours, ref, teacher, ref_teacher = make(), make(), make(), make()
    def groups(m):  # the last layer in its own group so that it can be frozen, rest split like get_params_groups
        last = list(m[2].parameters())
        rest = get_params_groups(torch.nn.Sequential(m[0], m[1]))
        return rest + [{"params": last}]
    opt = FusedAdam(groups(ours), lr=1e-3, weight_decay=0.04, decoupled=True)
    ropt = torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=0.04)
    ema = EMATeacher(teacher, opt, ours)
    g = torch.Generator(device="cuda").manual_seed(9)
    world = 4  # gradients arrive as a SUM over ranks: the sweep folds the 1 / world in before the clip
    for it in range(6):
        lr, wd, mom = 1e-3 * (1 + it), 0.04 + 0.01 * it, 0.9 + 0.01 * it
        for i, (a, b) in enumerate(zip(opt.param_groups, ropt.param_groups)):
            a["lr"] = b["lr"] = lr
            if i != 1:  # the no-decay group keeps wd 0 (LstmDistillation.py:543)
                a["weight_decay"] = b["weight_decay"] = wd
        frozen = it < 2
        opt.set_group_active(2, not frozen)
        x = torch.randn(11, 37, device="cuda", generator=g)
        opt.zero_grad(); ropt.zero_grad()
        (ours(x).pow(2).sum() * world).backward()
        ref(x).pow(2).sum().backward()
        for p in ref.parameters():  # utils.clip_gradients
            coef = 0.3 / (p.grad.norm(2) + 1e-6)
            if coef < 1:
                p.grad.mul_(coef)
        if frozen:  # cancel_gradients_last_layer: p.grad = None -> torch skips the parameter
            for p in ref[2].parameters():
                p.grad = None
        norms = opt.fused_step(clip=0.3, grad_scale=1.0 / world, ema=ema, ema_momentum=mom, want_norms=True)
        ropt.step()
        with torch.no_grad():
            for q, k in zip(ref.parameters(), ref_teacher.parameters()):
                k.mul_(mom).add_((1 - mom) * q.detach())
        assert torch.isfinite(norms).all()
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6)
    for a, b in zip(teacher.parameters(), ref_teacher.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6)
