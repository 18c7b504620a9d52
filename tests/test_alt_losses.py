"""FeatureDistributionLoss / CosineSimilarityLoss: the oracle restatement against golden vectors the
reference's own classes produced (CPU), and the fused kernels against the same vectors (GPU)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alt_losses.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=True)


def test_oracle_restatement_matches_reference_vectors(gold):
    from oracle.distill import cosine_similarity_loss, feature_distribution_loss
    sched = gold["fd_schedule"]
    for i in range(3):
        s = torch.from_numpy(gold[f"fd_student{i}"]).requires_grad_(True)
        p = torch.from_numpy(gold[f"fd_pred{i}"]).requires_grad_(True)
        loss = feature_distribution_loss(s, torch.from_numpy(gold[f"fd_teacher{i}"]), float(sched[int(gold[f"fd_epoch{i}"])]),
                                         torch.from_numpy(gold[f"fd_label{i}"]), p, float(gold["fd_alpha"]), float(gold["fd_beta"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"fd_loss{i}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), gold[f"fd_dstudent{i}"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(p.grad.numpy(), gold[f"fd_dpred{i}"], rtol=1e-5, atol=1e-9)
    for i in range(2):
        s = torch.from_numpy(gold[f"cos_student{i}"]).requires_grad_(True)
        loss = cosine_similarity_loss(s, torch.from_numpy(gold[f"cos_teacher{i}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"cos_loss{i}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), gold[f"cos_dstudent{i}"], rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_golden_vectors_regenerate_from_the_reference(gold, tmp_path):
    from oracle.ref_import import import_reference
    ref = import_reference("LstmDistillFromDinoV2Train")
    crit = ref.FeatureDistributionLoss(nepochs=12, warmup_teacher_temp=1.5, teacher_temp=0.22, warmup_teacher_temp_epochs=5)
    i = 1
    s = torch.from_numpy(gold[f"fd_student{i}"]).requires_grad_(True)
    p = torch.from_numpy(gold[f"fd_pred{i}"]).requires_grad_(True)
    loss = crit(s, torch.from_numpy(gold[f"fd_teacher{i}"]), int(gold[f"fd_epoch{i}"]), torch.from_numpy(gold[f"fd_label{i}"]), pred_label=p)
    np.testing.assert_allclose(loss.item(), gold[f"fd_loss{i}"], rtol=1e-6)


@pytest.mark.gpu
def test_feature_distribution_loss_kernel_matches_reference_vectors(gold):
    import cerebralsignalnetworks_b200 as csn
    crit = csn.FeatureDistributionLoss(nepochs=12, warmup_teacher_temp=1.5, teacher_temp=0.22, warmup_teacher_temp_epochs=5)
    np.testing.assert_allclose(crit.teacher_temp_schedule, gold["fd_schedule"])
    assert csn.HyperParams.alpha == float(gold["fd_alpha"]) and csn.HyperParams.beta == float(gold["fd_beta"])
    for i in range(3):
        s = torch.from_numpy(gold[f"fd_student{i}"]).cuda().requires_grad_(True)
        p = torch.from_numpy(gold[f"fd_pred{i}"]).cuda().requires_grad_(True)
        loss = crit(s, torch.from_numpy(gold[f"fd_teacher{i}"]).cuda(), int(gold[f"fd_epoch{i}"]),
                    torch.from_numpy(gold[f"fd_label{i}"]).cuda(), pred_label=p)
        (2.0 * loss).backward()  # upstream gradient != 1: the backward scales the saved gradients
        np.testing.assert_allclose(loss.item(), gold[f"fd_loss{i}"], rtol=2e-5)
        scale = np.abs(gold[f"fd_dstudent{i}"]).max()
        # the gradient is p_j (<p,q> - q_j): for a sharp row (p_j ~ 1) the bracket cancels to fp32 round-off in BOTH
        # implementations, so elements are compared against the largest gradient of the batch (1e-3 of it)
        np.testing.assert_allclose(s.grad.cpu().numpy() / 2.0, gold[f"fd_dstudent{i}"], rtol=2e-3, atol=1e-3 * scale)
        np.testing.assert_allclose(p.grad.cpu().numpy() / 2.0, gold[f"fd_dpred{i}"], rtol=2e-3, atol=1e-7)
        assert csn.HyperParams.T == crit.teacher_temp_schedule[int(gold[f"fd_epoch{i}"])]


@pytest.mark.gpu
def test_cosine_similarity_loss_kernel_matches_reference_vectors(gold):
    import cerebralsignalnetworks_b200 as csn
    crit = csn.CosineSimilarityLoss()
    for i in range(2):
        s = torch.from_numpy(gold[f"cos_student{i}"]).cuda().requires_grad_(True)
        loss = crit(s, torch.from_numpy(gold[f"cos_teacher{i}"]).cuda())
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"cos_loss{i}"], rtol=2e-5)
        scale = np.abs(gold[f"cos_dstudent{i}"]).max()
        np.testing.assert_allclose(s.grad.cpu().numpy(), gold[f"cos_dstudent{i}"], rtol=2e-3, atol=2e-5 * scale)


@pytest.mark.gpu
def test_model_with_class_head_trains_on_the_live_loss():
    """The live configuration of LstmDistillFromDinoV2Train.py:323-375: include_top=True model, FeatureDistributionLoss on
    (features, class logits), autograd through both heads and the BPTT kernels, compared with the torch-CPU restatement."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.distill import Model as RefModel, feature_distribution_loss
    torch.manual_seed(3)
    B, T, C, H, D = 6, 40, 16, 32, 24
    model = csn.Model(C, H, 1, D, include_top=True, compute_dtype=torch.float32).cuda()
    ref = RefModel(C, H, 1, D, include_top=True)
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, C, generator=g)
    feats = torch.randn(B, D, generator=g)
    label = torch.randint(0, 40, (B,), generator=g)
    crit = csn.FeatureDistributionLoss(nepochs=10, warmup_teacher_temp=1.5, teacher_temp=0.22, warmup_teacher_temp_epochs=5)
    out, cls = model(x.cuda())
    loss = crit(out, feats.cuda(), 2, label.cuda(), pred_label=cls)
    loss.backward()
    r_out, r_cls = ref(x)
    r_loss = feature_distribution_loss(r_out, feats, float(crit.teacher_temp_schedule[2]), label, r_cls)
    r_loss.backward()
    np.testing.assert_allclose(loss.item(), r_loss.item(), rtol=1e-4)
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        scale = q.grad.abs().max().item() + 1e-12
        assert (p.grad.cpu() - q.grad).abs().max().item() <= 5e-3 * scale, n
