"""The `loss_fn_kd` family of the older training scripts (rest of SURVEY.md 8f #3): the oracle restatements against golden
vectors produced by the reference's own functions (CPU), and the fused kernel against the same vectors (GPU)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kd_losses.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=True)


def test_oracle_restatements_match_reference_vectors(gold):
    from oracle.distill import kd_loss_hinton, kd_loss_retrieval, neg_cosine_loss
    for i in range(3):
        s = torch.from_numpy(gold[f"ret_student{i}"]).requires_grad_(True)
        loss = kd_loss_retrieval(s, torch.from_numpy(gold[f"ret_teacher{i}"]), float(gold["ret_T"]), float(gold["ret_w_soft"]),
                                 float(gold["ret_w_ce"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"ret_loss{i}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), gold[f"ret_dstudent{i}"], rtol=1e-5, atol=1e-9)
    for i in range(3):
        s = torch.from_numpy(gold[f"hin_student{i}"]).requires_grad_(True)
        loss = kd_loss_hinton(s, torch.from_numpy(gold[f"hin_label{i}"]), torch.from_numpy(gold[f"hin_teacher{i}"]),
                              float(gold[f"hin_T{i}"]), float(gold[f"hin_alpha{i}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"hin_loss{i}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), gold[f"hin_dstudent{i}"], rtol=1e-5, atol=1e-9)
    for i in range(2):
        s = torch.from_numpy(gold[f"neg_student{i}"]).requires_grad_(True)
        loss = neg_cosine_loss(s, torch.from_numpy(gold[f"neg_teacher{i}"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"neg_loss{i}"], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), gold[f"neg_dstudent{i}"], rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_golden_vectors_regenerate_from_the_reference(gold):
    from oracle.ref_import import import_reference
    ret = import_reference("LSTMDistillRetreival")
    i = 1
    loss = ret.loss_fn_kd(torch.from_numpy(gold[f"ret_student{i}"]), None, torch.from_numpy(gold[f"ret_teacher{i}"]), ret.Parameters)
    np.testing.assert_allclose(loss.item(), gold[f"ret_loss{i}"], rtol=1e-6)
    assert float(gold["ret_T"]) == ret.Parameters.temperature


def test_host_mirror_keeps_the_reference_knobs(gold):
    from cerebralsignalnetworks_b200.losses import Parameters
    assert Parameters.temperature == float(gold["ret_T"])
    assert Parameters.soft_target_loss_weight == float(gold["ret_w_soft"])
    assert Parameters.ce_loss_weight == float(gold["ret_w_ce"])


def _check_grad(got, want, rel=2e-3, floor=2e-5):
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=rel, atol=floor * scale)


@pytest.mark.gpu
def test_retrieval_kd_loss_kernel_matches_reference_vectors(gold):
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200.losses import Parameters
    for i in range(3):
        s = torch.from_numpy(gold[f"ret_student{i}"]).cuda().requires_grad_(True)
        loss = csn.loss_fn_kd(s, None, torch.from_numpy(gold[f"ret_teacher{i}"]).cuda(), Parameters)
        (3.0 * loss).backward()  # upstream gradient != 1
        np.testing.assert_allclose(loss.item(), gold[f"ret_loss{i}"], rtol=2e-5)
        _check_grad(s.grad.cpu().numpy() / 3.0, gold[f"ret_dstudent{i}"])


@pytest.mark.gpu
def test_hinton_kd_loss_kernel_matches_reference_vectors(gold):
    import cerebralsignalnetworks_b200 as csn

    class P:
        pass
    for i in range(3):
        P.alpha, P.temperature = float(gold[f"hin_alpha{i}"]), float(gold[f"hin_T{i}"])
        s = torch.from_numpy(gold[f"hin_student{i}"]).cuda().requires_grad_(True)
        loss = csn.loss_fn_kd_hinton(s, torch.from_numpy(gold[f"hin_label{i}"]).cuda(),
                                     torch.from_numpy(gold[f"hin_teacher{i}"]).cuda(), P)
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"hin_loss{i}"], rtol=2e-5)
        _check_grad(s.grad.cpu().numpy(), gold[f"hin_dstudent{i}"])


@pytest.mark.gpu
def test_negative_cosine_loss_matches_reference_vectors(gold):
    import cerebralsignalnetworks_b200 as csn
    for i in range(2):
        s = torch.from_numpy(gold[f"neg_student{i}"]).cuda().requires_grad_(True)
        loss = csn.cosine_similarity_loss(s, torch.from_numpy(gold[f"neg_teacher{i}"]).cuda())
        loss.backward()
        np.testing.assert_allclose(loss.item(), gold[f"neg_loss{i}"], rtol=2e-5, atol=2e-7)
        _check_grad(s.grad.cpu().numpy(), gold[f"neg_dstudent{i}"])


@pytest.mark.gpu
def test_kd_loss_rejects_bad_shapes():
    import cerebralsignalnetworks_b200 as csn
    s = torch.zeros(4, 2048, device="cuda")
    with pytest.raises(csn.CsnError):
        csn.loss_fn_kd(s, None, s.clone())
    with pytest.raises(ValueError):
        csn.loss_fn_kd(torch.zeros(4, 8, device="cuda"), None, torch.zeros(4, 9, device="cuda"))
