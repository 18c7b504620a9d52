"""Host-side mirror of the reference interface: names, signatures, state-dict compatibility, schedules,
loud failure without the CUDA path."""
import inspect

import numpy as np
import pytest
import torch

import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200 import _lib
from oracle import distill as od


def test_model_signature_and_state_dict_compat():
    from models.lstm import Model  # the import the reference scripts perform
    sig = inspect.signature(Model.__init__)
    assert list(sig.parameters)[1:6] == ["input_size", "lstm_size", "lstm_layers", "output_size", "include_top"]
    torch.manual_seed(3)
    ours = Model(input_size=12, lstm_size=16, lstm_layers=2, output_size=10, include_top=True)
    torch.manual_seed(3)
    ref = od.Model(input_size=12, lstm_size=16, lstm_layers=2, output_size=10, include_top=True)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert set(sd_o) == set(sd_r)
    for k in sd_r:  # same seed -> bit-identical initial weights
        assert torch.equal(sd_o[k], sd_r[k]), k
    ref.load_state_dict(sd_o)
    # Eval.py:309-313 strips "backbone." from a MultiCropWrapper checkpoint and loads with strict=False
    wrapped = {"backbone." + k: v for k, v in sd_o.items()}
    stripped = {k.replace("backbone.", ""): v for k, v in wrapped.items()}
    Model(12, 16, 2, 10, True).load_state_dict(stripped, strict=False)
    # MultiCropWrapper assigns fc/head
    ours.fc, ours.head = torch.nn.Identity(), torch.nn.Identity()


def test_dino_head_param_names_match_reference_layout(golden):
    g = golden("dino_head.npz")
    head = csn.DINOHead(16, 24, nlayers=3, hidden_dim=32, bottleneck_dim=8)
    names = {n for n, _ in head.named_parameters()}
    assert names == {k[len("param."):] for k in g.files if k.startswith("param.")}
    assert not head.last_layer.weight_g.requires_grad
    assert torch.all(head.last_layer.weight_g == 1)
    for n, p in head.named_parameters():
        assert tuple(p.shape) == g["param." + n].shape, n


def test_dino_head_with_batchnorm_has_the_reference_state_layout():
    """DINOHead(use_bn=True) (LstmDistillation.py:72-80): same module order and state_dict keys / shapes as the restated
    reference head (itself checked against the reference class in tests/test_oracle_distill.py), so checkpoints interchange."""
    import torch
    from oracle import distill as od
    from cerebralsignalnetworks_b200.dino import DINOHead
    ours = DINOHead(16, 24, use_bn=True, hidden_dim=32, bottleneck_dim=8)
    ref = od.DINOHead(16, 24, use_bn=True, hidden_dim=32, bottleneck_dim=8)
    so, sr = ours.state_dict(), ref.state_dict()
    assert list(so.keys()) == list(sr.keys())
    assert all(tuple(so[k].shape) == tuple(sr[k].shape) and so[k].dtype == sr[k].dtype for k in so)
    ours.load_state_dict(sr)
    assert torch.equal(ours.mlp[1].running_var, ref.mlp[1].running_var)


def test_dino_loss_signature_and_schedule(golden):
    g = golden("dino_loss_single.npz")
    sig = inspect.signature(csn.DINOLoss.__init__)
    assert list(sig.parameters)[1:9] == ["out_dim", "ncrops", "warmup_teacher_temp", "teacher_temp",
                                         "warmup_teacher_temp_epochs", "nepochs", "student_temp", "center_momentum"]
    crit = csn.DINOLoss(48, 4, 1.5, 0.22, 5, 12)
    np.testing.assert_array_equal(crit.teacher_temp_schedule, g["schedule"])
    assert crit.center.shape == (1, 48) and "center" in crit.state_dict()
    assert list(inspect.signature(crit.forward).parameters) == ["student_output", "teacher_output", "epoch"]
    with pytest.raises(ValueError):  # Q5: nepochs < warmup epochs makes np.ones(negative) raise, as in the reference
        csn.DINOLoss(8, 1, 1.5, 0.22, 50, 10)


def test_eegfilters_attributes(golden):
    g = golden("filters.npz")
    f = csn.EEGFilters(1000.0)
    np.testing.assert_allclose([f.low_cutoff, f.high_cutoff, f.fs, f.low_cutoff_norm, f.high_cutoff_norm],
                               g["eegfilters_attrs"], rtol=0, atol=0)
    np.testing.assert_allclose(f.sos(5.0, 95.0, 4), g["sos_5_95"], rtol=1e-12)
    assert f.sos(order=3).shape == (3, 6) and f.sos(order=5).shape == (5, 6)


def test_no_cpu_fallback():
    """The product path refuses to run without the CUDA device instead of silently computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from models.lstm import Model
    m = Model(8, 8, 1, 8, include_top=False)
    with pytest.raises(_lib.CsnError):
        m(torch.zeros(2, 5, 8))
    with pytest.raises(_lib.CsnError):
        csn.DINOLoss(8, 1, 1.5, 0.22, 2, 4)(torch.zeros(2, 8), torch.zeros(2, 8), 0)
    with pytest.raises(_lib.CsnError):
        csn.EEGFilters(1000.0).apply(torch.zeros(1, 1, 64))


def test_product_does_not_import_oracle():
    import os
    import re
    root = os.path.dirname(os.path.abspath(csn.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_cli_flags():
    from cerebralsignalnetworks_b200 import cli
    a = cli.build_parser().parse_args([])
    assert a.batch_size == 16 and a.num_epochs == 100 and a.learning_rate == 0.001 and a.seed == 43
    a = cli.build_parser().parse_args(["--batch_size", "256", "--num_epochs", "3"])
    assert a.batch_size == 256 and a.num_epochs == 3


def test_schedules_match_reference_goldens(golden, tmp_path):
    """cosine_scheduler against vectors produced by utils/utils.py:187-198; schedule application, last-layer freeze and
    checkpoint round trip as LstmDistillation.py:540-544, :602, :634-646 use them."""
    import torch
    from cerebralsignalnetworks_b200 import schedules as S
    g = golden("utils.npz")
    np.testing.assert_allclose(S.cosine_scheduler(0.0005, 1e-6, 10, 7, warmup_epochs=2), g["cos_a"], rtol=1e-12)
    np.testing.assert_allclose(S.cosine_scheduler(0.04, 0.4, 5, 3), g["cos_b"], rtol=1e-12)
    np.testing.assert_allclose(S.cosine_scheduler(0.996, 1.0, 4, 5), g["cos_c"], rtol=1e-12)

    class Opt:
        param_groups = [{"lr": 0.0, "weight_decay": 0.0}, {"lr": 0.0, "weight_decay": 0.0}]
    S.apply_schedules(Opt, 3, g["cos_a"], g["cos_b"])
    assert Opt.param_groups[0] == {"lr": float(g["cos_a"][3]), "weight_decay": float(g["cos_b"][3])}
    assert Opt.param_groups[1] == {"lr": float(g["cos_a"][3]), "weight_decay": 0.0}

    net = torch.nn.ModuleDict({"mlp": torch.nn.Linear(3, 3), "last_layer": torch.nn.Linear(3, 2)})
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    S.cancel_gradients_last_layer(0, net, freeze_last_layer=1)
    assert net["last_layer"].weight.grad is None and net["mlp"].weight.grad is not None
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    S.cancel_gradients_last_layer(1, net, freeze_last_layer=1)
    assert net["last_layer"].weight.grad is not None

    path = str(tmp_path / "checkpoint.pth")
    S.save_checkpoint(path, student=net, epoch=7)
    other = torch.nn.ModuleDict({"mlp": torch.nn.Linear(3, 3), "last_layer": torch.nn.Linear(3, 2)})
    run = {"epoch": 0}
    assert S.restart_from_checkpoint(path, run_variables=run, student=other) and run["epoch"] == 7
    assert torch.equal(other["mlp"].weight, net["mlp"].weight)
    assert S.restart_from_checkpoint(str(tmp_path / "missing.pth")) is False


def test_resident_dataset_host_side_batching_and_index_checks():
    """The host half of DeviceEEGDataset (no device needed): epoch batches with the DistributedSampler + drop_last split
    of LstmDistillation.py:406-414, and the index validation step_from_dataset relies on (the device does not re-check)."""
    import pytest
    import torch
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    ds = DeviceEEGDataset.__new__(DeviceEEGDataset)   # host logic only: no upload
    ds.N = 22
    full = list(ds.epoch_batches(4, shuffle=False))
    assert len(full) == 5 and torch.equal(torch.cat(full), torch.arange(20))            # drop_last
    assert len(list(ds.epoch_batches(4, shuffle=False, drop_last=False))) == 6
    shares = [list(ds.epoch_batches(4, shuffle=False, rank=r, world=2)) for r in range(2)]
    assert len(shares[0]) == len(shares[1]) == 2                                       # 22 // (4 * 2) global batches
    for k in range(2):  # every global batch of 8 is split contiguously over the ranks, no trial twice
        both = torch.cat([shares[0][k], shares[1][k]])
        assert torch.equal(both, torch.arange(8 * k, 8 * k + 8))
    a = list(ds.epoch_batches(4, generator=torch.Generator().manual_seed(1)))
    b = list(ds.epoch_batches(4, generator=torch.Generator().manual_seed(1)))
    assert all(torch.equal(x, y) for x, y in zip(a, b))                                # ranks agree given the same seed
    assert len(set(torch.cat(a).tolist())) == 20
    idx = ds.check_indices([0, 21, -22, 5])
    assert idx.dtype == torch.int64 and idx.device.type == "cpu" and idx.tolist() == [0, 21, -22, 5]
    from cerebralsignalnetworks_b200 import CsnError
    for bad, exc in (([22], IndexError), ([-23], IndexError), ([[0, 1]], CsnError)):
        with pytest.raises(exc):
            ds.check_indices(bad)


def test_dino_cli_has_the_reference_flags_and_defaults():
    """cli_dino mirrors LstmDistillation.py:187-348 (the flags SURVEY.md section 5 lists, with the reference's defaults)."""
    from cerebralsignalnetworks_b200 import cli_dino
    a = cli_dino.build_parser().parse_args([])
    assert (a.batch_size_per_gpu, a.epochs, a.out_dim) == (8, 200, 384)
    assert (a.warmup_teacher_temp, a.teacher_temp, a.warmup_teacher_temp_epochs) == (0.04, 0.04, 30)
    assert (a.lr, a.min_lr, a.warmup_epochs, a.weight_decay, a.weight_decay_end) == (0.0005, 1e-06, 10, 0.04, 0.4)
    assert (a.clip_grad, a.freeze_last_layer, a.momentum_teacher, a.local_crops_number, a.seed) == (3.0, 1, 0.996, 4, 43)
    b = cli_dino.build_parser().parse_args(["--batch_size_per_gpu", "64", "--epochs", "100", "--out_dim", "65536",
                                            "--norm_last_layer", "false"])
    assert (b.batch_size_per_gpu, b.epochs, b.out_dim, b.norm_last_layer) == (64, 100, 65536, False)
