"""GPU-resident dataset: oracle restatement against vectors produced by the reference's own
EEGDataset (CPU), and the gather kernel against the same vectors (GPU)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=True)


def _loaded(gold):
    raw = gold["eeg_raw"]
    return {"dataset": [{"eeg": torch.from_numpy(raw[i]), "image": i, "label": int(gold["labels"][i]), "subject": 1}
                        for i in range(len(raw))],
            "labels": ["a", "b", "c"], "images": [f"img{i}" for i in range(len(raw))]}


def test_oracle_item_transform_matches_reference_dataset(gold):
    from oracle.dataset import dataset_scalars, item_transform
    loaded = _loaded(gold)
    mean, std = dataset_scalars(loaded["dataset"])
    assert mean == pytest.approx(float(gold["mean_plain"]), rel=1e-12) and std == pytest.approx(float(gold["std_plain"]), rel=1e-12)
    lo, hi = int(gold["time_low"]), int(gold["time_high"])
    for i, it in enumerate(loaded["dataset"]):
        np.testing.assert_array_equal(item_transform(it["eeg"], lo, hi).numpy(), gold["items_plain"][i])
        np.testing.assert_allclose(item_transform(it["eeg"], lo, hi, mean, std).numpy(), gold["items_norm"][i], rtol=1e-6, atol=1e-7)


def test_oracle_channel_selection_matches_reference_dataset(gold):
    from oracle.dataset import item_transform_channels
    loaded = _loaded(gold)
    lo, hi, sel = int(gold["time_low"]), int(gold["time_high"]), [int(c) for c in gold["filter_channels"]]
    for i, it in enumerate(loaded["dataset"]):
        np.testing.assert_array_equal(item_transform_channels(it["eeg"], lo, hi, sel).numpy(), gold["items_chsel"][i])
        np.testing.assert_allclose(item_transform_channels(it["eeg"], lo, hi, sel, True).numpy(), gold["items_chsel_norm"][i],
                                   rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("chnorm", [False, True])
def test_device_dataset_channel_selection_matches_reference_items(gold, chnorm):
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    lo, hi, sel = int(gold["time_low"]), int(gold["time_high"]), [int(c) for c in gold["filter_channels"]]
    ds = DeviceEEGDataset(_loaded(gold), time_low=lo, time_high=hi, filter_channels=sel, apply_channel_wise_norm=chnorm)
    assert (ds.C, ds.samples, len(ds)) == (len(sel), hi - lo, 7)
    assert ds.mean == pytest.approx(float(gold["mean_plain"]), rel=1e-12)   # scalars of the ORIGINAL trials
    got = ds.batch([6, 1, -7])[0].cpu().numpy()                             # [B, len(sel), T], the reference's item layout
    want = gold["items_chsel_norm" if chnorm else "items_chsel"][[6, 1, 0]]
    if chnorm:
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    else:
        np.testing.assert_array_equal(got, want)
    with pytest.raises(IndexError):
        DeviceEEGDataset(_loaded(gold), time_low=lo, time_high=hi, filter_channels=[0, 6])


@pytest.mark.gpu
@pytest.mark.parametrize("norm", [False, True])
def test_device_dataset_batches_match_reference_items(gold, norm):
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    lo, hi = int(gold["time_low"]), int(gold["time_high"])
    ds = DeviceEEGDataset(_loaded(gold), time_low=lo, time_high=hi, apply_norm_with_stds_and_means=norm)
    assert len(ds) == 7 and ds.samples == hi - lo
    assert ds.mean == pytest.approx(float(gold["mean_plain"]), rel=1e-12) and ds.std == pytest.approx(float(gold["std_plain"]), rel=1e-12)
    items = gold["items_norm" if norm else "items_plain"]            # [N, T, C]
    idx = [3, 0, 6, 6, -1, 2]
    btc, labels, img = ds.batch_btc(idx)
    bct, labels2, _ = ds.batch(idx)
    want = items[[3, 0, 6, 6, 6, 2]]
    if norm:
        np.testing.assert_allclose(btc.cpu().numpy(), want, rtol=1e-6, atol=1e-6)
    else:
        np.testing.assert_array_equal(btc.cpu().numpy(), want)
    assert torch.equal(bct, btc.transpose(1, 2))
    np.testing.assert_array_equal(labels.cpu().numpy(), gold["labels"][[3, 0, 6, 6, 6, 2]])
    assert torch.equal(labels, labels2) and img.tolist() == [3, 0, 6, 6, 6, 2]
    full = DeviceEEGDataset(_loaded(gold), time_low=0, time_high=50)   # the reference's default state: no crop
    np.testing.assert_array_equal(full.batch(range(7))[0].cpu().numpy(), gold["eeg_raw"].astype(np.float32))
    with pytest.raises(IndexError):
        ds.batch([7])
    assert ds.batch([])[0].shape == (0, 6, hi - lo)


@pytest.mark.gpu
def test_device_dataset_feeds_the_train_step(gold):
    """Wide trials (more than one channel tile, T not a multiple of 32) + epoch batching + one fused step."""
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    g = torch.Generator().manual_seed(1)
    N, C, T_raw = 20, 40, 75
    loaded = {"dataset": [{"eeg": torch.randn(C, T_raw, generator=g), "image": i, "label": i % 4} for i in range(N)],
              "labels": list("abcd"), "images": [str(i) for i in range(N)]}
    ds = DeviceEEGDataset(loaded, time_low=7, time_high=71)
    batches = list(ds.epoch_batches(8, generator=torch.Generator().manual_seed(0)))
    assert len(batches) == 2 and all(len(b) == 8 for b in batches)
    shares = [list(ds.epoch_batches(4, shuffle=False, rank=r, world=2)) for r in range(2)]
    assert shares[0][0].tolist() == [0, 1, 2, 3] and shares[1][0].tolist() == [4, 5, 6, 7]
    eeg, labels, _ = ds.batch(batches[0])
    ref = torch.stack([loaded["dataset"][i]["eeg"][:, 7:71] for i in batches[0].tolist()])
    assert torch.equal(eeg.cpu(), ref)
    model = csn.Model(C, 32, 1, 24, include_top=False, compute_dtype=torch.float32).cuda()
    step = csn.DistillTrainStep(model, csn.DINOLoss(24, 1, 1.5, 0.22, 5, 10).cuda(), lr=1e-3,
                                sos=csn.EEGFilters(1000.0).sos(5.0, 95.0, 4), use_cuda_graph=False)
    loss = step.step(eeg, torch.randn(8, 24, device="cuda"), 0)
    assert torch.isfinite(loss)


@pytest.mark.gpu
def test_cli_trains_from_a_pth_dataset(tmp_path, capsys):
    """--batch_size / --num_epochs CLI (LstmDistillFromDinoV2Train.py:150-231) on a ConvertToPth-style file served from HBM."""
    from cerebralsignalnetworks_b200 import cli
    g = torch.Generator().manual_seed(2)
    N, C, T_raw = 24, 16, 120
    loaded = {"dataset": [{"eeg": torch.randn(C, T_raw, generator=g), "image": i % 10, "label": i % 4, "subject": 1} for i in range(N)],
              "labels": list("abcd"), "images": [f"n0_{i}" for i in range(10)], "image_features": torch.randn(10, 24, generator=g)}
    path = str(tmp_path / "spampinato-synthetic.pth")
    torch.save(loaded, path)
    cli.main(["--eeg_dataset", path, "--batch_size", "8", "--num_epochs", "2", "--input_size", "16", "--lstm_size", "32",
              "--output_size", "24", "--time_low", "10", "--time_high", "110", "--log_dir", str(tmp_path / "logs"),
              "--warmup_teacher_temp_epochs", "1", "--precision", "fp32"])
    out = capsys.readouterr().out
    assert "EPOCH 0 train_loss" in out and "EPOCH 1 train_loss" in out
    sd = torch.load(str(tmp_path / "logs" / "lstm_dinov2_last.pth"))
    assert "lstm.weight_ih_l0" in sd and all(torch.isfinite(v).all() for v in sd.values())


@pytest.mark.gpu
@pytest.mark.parametrize("norm", [False, True])
def test_filter_with_fused_gather_is_bit_identical_to_gather_then_filter(norm):
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200 import ops
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    torch.manual_seed(4)
    N, C, T_raw, lo, hi = 23, 64, 120, 8, 108
    ds = DeviceEEGDataset.from_tensor(torch.randn(N, C, T_raw) * 3 + 1, time_low=lo, time_high=hi, apply_norm_with_stds_and_means=norm)
    assert ds.fused_filter_ok()
    sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 4)
    idx = torch.tensor([22, 0, -1, 7, 7, 13, -23, 4, 19])
    mean, std = ds.norm_scalars()
    for dtype in (torch.float32, torch.bfloat16):
        fused = ops.sosfilt_gather(ds.eeg, idx.cuda(), lo, hi, sos, mean, std, out_layout="TBC", out_dtype=dtype)
        eeg, _, _ = ds.batch(idx)
        two = ops.sosfilt(eeg, sos, out_layout="TBC", out_dtype=dtype)
        assert fused.shape == two.shape == (hi - lo, idx.numel(), C)
        assert torch.equal(fused, two)
    # shapes the fused path does not serve are refused, not silently mis-read
    bad = DeviceEEGDataset.from_tensor(torch.randn(5, 24, 40), time_low=4, time_high=36)
    assert not bad.fused_filter_ok()
    with pytest.raises(csn.CsnError):
        ops.sosfilt_gather(bad.eeg, torch.tensor([0, 1]).cuda(), 4, 36, sos)


@pytest.mark.gpu
@pytest.mark.parametrize("C", [16, 32])   # 16: gather kernel + ordinary step; 32: gather fused into the filter
def test_step_from_resident_dataset_equals_step_on_gathered_batch(C):
    """DistillTrainStep.step_from_dataset (indices -> device gather) must be the same step as
    step(dataset.batch(indices)): same loss and same weights after four updates (eager, capture, replays)."""
    import copy
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
    torch.manual_seed(2)
    N, T_raw, lo, hi, B, H, D = 40, 72, 4, 68, 8, 32, 24
    ds = DeviceEEGDataset.from_tensor(torch.randn(N, C, T_raw), time_low=lo, time_high=hi, apply_norm_with_stds_and_means=True)
    assert ds.samples == hi - lo and len(ds) == N
    sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 2)
    feats = [torch.randn(B, D).pin_memory() for _ in range(2)]
    batches = [torch.tensor([3, 39, 0, 17, -1, 8, 21, 5]), torch.tensor([9, 2, 30, 31, 11, 4, 6, 12])]
    results = []
    for mode in ("dataset", "batch"):
        torch.manual_seed(7)
        model = csn.Model(C, H, 1, D, compute_dtype=torch.float32).cuda()
        loss_mod = csn.DINOLoss(D, 1, 1.5, 0.22, 5, 10).cuda()
        step = csn.DistillTrainStep(model, loss_mod, lr=1e-2, sos=sos)
        losses = []
        for k in range(4):  # eager first call, then graph capture + replays
            idx, f = batches[k % 2], feats[k % 2]
            if mode == "dataset":
                losses.append(float(step.step_from_dataset(ds, idx, f, epoch=0)))
            else:
                eeg, _, _ = ds.batch(idx)
                losses.append(float(step.step(eeg, f.cuda(), epoch=0)))
        results.append((losses, copy.deepcopy({k: v.cpu() for k, v in model.state_dict().items()})))
    np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-6)
    for k in results[0][1]:
        np.testing.assert_allclose(results[0][1][k].numpy(), results[1][1][k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
    with pytest.raises(IndexError):
        step.step_from_dataset(ds, torch.tensor([0, 1, 2, 3, 4, 5, 6, N]), feats[0])
