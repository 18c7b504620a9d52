"""Generate tests/golden/*.npz from the REFERENCE'S OWN classes.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):
    python -m oracle.make_golden
The vectors are small, committed, and travel to the GPU box where /root/reference does not exist.
Library versions are recorded in every file.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy
import torch
import torch.distributed as dist

from .ref_import import import_reference

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
VERS = dict(torch=torch.__version__, scipy=scipy.__version__, numpy=np.__version__)


def _init_pg():
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29591")
        dist.init_process_group("gloo", rank=0, world_size=1)


def golden_dino_single():
    ref = import_reference("LstmDistillFromDinoV2Train")
    g = torch.Generator().manual_seed(43)
    B, K, E = 8, 48, 12
    crit = ref.DINOLoss(out_dim=K, ncrops=4, warmup_teacher_temp=1.5, teacher_temp=0.22,
                        warmup_teacher_temp_epochs=5, nepochs=E)
    out = {"schedule": crit.teacher_temp_schedule.copy()}
    for step, epoch in enumerate([0, 3, 7]):
        s = torch.randn(B, K, generator=g, requires_grad=True)
        t = torch.randn(B, K, generator=g) * 2.0
        c_before = crit.center.clone()
        loss = crit(s, t, epoch)
        loss.backward()
        out.update({f"student{step}": s.detach().numpy(), f"teacher{step}": t.numpy(), f"epoch{step}": np.int64(epoch),
                    f"center_before{step}": c_before.numpy(), f"loss{step}": loss.detach().numpy(),
                    f"grad{step}": s.grad.numpy(), f"center_after{step}": crit.center.numpy()})
    np.savez(os.path.join(GOLD, "dino_loss_single.npz"), versions=str(VERS), **out)


def golden_dino_multicrop():
    ref = import_reference("LstmDistillation")
    g = torch.Generator().manual_seed(44)
    V, B, K, E = 6, 4, 40, 10
    crit = ref.DINOLoss(out_dim=K, ncrops=V, warmup_teacher_temp=0.04, teacher_temp=0.07,
                        warmup_teacher_temp_epochs=4, nepochs=E)
    out = {"schedule": crit.teacher_temp_schedule.copy()}
    for step, epoch in enumerate([0, 2]):
        s = torch.randn(V, B, K, generator=g, requires_grad=True)
        t = torch.randn(2, B, K, generator=g)
        c_before = crit.center.clone()
        loss = crit(s, t, epoch)
        loss.backward()
        out.update({f"student{step}": s.detach().numpy(), f"teacher{step}": t.numpy(), f"epoch{step}": np.int64(epoch),
                    f"center_before{step}": c_before.numpy(), f"loss{step}": loss.detach().numpy(),
                    f"grad{step}": s.grad.numpy(), f"center_after{step}": crit.center.numpy()})
    np.savez(os.path.join(GOLD, "dino_loss_multicrop.npz"), versions=str(VERS), **out)


def golden_dino_head():
    ref = import_reference("LstmDistillation")
    torch.manual_seed(45)
    head = ref.DINOHead(in_dim=16, out_dim=24, nlayers=3, hidden_dim=32, bottleneck_dim=8)
    # std=.02 init makes everything tiny; perturb so the test is sensitive
    with torch.no_grad():
        for p in head.parameters():
            p.add_(torch.randn_like(p) * 0.3)
        head.last_layer.weight_g.fill_(1)
    x = torch.randn(10, 16, requires_grad=True)
    y = head(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy()}
    for n, p in head.named_parameters():
        out["param." + n] = p.detach().numpy()
        if p.grad is not None:
            out["grad." + n] = p.grad.numpy()
    np.savez(os.path.join(GOLD, "dino_head.npz"), versions=str(VERS), **out)


def golden_dino_head_bn():
    """The use_bn=True head of the reference (LstmDistillation.py:72-80) in training mode: output, gradients and the
    running statistics after one forward pass."""
    ref = import_reference("LstmDistillation")
    torch.manual_seed(46)
    head = ref.DINOHead(in_dim=16, out_dim=24, use_bn=True, nlayers=3, hidden_dim=32, bottleneck_dim=8)
    with torch.no_grad():
        for p in head.parameters():
            p.add_(torch.randn_like(p) * 0.3)
        head.last_layer.weight_g.fill_(1)
    out = {}
    for n, b in head.state_dict().items():
        out["state0." + n] = b.detach().clone().numpy()
    x = torch.randn(12, 16, requires_grad=True)
    y = head(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    out.update({"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy()})
    for n, p in head.named_parameters():
        if p.grad is not None:
            out["grad." + n] = p.grad.numpy()
    for n, b in head.named_buffers():
        out["buf1." + n] = b.detach().numpy()
    np.savez(os.path.join(GOLD, "dino_head_bn.npz"), versions=str(VERS), **out)


def golden_utils():
    u = import_reference("utils.utils")
    out = {
        "cos_a": u.cosine_scheduler(0.0005, 1e-6, 10, 7, warmup_epochs=2),
        "cos_b": u.cosine_scheduler(0.04, 0.4, 5, 3),
        "cos_c": u.cosine_scheduler(0.996, 1.0, 4, 5),
    }
    # clip_gradients (utils/utils.py:132-141): per-parameter clipping
    torch.manual_seed(46)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.randn_like(p) * (3.0 if i % 2 == 0 else 0.1)
        out[f"clip_gin{i}"] = p.grad.clone().numpy()
    norms = u.clip_gradients(lin, 1.5)
    out["clip_norms"] = np.asarray(norms)
    for i, p in enumerate(lin.parameters()):
        out[f"clip_gout{i}"] = p.grad.numpy()
    np.savez(os.path.join(GOLD, "utils.npz"), versions=str(VERS), **out)


def golden_filters():
    """scipy outputs (the arithmetic the reference calls) + the reference's own remove_noise."""
    from scipy import signal

    util = import_reference("utils.Utilities")
    rng = np.random.default_rng(47)
    x = rng.normal(size=(3, 5, 440)).astype(np.float32)  # [B, C, T]
    nyq = 500.0
    sos = signal.butter(4, [5.0 / nyq, 95.0 / nyq], btype="bandpass", output="sos")
    out = {"x": x, "sos_5_95": sos,
           "sosfilt_5_95": signal.sosfilt(sos, x.astype(np.float64), axis=-1),
           "sosfiltfilt_5_95": signal.sosfiltfilt(sos, x.astype(np.float64), axis=-1)}
    # reference remove_noise takes [S, T, C]
    eeg_stc = np.ascontiguousarray(x.transpose(0, 2, 1)).astype(np.float64)
    out["remove_noise_1_50"] = util.Utilities().remove_noise(eeg_stc, 1000.0)  # [S, T, C]
    sos150 = signal.butter(4, [1.0 / nyq, 50.0 / nyq], btype="bandpass", output="sos")
    out["sos_1_50"] = sos150
    # EEGFilters attributes (utils/EEGFilters.py:10-16)
    f = import_reference("utils.EEGFilters").EEGFilters(1000.0)
    out["eegfilters_attrs"] = np.array([f.low_cutoff, f.high_cutoff, f.fs, f.low_cutoff_norm, f.high_cutoff_norm])
    np.savez(os.path.join(GOLD, "filters.npz"), versions=str(VERS), **out)


def golden_lstm_step():
    """One full distill step of the CPU oracle (restated Model on torch.nn.LSTM + the REFERENCE
    DINOLoss + Adam): inputs, initial weights, embeddings, loss, grads, post-step weights."""
    from .distill import Model, synthetic_batch
    from .filters import design_bandpass_sos
    from scipy.signal import sosfilt

    ref = import_reference("LstmDistillFromDinoV2Train")
    for tag, (B, C, T, H, L, D, top) in {"l1": (4, 16, 60, 32, 1, 24, False), "l2top": (3, 12, 40, 32, 2, 20, True)}.items():
        torch.manual_seed(48)
        model = Model(C, H, L, D, include_top=top)
        eeg, feats, labels = synthetic_batch(B, C, T, D, seed=43)
        sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
        xf = sosfilt(sos, eeg.astype(np.float64), axis=-1).astype(np.float32)
        crit = ref.DINOLoss(out_dim=D, ncrops=1, warmup_teacher_temp=1.5, teacher_temp=0.22,
                            warmup_teacher_temp_epochs=5, nepochs=10)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        out = {"eeg": eeg, "feats": feats, "filtered": xf}
        for n, p in model.named_parameters():
            out["w0." + n] = p.detach().clone().numpy()
        x = torch.from_numpy(xf).transpose(1, 2).contiguous()
        opt.zero_grad()
        y = model(x)
        emb = y[0] if isinstance(y, tuple) else y
        loss = crit(emb, torch.from_numpy(feats), 1)
        loss.backward()
        opt.step()
        out["emb"] = emb.detach().numpy()
        if isinstance(y, tuple):
            out["cls"] = y[1].detach().numpy()
        out["loss"] = loss.detach().numpy()
        out["center_after"] = crit.center.numpy()
        for n, p in model.named_parameters():
            out["g." + n] = (p.grad.numpy() if p.grad is not None else np.zeros_like(p.detach().numpy()))
            out["w1." + n] = p.detach().numpy()
        np.savez(os.path.join(GOLD, f"distill_step_{tag}.npz"), versions=str(VERS), **out)


def golden_alt_losses():
    """FeatureDistributionLoss (live in LstmDistillFromDinoV2Train.py:371) and CosineSimilarityLoss (:36-43), run
    through the REFERENCE'S OWN classes: inputs, loss, gradients w.r.t. the student features and the class logits."""
    ref = import_reference("LstmDistillFromDinoV2Train")
    g = torch.Generator().manual_seed(49)
    out = {}
    crit = ref.FeatureDistributionLoss(nepochs=12, warmup_teacher_temp=1.5, teacher_temp=0.22, warmup_teacher_temp_epochs=5)
    out["fd_schedule"] = crit.teacher_temp_schedule.copy()
    for i, (B, K, C, epoch) in enumerate([(6, 48, 40, 0), (5, 384, 40, 3), (9, 100, 7, 8)]):
        s = torch.randn(B, K, generator=g, requires_grad=True)
        t = torch.randn(B, K, generator=g) * 1.5
        pred = torch.randn(B, C, generator=g, requires_grad=True)
        label = torch.randint(0, C, (B,), generator=g)
        loss = crit(s, t, epoch, label, pred_label=pred)
        loss.backward()
        out.update({f"fd_student{i}": s.detach().numpy(), f"fd_teacher{i}": t.numpy(), f"fd_pred{i}": pred.detach().numpy(),
                    f"fd_label{i}": label.numpy(), f"fd_epoch{i}": np.int64(epoch), f"fd_loss{i}": loss.detach().numpy(),
                    f"fd_dstudent{i}": s.grad.numpy(), f"fd_dpred{i}": pred.grad.numpy()})
    out["fd_alpha"], out["fd_beta"] = np.float64(ref.HyperParams.alpha), np.float64(ref.HyperParams.beta)
    cos = ref.CosineSimilarityLoss()
    for i, (B, K) in enumerate([(4, 24), (7, 384)]):
        s = torch.randn(B, K, generator=g, requires_grad=True)
        t = torch.randn(B, K, generator=g)
        loss = cos(s, t)
        loss.backward()
        out.update({f"cos_student{i}": s.detach().numpy(), f"cos_teacher{i}": t.numpy(), f"cos_loss{i}": loss.detach().numpy(),
                    f"cos_dstudent{i}": s.grad.numpy()})
    np.savez(os.path.join(GOLD, "alt_losses.npz"), versions=str(VERS), **out)


def golden_kd_losses():
    """The `loss_fn_kd` family of the older training scripts, run through the REFERENCE'S OWN functions:
    LSTMDistillRetreival.py:40-70 (soft targets + smooth-L1), LstmDistillFromDinoV2TrainSpampinato.py:107-121 (Hinton KD +
    hard-label CE), LSTMDistill.py:60-98 (negative cosine)."""
    import warnings
    g = torch.Generator().manual_seed(51)
    out = {}
    ret = import_reference("LSTMDistillRetreival")
    out["ret_T"], out["ret_w_soft"], out["ret_w_ce"] = (np.float64(ret.Parameters.temperature),
                                                         np.float64(ret.Parameters.soft_target_loss_weight),
                                                         np.float64(ret.Parameters.ce_loss_weight))
    for i, (B, K, spread) in enumerate([(5, 40, 1.0), (7, 384, 0.6), (3, 768, 2.5)]):
        s = (torch.randn(B, K, generator=g) * spread).requires_grad_(True)
        t = torch.randn(B, K, generator=g) * spread
        loss = ret.loss_fn_kd(s, None, t, ret.Parameters)
        loss.backward()
        out.update({f"ret_student{i}": s.detach().numpy(), f"ret_teacher{i}": t.numpy(), f"ret_loss{i}": loss.detach().numpy(),
                    f"ret_dstudent{i}": s.grad.numpy()})
    spa = import_reference("LstmDistillFromDinoV2TrainSpampinato")

    class P:
        pass
    for i, (B, K, alpha, T) in enumerate([(6, 40, 0.9, 4.0), (4, 100, 0.5, 2.0), (5, 40, 1.0, 3.0)]):
        P.alpha, P.temperature = alpha, T
        s = torch.randn(B, K, generator=g, requires_grad=True)
        t = torch.randn(B, K, generator=g) * 2.0
        y = torch.randint(0, K, (B,), generator=g)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # KLDivLoss warns that reduction='mean' is not the batch mean
            loss = spa.loss_fn_kd(s, y, t, P)
        loss.backward()
        out.update({f"hin_student{i}": s.detach().numpy(), f"hin_teacher{i}": t.numpy(), f"hin_label{i}": y.numpy(),
                    f"hin_alpha{i}": np.float64(alpha), f"hin_T{i}": np.float64(T), f"hin_loss{i}": loss.detach().numpy(),
                    f"hin_dstudent{i}": s.grad.numpy()})
    dis = import_reference("LSTMDistill")
    for i, (B, K) in enumerate([(4, 24), (6, 384)]):
        s = torch.randn(B, K, generator=g, requires_grad=True)
        t = torch.randn(B, K, generator=g)
        loss = dis.loss_fn_kd(s, t, None, None, dis.Parameters)
        loss.backward()
        out.update({f"neg_student{i}": s.detach().numpy(), f"neg_teacher{i}": t.numpy(), f"neg_loss{i}": loss.detach().numpy(),
                    f"neg_dstudent{i}": s.grad.numpy()})
    np.savez(os.path.join(GOLD, "kd_losses.npz"), versions=str(VERS), **out)


def golden_model_analogues():
    """`models.lstm.Model` is absent from the reference, but two scripts ship the class it was derived from:
    LSTMDistillRetreival.LSTMModel (:85-110: nn.LSTM -> last timestep -> fc) and LSTMDistill.LSTMModel (:112-142: fc over
    all timesteps, class head on the features, ReLU on the returned features).  Both start with `x.view(B, C, T)`; on a
    SQUARE input (T == C) that view is the identity, so there their forward is exactly what oracle.distill.Model restates.
    Inputs, weights, outputs and weight gradients of the reference's own classes on square inputs."""
    import contextlib
    import io
    import warnings
    out = {}
    torch.manual_seed(53)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # Variable(volatile=...) of the 0.3-era code
        ret = import_reference("LSTMDistillRetreival")
        m = ret.LSTMModel(input_size=24, hidden_size=16, n_layers=2, out_features=12)
        x = torch.randn(3, 24, 24)
        y = m(x)
        y.pow(2).sum().backward()
        out.update({"ret_x": x.numpy(), "ret_y": y.detach().numpy(), "ret_hidden": np.int64(16), "ret_layers": np.int64(2)})
        out.update({f"ret_w_{k}": v.detach().numpy() for k, v in m.state_dict().items()})
        out.update({f"ret_g_{k}": p.grad.numpy() for k, p in m.named_parameters()})
        dis = import_reference("LSTMDistill")
        m2 = dis.LSTMModel(input_size=20, hidden_size=16, n_layers=1, out_features=10, number_of_classes=7)
        x2 = torch.randn(4, 20, 20)
        with contextlib.redirect_stdout(io.StringIO()):   # the forward prints tensor sizes
            f2, c2 = m2(x2)
        out.update({"dis_x": x2.numpy(), "dis_feat_last": f2[:, -1, :].detach().numpy(), "dis_cls_last": c2[:, -1, :].detach().numpy(),
                    "dis_hidden": np.int64(16), "dis_layers": np.int64(1)})
        out.update({f"dis_w_{k}": v.detach().numpy() for k, v in m2.state_dict().items()})
    np.savez(os.path.join(GOLD, "model_analogues.npz"), versions=str(VERS), **out)


class _FlatL2Stub:
    """What the reference's evaluate() needs from faiss.IndexFlatL2 (utils/Utilities.py:45-58), served by the float64
    exhaustive search of oracle/retrieval.py -- faiss itself is absent from the image."""

    def __init__(self, d):
        self.d, self.is_trained, self.ntotal = d, True, 0
        self._g = np.zeros((0, d), np.float32)

    def add(self, x):
        self._g = np.concatenate([self._g, np.asarray(x, dtype=np.float32).reshape(-1, self.d)])
        self.ntotal = len(self._g)

    def search(self, q, k):
        from .retrieval import flat_search
        D, I = flat_search(self._g, np.asarray(q, dtype=np.float32), k)
        return D.astype(np.float32), I


def run_reference_evaluate(gallery, query, gallery_ids, query_ids, k, n_classes):
    """The reference's OWN Utilities.evaluate (recall / precision bookkeeping, :60-160) on plain arrays; only the faiss
    index is a stand-in.  Returns (Recall_Total, Precision_Total)."""
    import contextlib
    import io
    import types
    saved = sys.modules.get("faiss")
    stub = types.ModuleType("faiss")
    stub.IndexFlatL2 = _FlatL2Stub
    sys.modules["faiss"] = stub
    try:
        util = import_reference("utils.Utilities")
        util.faiss = stub
        names = {i: f"class{i}" for i in range(n_classes)}
        ds = types.SimpleNamespace(class_id_to_str=names, class_str_to_id={v: kk for kk, v in names.items()})
        lab = lambda ids: [{"ClassId": int(i), "ClassName": names[int(i)]} for i in ids]
        with contextlib.redirect_stdout(io.StringIO()):
            r, p = util.evaluate(types.SimpleNamespace(topK=k), gallery, query, lab(gallery_ids), lab(query_ids), ds)
        return float(r), float(p)
    finally:
        if saved is not None:
            sys.modules["faiss"] = saved
        else:
            sys.modules.pop("faiss", None)


def golden_retrieval():
    """Recall / precision of the reference's own evaluate() (utils/Utilities.py:28-169) on two synthetic galleries."""
    from .retrieval import flat_search
    rng = np.random.default_rng(52)
    out = {}
    for i, (nb, nq, d, k, ncls) in enumerate([(120, 40, 16, 5, 6), (300, 75, 32, 3, 11)]):
        centres = rng.standard_normal((ncls, d)).astype(np.float32) * (0.35, 0.25)[i]
        g_ids, q_ids = rng.integers(0, ncls, nb), rng.integers(0, ncls, nq)
        g = (centres[g_ids] + rng.standard_normal((nb, d))).astype(np.float32)
        q = (centres[q_ids] + rng.standard_normal((nq, d))).astype(np.float32)
        r, p = run_reference_evaluate(g, q, g_ids, q_ids, k, ncls)
        _, I = flat_search(g, q, k)
        out.update({f"gallery{i}": g, f"query{i}": q, f"gallery_ids{i}": g_ids, f"query_ids{i}": q_ids, f"k{i}": np.int64(k),
                    f"I{i}": I, f"recall{i}": np.float64(r), f"precision{i}": np.float64(p)})
    np.savez(os.path.join(GOLD, "retrieval_scores.npz"), versions=str(VERS), **out)


def golden_dataset():
    """The reference's own EEGDataset (utils/PerilsEEGDataset.py) on a small synthetic .pth file in the schema of
    ConvertToPth.py:170-201: dataset-level mean / std and the EEG tensor of every item, with and without normalisation."""
    import tempfile
    from PIL import Image
    mod = import_reference("utils.PerilsEEGDataset")
    rng = np.random.default_rng(50)
    N, C, T_raw, lo, hi = 7, 6, 50, 5, 41
    classes = ["n01440764", "n01443537", "n01484850"]
    images = [f"{classes[i % 3]}_{100 + i}" for i in range(N)]
    data = {"dataset": [], "labels": classes, "images": images, "means": [], "stddevs": []}
    for i in range(N):
        data["dataset"].append({"eeg": torch.from_numpy(rng.normal(3.0, 2.0, size=(C, T_raw))), "image": i,
                                "label": i % 3, "subject": 1})
    out = {"eeg_raw": np.stack([d["eeg"].numpy() for d in data["dataset"]]), "labels": np.array([d["label"] for d in data["dataset"]]),
           "time_low": np.int64(lo), "time_high": np.int64(hi)}
    with tempfile.TemporaryDirectory() as tmp:
        pth = os.path.join(tmp, "spampinato-1-synthetic.pth")
        torch.save(data, pth)
        root = os.path.join(tmp, "images")
        with open(os.path.join(tmp, "labels.txt"), "w"):
            pass
        os.makedirs(root)
        with open(os.path.join(root, "labels.txt"), "w") as f:
            for k, c in enumerate(classes):
                f.write(f"{c} {k + 1} class{k}\n")
        for name in images:
            os.makedirs(os.path.join(root, name.split("_")[0]), exist_ok=True)
            Image.new("RGB", (8, 8), (10, 20, 30)).save(os.path.join(root, name.split("_")[0], name + ".JPEG"))
        for tag, norm in (("plain", False), ("norm", True)):
            ds = mod.EEGDataset(pth, None, subset="train", time_low=lo, time_high=hi, imagesRoot=root,
                                apply_norm_with_stds_and_means=norm, inference_mode=False)
            ds.isDataTransformed = False  # the constructor marks the data "transformed"; __getitem__ crops only when it is not
            out[f"mean_{tag}"], out[f"std_{tag}"] = np.float64(float(ds.mean)), np.float64(float(ds.std))
            items = [ds[i] for i in range(len(ds))]
            out[f"items_{tag}"] = np.stack([it[0].numpy() for it in items])   # [N, T, C]
            out[f"labels_{tag}"] = np.array([int(it[1]) for it in items])
        # the `filter_channels` branch of __getitem__ (:554-565), with and without the per-channel z-score.  The flags are
        # set AFTER construction: passing apply_channel_wise_norm to the constructor runs transformEEGDataToChannelWiseNorm
        # (:463-510), a different (label-wise) transform that rewrites the stored trials.
        sel = [4, 0, 2]
        out["filter_channels"] = np.array(sel)
        for tag, chn in (("chsel", False), ("chsel_norm", True)):
            ds = mod.EEGDataset(pth, None, subset="train", time_low=lo, time_high=hi, imagesRoot=root,
                                apply_norm_with_stds_and_means=False, inference_mode=False)
            ds.isDataTransformed = False
            ds.filter_channels = sel
            ds.apply_channel_wise_norm = chn
            out[f"items_{tag}"] = np.stack([ds[i][0].numpy() for i in range(len(ds))])   # [N, len(sel), T]
    np.savez(os.path.join(GOLD, "dataset.npz"), versions=str(VERS), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    _init_pg()
    if "--only-dataset" in sys.argv:
        golden_dataset()
        print("dataset.npz written")
        return 0
    if "--only-model" in sys.argv:
        golden_model_analogues()
        print("model_analogues.npz written")
        return 0
    if "--only-retrieval" in sys.argv:
        golden_retrieval()
        print("retrieval_scores.npz written")
        return 0
    if "--only-kd-losses" in sys.argv:
        golden_kd_losses()
        print("kd_losses.npz written")
        return 0
    if "--only-head-bn" in sys.argv:
        golden_dino_head_bn()
        print("dino_head_bn.npz written")
        return 0
    if "--only-alt-losses" in sys.argv:
        golden_alt_losses()
        print("alt_losses.npz written")
        return 0
    golden_dino_single()
    golden_dino_multicrop()
    golden_dino_head()
    golden_dino_head_bn()
    golden_utils()
    golden_filters()
    golden_lstm_step()
    golden_alt_losses()
    golden_kd_losses()
    golden_retrieval()
    golden_model_analogues()
    golden_dataset()
    print("golden vectors written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  ", f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    sys.exit(main())
