"""MultiCropDistillStep -- the data-parallel DINO step of LstmDistillation.py:537-626 -- against the CPU oracle: the
restated Model / DINOHead / MultiCropWrapper / multi-crop DINOLoss (oracle/distill.py, pinned by reference-generated
goldens) driven through the reference loop's own sequence (crops -> teacher / student -> loss -> backward -> DDP mean ->
clip_gradients -> cancel_gradients_last_layer -> AdamW with per-iteration lr / wd -> EMA teacher -> centre update).
fp32 compute mode, so the comparison is tight; the world-size-2 case runs two real ranks on two GPUs (two-shot exchange
over peer memory) against the SAME loop with two emulated ranks on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import distill as od

pytestmark = pytest.mark.gpu

C, H, L, K, HID, BOT = 16, 32, 2, 96, 48, 16
T, GLEN, LLEN = 40, 24, 16
N_IT = 4


def _schedules():
    lr = np.linspace(2e-3, 1e-3, N_IT)
    wd = np.linspace(0.04, 0.1, N_IT)
    mom = np.linspace(0.9, 0.95, N_IT)
    return lr, wd, mom


def _crop_starts():
    rng = np.random.RandomState(3)
    return [[int(rng.randint(0, T - GLEN)) for _ in range(2)] + [int(rng.randint(0, T - LLEN)) for _ in range(4)] for _ in range(N_IT)]


def _crops(eeg, starts):
    g = [eeg[:, s:s + GLEN, :] for s in starts[:2]]
    l = [eeg[:, s:s + LLEN, :] for s in starts[2:]]
    return g, l


def _oracle_run(eeg_per_rank, state, world):
    """The reference loop with `world` emulated ranks: identical replicas, per-rank shards, gradients averaged, centre
    statistics summed over ranks (DINOLoss.update_center's all_reduce), then one optimiser / EMA step on every replica."""
    lr, wd, mom = _schedules()
    starts = _crop_starts()

    def make():
        return od.MultiCropWrapper(od.Model(C, H, L, H, include_top=False), od.DINOHead(H, K, hidden_dim=HID, bottleneck_dim=BOT))
    student, teacher = make(), make()
    student.load_state_dict(state)
    teacher.load_state_dict(state)
    for p in teacher.parameters():
        p.requires_grad = False
    reg, noreg = [], []
    for n, p in student.named_parameters():
        if p.requires_grad:
            (noreg if n.endswith(".bias") or p.dim() == 1 else reg).append(p)
    opt = torch.optim.AdamW([{"params": reg}, {"params": noreg, "weight_decay": 0.0}])
    # one loss object per rank (each keeps its own centre; the all-reduce makes them equal)
    pending = []
    crits = [od.DINOLossMultiCrop(K, 6, 0.04, 0.07, 2, 4, world_sum=lambda t: pending.append(t), world_size=world) for _ in range(world)]
    losses = []
    for it in range(N_IT):
        epoch = it // 2
        for i, g in enumerate(opt.param_groups):
            g["lr"] = lr[it]
            if i == 0:
                g["weight_decay"] = wd[it]
        opt.zero_grad()
        rank_losses, sums = [], []
        for r in range(world):
            gv, lv = _crops(eeg_per_rank[r][it], starts[it])
            with torch.no_grad():
                t_out = torch.stack([teacher(v) for v in gv], dim=0)
            s_out = torch.stack([student(v) for v in gv + lv], dim=0)
            # loss WITHOUT the in-place centre update (done below with the cross-rank sum)
            crit = crits[r]
            student_out = (s_out / crit.student_temp).chunk(6)
            q = torch.softmax((t_out - crit.center) / crit.teacher_temp_schedule[epoch], dim=-1).detach()
            total = sum(torch.sum(-q * torch.log_softmax(student_out[v], dim=-1), dim=-1).mean() for v in range(1, 6)) / 5
            (total / world).backward()  # DDP: mean of the ranks' gradients
            rank_losses.append(float(total))
            sums.append(torch.sum(t_out, dim=0, keepdim=True))
        bc = sum(sums) / (2 * world)  # len(teacher_output) = 2 stacked global views
        for crit in crits:
            crit.center = crit.center * crit.center_momentum + bc * (1 - crit.center_momentum)
        for p in student.parameters():  # utils.clip_gradients
            if p.grad is not None:
                coef = 3.0 / (p.grad.norm(2) + 1e-6)
                if coef < 1:
                    p.grad.mul_(coef)
        if epoch < 1:  # cancel_gradients_last_layer
            for n, p in student.named_parameters():
                if "last_layer" in n:
                    p.grad = None
        opt.step()
        with torch.no_grad():
            for q_, k_ in zip(student.parameters(), teacher.parameters()):
                k_.mul_(mom[it]).add_((1 - mom[it]) * q_.detach())
        losses.append(rank_losses)
    return (np.array(losses), {n: p.detach().clone() for n, p in student.named_parameters()},
            {n: p.detach().clone() for n, p in teacher.named_parameters()}, crits[0].center.clone())


def _our_run(eeg_its, state, dev, world):
    import cerebralsignalnetworks_b200 as csn
    lr, wd, mom = _schedules()
    starts = _crop_starts()

    def make():
        return csn.MultiCropWrapper(csn.Model(C, H, L, H, include_top=False, compute_dtype=torch.float32),
                                    csn.DINOHead(H, K, hidden_dim=HID, bottleneck_dim=BOT)).to(dev)
    student, teacher = make(), make()
    student.load_state_dict(state)
    crit = csn.DINOLoss(K, 6, 0.04, 0.07, 2, 4).to(dev)
    step = csn.MultiCropDistillStep(student, teacher, crit, lr, wd, mom, clip_grad=3.0, freeze_last_layer=1,
                                    global_len=GLEN, local_len=LLEN, batch_size=eeg_its[0].shape[0])
    losses = []
    for it in range(N_IT):
        eeg = eeg_its[it].to(dev)
        losses.append(float(step.step(eeg, epoch=it // 2, it=it, crops=_crops(eeg, starts[it]))))
    torch.cuda.synchronize()
    return (np.array(losses), {n: p.detach().cpu() for n, p in student.named_parameters()},
            {n: p.detach().cpu() for n, p in teacher.named_parameters()}, crit.center.detach().cpu())


def _initial_state():
    torch.manual_seed(17)
    m = od.MultiCropWrapper(od.Model(C, H, L, H, include_top=False), od.DINOHead(H, K, hidden_dim=HID, bottleneck_dim=BOT))
    return {k: v.clone() for k, v in m.state_dict().items()}


def _check(ours, ref, rank=0):
    losses, sp, tp, center = ours
    r_losses, r_sp, r_tp, r_center = ref
    np.testing.assert_allclose(losses, r_losses[:, rank], rtol=2e-4)
    for n in r_sp:
        scale = r_sp[n].abs().max().item() + 1e-12
        assert (sp[n] - r_sp[n]).abs().max().item() <= 2e-3 * scale + 2e-5, ("student", n)
    for n in r_tp:
        scale = r_tp[n].abs().max().item() + 1e-12
        assert (tp[n] - r_tp[n]).abs().max().item() <= 2e-3 * scale + 2e-5, ("teacher", n)
    np.testing.assert_allclose(center.numpy().reshape(r_center.shape), r_center.numpy(), rtol=1e-4, atol=1e-6)


def test_multicrop_step_matches_reference_loop_single_rank():
    state = _initial_state()
    g = torch.Generator().manual_seed(5)
    eeg = [torch.randn(3, T, C, generator=g) for _ in range(N_IT)]
    ref = _oracle_run([eeg], state, world=1)
    ours = _our_run(eeg, state, torch.device("cuda", 0), world=1)
    _check(ours, ref)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    state = _initial_state()
    g = torch.Generator().manual_seed(50 + rank)
    eeg = [torch.randn(3, T, C, generator=g) for _ in range(N_IT)]
    losses, sp, tp, center = _our_run(eeg, state, dev, world)
    torch.save({"losses": losses, "sp": sp, "tp": tp, "center": center}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multicrop_step_two_ranks_match_reference_loop(tmp_path):
    """Two real ranks (two-shot exchange of head / backbone gradients and the per-row centre statistics over peer memory)
    equal the reference loop with two emulated ranks; the replicas are bit-identical."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = [torch.load(os.path.join(str(tmp_path), f"rank{k}.pt"), weights_only=False) for k in range(world)]
    state = _initial_state()
    eeg = []
    for r in range(world):
        g = torch.Generator().manual_seed(50 + r)
        eeg.append([torch.randn(3, T, C, generator=g) for _ in range(N_IT)])
    ref = _oracle_run(eeg, state, world)
    for r in range(world):
        _check((out[r]["losses"], out[r]["sp"], out[r]["tp"], out[r]["center"]), ref, rank=r)
    for n in out[0]["sp"]:
        assert torch.equal(out[0]["sp"][n], out[1]["sp"][n]), n
        assert torch.equal(out[0]["tp"][n], out[1]["tp"][n]), n
    assert torch.equal(out[0]["center"], out[1]["center"])


def test_dino_cli_trains_and_checkpoints(tmp_path, capsys):
    """`cli_dino` (the LstmDistillation.py entry point's flags) runs two synthetic epochs and writes log.txt + checkpoint.pth
    with the reference's checkpoint keys (LstmDistillation.py:634-646)."""
    import json
    from cerebralsignalnetworks_b200 import cli_dino
    out = str(tmp_path)
    cli_dino.main(["--batch_size_per_gpu", "4", "--epochs", "2", "--out_dim", "64", "--channels", "16", "--samples", "330",
                   "--lstm_size", "32", "--lstm_layers", "2", "--trials_per_epoch", "12", "--warmup_epochs", "1",
                   "--warmup_teacher_temp_epochs", "1", "--saveckp_freq", "1", "--output_dir", out, "--precision", "fp32"])
    lines = [json.loads(l) for l in open(os.path.join(out, "log.txt"))]
    assert len(lines) == 2 and all(np.isfinite(l["train_loss"]) for l in lines)
    ck = torch.load(os.path.join(out, "checkpoint.pth"), weights_only=False)
    assert {"student", "teacher", "epoch", "args", "dino_loss"} <= set(ck)
    assert any(k.startswith("backbone.lstm.weight_ih_l0") for k in ck["teacher"])
