"""`python -m cerebralsignalnetworks_b200.cli_dino --batch_size_per_gpu 64 --epochs 100 --out_dim 65536`

The CLI surface of the reference's data-parallel DINO trainer (LstmDistillation.py:187-348; README.md:19) driving
MultiCropDistillStep: --out_dim, --norm_last_layer, --momentum_teacher, --warmup_teacher_temp, --teacher_temp,
--warmup_teacher_temp_epochs, --weight_decay, --weight_decay_end, --clip_grad, --batch_size_per_gpu, --epochs,
--freeze_last_layer, --lr, --warmup_epochs, --min_lr, --local_crops_number, --seed, --output_dir, --saveckp_freq, with
the reference's defaults.  Launch under torchrun for data parallel (one process per GPU); the learning rate is scaled by
the global batch / 256 as at LstmDistillation.py:483-485.  Trials are synthetic ([N, 495, 96], the EEGDataset window of
:381-387) unless --eeg_dataset points at a .pth written by ConvertToPth.py:170-201.
"""
from __future__ import annotations

import argparse
import json
import os
import time


def bool_flag(s):  # utils/utils.py:201-212
    if s.lower() in {"off", "false", "0"}:
        return False
    if s.lower() in {"on", "true", "1"}:
        return True
    raise argparse.ArgumentTypeError("invalid value for a boolean flag")


def build_parser():
    p = argparse.ArgumentParser("EEG DINO self-distillation (B200)")
    p.add_argument("--out_dim", default=384, type=int)
    p.add_argument("--norm_last_layer", default=True, type=bool_flag)
    p.add_argument("--momentum_teacher", default=0.996, type=float)
    p.add_argument("--use_bn_in_head", default=False, type=bool_flag)
    p.add_argument("--warmup_teacher_temp", default=0.04, type=float)
    p.add_argument("--teacher_temp", default=0.04, type=float)
    p.add_argument("--warmup_teacher_temp_epochs", default=30, type=int)
    p.add_argument("--use_fp16", type=bool_flag, default=True, help="accepted for compatibility: the B200 path runs bf16 "
                   "operands with fp32 accumulation and master weights, no GradScaler")
    p.add_argument("--weight_decay", type=float, default=0.04)
    p.add_argument("--weight_decay_end", type=float, default=0.4)
    p.add_argument("--clip_grad", type=float, default=3.0)
    p.add_argument("--batch_size_per_gpu", default=8, type=int)
    p.add_argument("--epochs", default=200, type=int)
    p.add_argument("--freeze_last_layer", default=1, type=int)
    p.add_argument("--lr", default=0.0005, type=float)
    p.add_argument("--warmup_epochs", default=10, type=int)
    p.add_argument("--min_lr", type=float, default=1e-06)
    p.add_argument("--optimizer", default="adamw", type=str, choices=["adamw"])
    p.add_argument("--local_crops_number", type=int, default=4)
    p.add_argument("--eeg_dataset", type=str, default="")
    p.add_argument("--seed", default=43, type=int)
    p.add_argument("--output_dir", default="./output", type=str)
    p.add_argument("--saveckp_freq", default=10, type=int)
    p.add_argument("--num_workers", default=0, type=int)
    p.add_argument("--dist_url", default="env://", type=str)
    p.add_argument("--local_rank", default=0, type=int)
    # shapes of the synthetic run (the reference hard-codes them: Model(96, 128, 4, 128), time window 0..495)
    p.add_argument("--channels", default=96, type=int)
    p.add_argument("--samples", default=495, type=int)
    p.add_argument("--lstm_size", default=128, type=int)
    p.add_argument("--lstm_layers", default=4, type=int)
    p.add_argument("--trials_per_epoch", default=2048, type=int, help="synthetic trials per epoch (global)")
    p.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    return p


def main(argv=None):
    import numpy as np
    import torch
    import torch.distributed as dist

    from . import DINOHead, DINOLoss, Model, MultiCropDistillStep, MultiCropWrapper
    from .schedules import cosine_scheduler

    args, _ = build_parser().parse_known_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)  # same initial weights on every rank (DDP's broadcast)
    np.random.seed(args.seed)     # same crop starts on every rank
    os.makedirs(args.output_dir, exist_ok=True)
    dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32

    def make():
        return MultiCropWrapper(Model(args.channels, args.lstm_size, args.lstm_layers, args.lstm_size, include_top=False,
                                      compute_dtype=dtype),
                                DINOHead(args.lstm_size, args.out_dim, use_bn=args.use_bn_in_head,
                                         norm_last_layer=args.norm_last_layer, compute_dtype=dtype)).to(dev)
    student, teacher = make(), make()
    B = args.batch_size_per_gpu
    ds = None
    if args.eeg_dataset:
        from .dataset import DeviceEEGDataset
        ds = DeviceEEGDataset(torch.load(args.eeg_dataset, map_location="cpu", weights_only=False), time_low=0,
                              time_high=args.samples)
        n_trials = len(ds)
    else:
        n_trials = args.trials_per_epoch
    iters = max(1, n_trials // (B * world))
    nepochs = max(args.epochs, args.warmup_teacher_temp_epochs + 1)
    loss = DINOLoss(args.out_dim, args.local_crops_number + 2, args.warmup_teacher_temp, args.teacher_temp,
                    args.warmup_teacher_temp_epochs, nepochs).to(dev)
    lr_sched = cosine_scheduler(args.lr * (B * world) / 256.0, args.min_lr, args.epochs, iters, warmup_epochs=args.warmup_epochs)
    wd_sched = cosine_scheduler(args.weight_decay, args.weight_decay_end, args.epochs, iters)
    mom_sched = cosine_scheduler(args.momentum_teacher, 1.0, args.epochs, iters)
    step = MultiCropDistillStep(student, teacher, loss, lr_sched, wd_sched, mom_sched, clip_grad=args.clip_grad,
                                freeze_last_layer=args.freeze_last_layer, n_local=args.local_crops_number, batch_size=B)
    gen = torch.Generator(device=dev).manual_seed(args.seed + 1 + rank)
    for epoch in range(args.epochs):
        t0, loss_sum = time.time(), torch.zeros((), device=dev)
        batches = None
        if ds is not None:
            order = torch.Generator().manual_seed(args.seed)  # the reference never calls sampler.set_epoch: same order every epoch
            batches = list(ds.epoch_batches(B, shuffle=True, generator=order, rank=rank, world=world))[:iters]
        for i in range(iters):
            if batches is not None:
                eeg = ds.batch_btc(batches[i])[0]
            else:
                eeg = torch.randn(B, args.samples, args.channels, device=dev, generator=gen)
            loss_sum += step.step(eeg, epoch=epoch, it=epoch * iters + i)
        torch.cuda.synchronize()
        if rank == 0:
            dt = time.time() - t0
            stats = {"epoch": epoch, "train_loss": float(loss_sum) / iters, "train_lr": float(lr_sched[epoch * iters]),
                     "train_wd": float(wd_sched[epoch * iters]), "trials_per_s": iters * B * world / dt}
            print(json.dumps(stats), flush=True)
            with open(os.path.join(args.output_dir, "log.txt"), "a") as f:  # LstmDistillation.py:647-651
                f.write(json.dumps(stats) + "\n")
            if (epoch + 1) % args.saveckp_freq == 0 or epoch + 1 == args.epochs:
                torch.save({"student": student.state_dict(), "teacher": teacher.state_dict(), "epoch": epoch + 1,
                            "args": vars(args), "dino_loss": loss.state_dict()},
                           os.path.join(args.output_dir, "checkpoint.pth"))
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
