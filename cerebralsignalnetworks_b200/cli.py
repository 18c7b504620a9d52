"""`python -m cerebralsignalnetworks_b200.cli --batch_size 16 --num_epochs 50`

The reference trainer's CLI surface (LstmDistillFromDinoV2Train.py:150-231: --batch_size, --num_epochs,
--learning_rate, --seed, --log_dir ...) driving the B200 train step.  There is no dataset or DINOv2 checkpoint
offline, so trials and teacher features are synthetic unless --eeg_dataset points at a .pth written by
ConvertToPth.py:170-201: its trials are then served from HBM by dataset.DeviceEEGDataset (crop
[--time_low, --time_high)), and the teacher features come from its `image_features` entry ([n_images, output_size],
one row per entry of "images") when present, else from a seeded random table (one fixed row per image).  Launch under
torchrun for data parallel: one process per GPU, batch sharded, gradients + centre exchanged once per step.
"""
from __future__ import annotations

import argparse
import os
import time


def build_parser():
    p = argparse.ArgumentParser("EEG -> DINOv2 feature distillation (B200)")
    p.add_argument("--learning_rate", type=float, default=0.001, help="Initial learning rate.")
    p.add_argument("--num_epochs", type=int, default=100, help="Number of epochs to run trainer.")
    p.add_argument("--batch_size", type=int, default=16, help="Global batch size (sharded over ranks).")
    p.add_argument("--log_dir", type=str, default="./logs/DinoV2LstmDistill_b200/")
    p.add_argument("--eeg_dataset", type=str, default="")
    p.add_argument("--seed", default=43, type=int, help="Random seed.")
    p.add_argument("--input_size", type=int, default=128)
    p.add_argument("--lstm_size", type=int, default=128)
    p.add_argument("--lstm_layers", type=int, default=1)
    p.add_argument("--output_size", type=int, default=384, help="teacher feature width (384 ViT-S, 768 ViT-B)")
    p.add_argument("--samples", type=int, default=440)
    p.add_argument("--time_low", type=int, default=20, help="first sample kept of a stored trial (EEGDataset time_low)")
    p.add_argument("--time_high", type=int, default=460, help="one past the last sample kept (EEGDataset time_high)")
    p.add_argument("--trials_per_epoch", type=int, default=2048)
    p.add_argument("--warmup_teacher_temp", type=float, default=1.5)
    p.add_argument("--teacher_temp", type=float, default=0.22)
    p.add_argument("--warmup_teacher_temp_epochs", type=int, default=50)
    p.add_argument("--band", type=float, nargs=2, default=[5.0, 95.0], help="band-pass cut-offs in Hz")
    p.add_argument("--fs", type=float, default=1000.0)
    p.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    p.add_argument("--dist_url", default="env://", type=str)
    p.add_argument("--local_rank", default=0, type=int)
    return p


def main(argv=None):
    import numpy as np
    import torch
    import torch.distributed as dist

    from . import DINOLoss, DistillTrainStep, EEGFilters, Model

    FLAGS, _ = build_parser().parse_known_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(FLAGS.seed)
    np.random.seed(FLAGS.seed)
    os.makedirs(FLAGS.log_dir, exist_ok=True)

    nepochs = max(FLAGS.num_epochs, FLAGS.warmup_teacher_temp_epochs + 1)
    dtype = torch.bfloat16 if FLAGS.precision == "bf16" else torch.float32
    model = Model(FLAGS.input_size, FLAGS.lstm_size, FLAGS.lstm_layers, FLAGS.output_size, include_top=False,
                  compute_dtype=dtype).cuda()
    loss = DINOLoss(FLAGS.output_size, 1, FLAGS.warmup_teacher_temp, FLAGS.teacher_temp,
                    FLAGS.warmup_teacher_temp_epochs, nepochs).cuda()
    sos = EEGFilters(FLAGS.fs).sos(FLAGS.band[0], FLAGS.band[1], 4)
    step = DistillTrainStep(model, loss, lr=FLAGS.learning_rate, sos=sos)

    per_rank = FLAGS.batch_size // world
    if per_rank * world != FLAGS.batch_size:
        raise SystemExit("--batch_size must be divisible by the number of ranks")
    gen = torch.Generator(device="cuda").manual_seed(FLAGS.seed + rank)
    ds = feats_table = None
    if FLAGS.eeg_dataset:
        from .dataset import DeviceEEGDataset
        loaded = torch.load(FLAGS.eeg_dataset, map_location="cpu", weights_only=False)
        ds = DeviceEEGDataset(loaded, time_low=FLAGS.time_low, time_high=FLAGS.time_high)
        if ds.C != FLAGS.input_size:
            raise SystemExit("--input_size %d does not match the dataset's %d channels" % (FLAGS.input_size, ds.C))
        n_images = max(len(ds.image_names), int(ds.image_index.max()) + 1)
        if "image_features" in loaded:
            feats_table = torch.as_tensor(loaded["image_features"]).float().reshape(n_images, -1).cuda()
            if feats_table.shape[1] != FLAGS.output_size:
                raise SystemExit("image_features are %d-d, --output_size is %d" % (feats_table.shape[1], FLAGS.output_size))
        else:
            if rank == 0:
                print("no image_features in the dataset file: using a seeded random teacher-feature table", flush=True)
            feats_table = torch.randn(n_images, FLAGS.output_size, device="cuda",
                                      generator=torch.Generator(device="cuda").manual_seed(FLAGS.seed))
        steps = len(ds) // FLAGS.batch_size
        if steps == 0:
            raise SystemExit("--batch_size %d exceeds the dataset (%d trials)" % (FLAGS.batch_size, len(ds)))
    t_axis = torch.arange(FLAGS.samples, device="cuda") / FLAGS.fs
    if ds is None:
        steps = max(1, FLAGS.trials_per_epoch // FLAGS.batch_size)
    for epoch in range(FLAGS.num_epochs):
        # step() hands back the step's STATIC loss tensor when the step is a replayed CUDA graph (every replay overwrites
        # it), so the epoch mean is accumulated on the device, step by step, not stacked from aliases afterwards
        t0, loss_sum, n_steps = time.time(), torch.zeros((), device="cuda"), 0
        if ds is not None:
            order = torch.Generator().manual_seed(FLAGS.seed + epoch)  # same permutation on every rank
            for idx in ds.epoch_batches(per_rank, shuffle=True, generator=order, rank=rank, world=world):
                eeg, _, img = ds.batch(idx)
                loss_sum += step.step(eeg, feats_table[img].contiguous(), epoch)
                n_steps += 1
        for _ in range(steps if ds is None else 0):
            eeg = torch.randn(per_rank, FLAGS.input_size, FLAGS.samples, device="cuda", generator=gen)
            eeg += 0.5 * torch.sin(2 * torch.pi * 40.0 * t_axis)  # utils/PerilsEEGDataset.py:140-147
            feats = torch.randn(per_rank, FLAGS.output_size, device="cuda", generator=gen)
            loss_sum += step.step(eeg, feats, epoch)
            n_steps += 1
        torch.cuda.synchronize()
        if rank == 0:
            mean = float(loss_sum) / max(1, n_steps)
            dt = time.time() - t0
            print(f"EPOCH {epoch} train_loss: {mean:.6f} T: {loss.teacher_temp_schedule[epoch]:.4f} "
                  f"({steps * FLAGS.batch_size / dt:.0f} trials/s)", flush=True)
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(FLAGS.log_dir, "lstm_dinov2_last.pth"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
