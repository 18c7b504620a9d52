"""Two-GPU check of the fused peer all-reduce + Adam + centre-EMA kernel against the ncclAllReduce path: same
seeds, same shards, eager and replayed steps; weights / moments / centre must agree between the two exchanges and be
bit-identical across ranks.  Skipped on a single-GPU box (the driver's 1-GPU test run)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    results = {}
    for mode in ("nccl", "peer"):
        torch.manual_seed(11)
        model = csn.Model(16, 32, 1, 24, include_top=False, compute_dtype=torch.float32).to(dev)
        crit = csn.DINOLoss(24, 1, 1.5, 0.22, 5, 10).to(dev)
        step = csn.DistillTrainStep(model, crit, lr=1e-2, sos=sos, use_cuda_graph=True, dp_exchange=mode)
        assert step.dp_exchange == ("peer_fused_adam" if mode == "peer" else "nccl_allreduce")
        g = torch.Generator(device=dev).manual_seed(100 + rank)  # every rank its own shard
        losses = []
        for i in range(7):  # call 1 eager, call 2 captures, calls 3.. replay; epoch change re-captures
            eeg = torch.randn(4, 16, 48, device=dev, generator=g)
            feats = torch.randn(4, 24, device=dev, generator=g)
            losses.append(float(step.step(eeg, feats, epoch=0 if i < 5 else 1)))
        torch.cuda.synchronize()
        results[mode] = dict(losses=np.array(losses), p=step.flat_p.cpu().numpy(), m=step.exp_avg.cpu().numpy(),
                             v=step.exp_avg_sq.cpu().numpy(), c=crit.center.cpu().numpy().ravel(),
                             steps=int(step._step_dev.item()))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{m}_{k}": v for m, r in results.items() for k, v in r.items()})
    dist.barrier()
    os._exit(0)  # leave without tearing NCCL down under captured graphs (see bench.py)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_fused_adam_matches_nccl_exchange(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), f"rank{k}.npz")) for k in range(world)]
    for k in range(world):
        assert int(r[k]["peer_steps"]) == 7 and int(r[k]["nccl_steps"]) == 7
        # the two exchanges agree (summation order differs: ring vs rank order)
        for name in ("p", "m", "v", "c"):
            np.testing.assert_allclose(r[k][f"peer_{name}"], r[k][f"nccl_{name}"], rtol=2e-5, atol=1e-7, err_msg=name)
        np.testing.assert_allclose(r[k]["peer_losses"], r[k]["nccl_losses"], rtol=1e-5)
    # replicas stay bit-identical under the peer exchange (every rank sums in rank order)
    for name in ("p", "m", "v", "c"):
        assert np.array_equal(r[0][f"peer_{name}"], r[1][f"peer_{name}"]), name
