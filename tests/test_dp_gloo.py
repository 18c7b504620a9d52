"""world_size-2 gloo run of the data-parallel combine (the reference's own backend is gloo, utils/utils.py:491-497):
per-rank batch shards -> flat [grads | centre sums] -> one all-reduce -> must equal the single-process full batch."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cerebralsignalnetworks_b200 import dp
from oracle import distill as od


def _flat_for_shard(eeg, feats, seed):
    torch.manual_seed(seed)
    model = od.Model(8, 12, 1, 10, include_top=False)
    x = torch.from_numpy(eeg).transpose(1, 2).contiguous()
    t = torch.from_numpy(feats)
    import torch.nn.functional as F
    q = F.softmax((t - torch.zeros(1, 10)) / 0.86, dim=-1)
    loss = torch.sum(-q * F.log_softmax(model(x) / 0.1, dim=-1), dim=-1).mean()
    loss.backward()
    grads = torch.cat([p.grad.flatten() for p in model.parameters()])
    return torch.cat([grads, t.sum(0)]), grads.numel()


def _worker(rank, world, port, eeg, feats, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    a, b = dp.shard_range(eeg.shape[0], rank, world)
    flat, n_param = _flat_for_shard(eeg[a:b], feats[a:b], seed=7)
    dp.allreduce_flat_(flat)
    g, c = dp.split_flat(flat, n_param, world, b - a)
    if rank == 0:
        out["grads"], out["center"] = g.numpy().copy(), c.numpy().copy()
    dist.destroy_process_group()


def test_two_rank_combine_equals_full_batch():
    eeg, feats, _ = od.synthetic_batch(8, 8, 20, 10, seed=3)
    full, n_param = _flat_for_shard(eeg, feats, seed=7)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29631, eeg, feats, out), nprocs=2, join=True)
    np.testing.assert_allclose(out["grads"], full[:n_param].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out["center"], (full[n_param:] / 8).numpy(), rtol=1e-5, atol=1e-6)


def test_shard_range_drop_last():
    assert dp.shard_range(10, 0, 4) == (0, 2) and dp.shard_range(10, 3, 4) == (6, 8)
    assert dp.shard_range(1024, 7, 8) == (896, 1024)
    assert dp.world_size() == 1
