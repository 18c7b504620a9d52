"""Data-parallel plumbing of the train step: batch sharding and the single flat all-reduce.

Reference semantics (LstmDistillation.py:406-414,445): DistributedSampler(drop_last=True) shards the batch,
DDP averages gradients, DINOLoss.update_center all-reduces the teacher column sums and divides by
len(batch) * world (LstmDistillation.py:154-156).  Here both reductions ride in ONE buffer
[gradients | centre sums]; the division by `world` happens inside the fused Adam / centre-EMA kernels."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_range(n_global: int, rank: int, world: int):
    """Even contiguous shard of a global batch; the remainder is dropped (drop_last=True)."""
    per = n_global // world
    return rank * per, (rank + 1) * per


def allreduce_flat_(flat: torch.Tensor) -> torch.Tensor:
    """SUM over ranks, in place (NCCL on GPUs, gloo in the CPU tests)."""
    if world_size() > 1:
        dist.all_reduce(flat)
    return flat


def split_flat(flat: torch.Tensor, n_param: int, world: int, local_batch: int):
    """-> (mean gradients, batch centre) with the reference's normalisation."""
    return flat[:n_param] / world, flat[n_param:] / (local_batch * world)


class PeerExchange:
    """Peer-mapped flat buffer [gradients | centre sums] + flag block for the fused all-reduce/Adam kernel
    (csn_dp_adam_step_peer).  torch's symmetric-memory allocator does the plumbing (allocation + exchange of the
    mappings over the process group's store); the arithmetic and the cross-rank ordering are ours."""

    def __init__(self, n_floats: int, device: torch.device):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        group = dist.group.WORLD
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 8:
            raise RuntimeError("peer exchange supports up to 8 ranks (one NVSwitch domain)")
        self.buf = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self._h_buf = symm_mem.rendezvous(self.buf, group)
        self._h_flags = symm_mem.rendezvous(self.flags, group)
        self.buf.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(device)
        dist.barrier()  # nobody signals before every flag block is zero
        gp, fp = list(self._h_buf.buffer_ptrs), list(self._h_flags.buffer_ptrs)
        if len(gp) != self.world or len(fp) != self.world or gp[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected mapping")
        self.grad_ptrs = (ctypes.c_void_p * self.world)(*gp)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*fp)
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)


class LocalExchange:
    """World size 1: the same fused kernel (sum over one rank + Adam + centre EMA + step counter in ONE launch instead
    of three), on ordinary device memory."""

    def __init__(self, n_floats: int, device: torch.device):
        import ctypes

        self.world, self.rank = 1, 0
        self.buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.flags = torch.zeros(64, dtype=torch.int32, device=device)
        self.grad_ptrs = (ctypes.c_void_p * 1)(self.buf.data_ptr())
        self.flag_ptrs = (ctypes.c_void_p * 1)(self.flags.data_ptr())
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)


class TwoShotExchange:
    """Peer-mapped flat fp32 buffer + flag block for csn_dp_allreduce_twoshot (the large exchanges of the
    LstmDistillation step: 22.3 M gradients + per-row centre statistics at cfg3).  Up to two exchanges (flag sets) may be in
    flight at once on different streams.  World size 1: an ordinary device buffer, the exchange is a no-op."""

    def __init__(self, n_floats: int, device: torch.device):
        import ctypes

        self.world = world_size()
        self.rank = dist.get_rank() if self.world > 1 else 0
        n_floats = (n_floats + 3) // 4 * 4
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem

            if self.world > 8:
                raise RuntimeError("peer exchange supports up to 8 ranks (one NVSwitch domain)")
            group = dist.group.WORLD
            self.buf = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
            self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
            self._h_buf = symm_mem.rendezvous(self.buf, group)
            self._h_flags = symm_mem.rendezvous(self.flags, group)
            self.buf.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(device)
            dist.barrier()  # nobody signals before every flag block is zero
            bp, fp = list(self._h_buf.buffer_ptrs), list(self._h_flags.buffer_ptrs)
            if len(bp) != self.world or bp[self.rank] != self.buf.data_ptr():
                raise RuntimeError("symmetric-memory rendezvous returned an unexpected mapping")
        else:
            self.buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
            self.flags = torch.zeros(64, dtype=torch.int32, device=device)
            bp, fp = [self.buf.data_ptr()], [self.flags.data_ptr()]
        self.buf_ptrs = (ctypes.c_void_p * self.world)(*bp)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*fp)
        self.epochs = torch.zeros(2, dtype=torch.int32, device=device)   # one exchange counter per flag set
        self.tickets = torch.zeros(4, dtype=torch.int32, device=device)  # two CTA-arrival words per flag set

    def all_reduce_(self, offset: int, n: int, flag_set: int = 0):
        """SUM over ranks of buf[offset : offset + n], in place, on the current stream."""
        import ctypes

        from . import _lib
        if self.world == 1 or n == 0:
            return
        vp = ctypes.c_void_p
        _lib.call("csn_dp_allreduce_twoshot", self.buf_ptrs, self.flag_ptrs, self.world, self.rank, int(offset), int(n),
                  vp(self.epochs.data_ptr() + 4 * flag_set), int(flag_set), vp(self.tickets.data_ptr() + 8 * flag_set),
                  vp(torch.cuda.current_stream().cuda_stream))

    def wait_done(self, flag_set: int = 0):
        """Every peer has gathered the last exchange of this flag set: the buffer range may be overwritten (stream op)."""
        import ctypes

        from . import _lib
        if self.world == 1:
            return
        vp = ctypes.c_void_p
        _lib.call("csn_dp_wait_done", vp(self.flags.data_ptr()), self.world, vp(self.epochs.data_ptr() + 4 * flag_set),
                  int(flag_set), vp(torch.cuda.current_stream().cuda_stream))
