"""The distillation train step as one fused, data-parallel driver.

Composition (BASELINE.json north_star; step order of LstmDistillFromDinoV2Train.py:358-375):
    band-pass (fused transpose + cast to the encoder's [T,B,C] layout) -> LSTM encoder -> projection
    -> DINO cross-entropy fwd+bwd (+ centre statistics) -> projection bwd -> BPTT
    -> ONE all-reduce over the flat [gradients | centre statistics] buffer (NCCL when world > 1)
    -> fused Adam over the flat parameter buffer -> centre EMA.
Every arithmetic stage is a libcsn_b200 kernel; torch provides memory, streams, torch.distributed and
(optionally) CUDA-graph capture of the whole sequence so the ~30 launches cost one submission.
"""
from __future__ import annotations

import numpy as np
import os

import torch
import torch.distributed as dist

from . import _lib, dp, ops
from ._lib import ACT_NONE, ACT_RELU, DINO_SINGLE
from .functional import encoder_bwd, encoder_fwd, linear_bwd, linear_fwd

_IDENTITY_SOS = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])


class DistillTrainStep:
    def __init__(self, model, loss, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 sos=None, zero_phase=False, use_cuda_graph=True, dp_exchange="auto"):
        _lib.require_gpu()
        self.model, self.loss = model, loss
        self.lr, self.betas, self.eps = lr, betas, eps
        self.weight_decay, self.decoupled = weight_decay, decoupled
        self.sos = _IDENTITY_SOS if sos is None else np.asarray(sos, dtype=np.float64)
        self.zero_phase = zero_phase
        self.step_count = 0
        self.world = dp.world_size()
        dev = next(model.parameters()).device
        self.device = dev

        # ---- flat parameter / gradient / moment buffers; module parameters become views ----
        # Parameters the DINO loss cannot reach (the class head of include_top models) get no gradient in the reference
        # (p.grad is None: torch's optimiser SKIPS them -- no moments, no weight decay).  They sit behind the optimised
        # range of the flat buffer, so the fused Adam never touches them.
        unreached = {id(p) for p in model.classifier.parameters()} if getattr(model, "include_top", False) else set()
        params = [p for p in model.parameters() if id(p) not in unreached]
        frozen = [p for p in model.parameters() if id(p) in unreached]
        offs, n_all = [], 0
        for i, p in enumerate(params + frozen):
            if i == len(params):
                total = n_all  # end of the optimised range
            offs.append(n_all)
            n_all += (p.numel() + 3) // 4 * 4
        if not frozen:
            total = n_all
        self.n_param = total
        K = loss.center.shape[-1]
        self.flat_p = torch.zeros(n_all, dtype=torch.float32, device=dev)
        # [grads | sum_b teacher].  Data parallel: the buffer is peer-mapped so that every rank's fused Adam reads all
        # ranks' gradients straight over NVLink (dp_exchange "peer"); "nccl" keeps one ncclAllReduce + local kernels;
        # "auto" tries peer and falls back to nccl when symmetric memory cannot be set up (no P2P / gloo group).
        self.peer = None
        if self.world > 1 and dev.type == "cuda" and dp_exchange in ("auto", "peer"):
            try:
                self.peer = dp.PeerExchange(total + K, dev)
            except Exception as exc:  # noqa: BLE001 - report and fall back to the NCCL exchange
                if dp_exchange == "peer":
                    raise
                import warnings
                warnings.warn("peer gradient exchange unavailable (%s: %s); using ncclAllReduce" % (type(exc).__name__, exc))
        self.dp_exchange = "peer_fused_adam" if self.peer is not None else ("nccl_allreduce" if self.world > 1 else "single")
        if self.world == 1 and dev.type == "cuda":
            self.peer = dp.LocalExchange(total + K, dev)  # same fused optimiser kernel, one rank
        self.flat_g = self.peer.buf if self.peer is not None else torch.zeros(total + K, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._grad_views = {}
        for p, o in zip(params + frozen, offs):
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            if id(p) in unreached:
                self._grad_views[id(p)] = torch.zeros_like(view)  # never written: the loss does not reach it
            else:
                self._grad_views[id(p)] = self.flat_g[o:o + p.numel()].view(p.shape)
        self.flat_p_opt = self.flat_p[:total]
        self.batch_center = self.flat_g[total:]
        self.center = loss.center.reshape(-1)
        if self.center.numel() != K:
            raise _lib.CsnError("DistillTrainStep drives the single-view loss (centre [1, K])")

        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_tau = None
        self._static = None
        self._idx_ring = None      # step_from_dataset: pinned index slots + their device copy
        self._side = None          # side stream of the fused head's weight-gradient GEMM
        self.fused_head = os.environ.get("CSN_NO_FUSED_HEAD", "0") != "1"
        self._ds_graphs = {}       # (dataset, tau, input slot) -> (graph, loss) of the resident-dataset step
        self._resident = {}        # (eeg ptr, teacher ptr) -> (eeg, teacher): buffers with their own captured graph
        self._resident_graphs = {}  # (eeg ptr, teacher ptr, tau) -> (graph, loss tensor)
        self._pool = None
        self._warm = False
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # device-side Adam step count
        self._adam_consts = torch.zeros(2, dtype=torch.float32, device=dev)
        self._loss_out = torch.zeros((), dtype=torch.float32, device=dev)
        self._stage_events = None

    def grad_of(self, p):
        return self._grad_views[id(p)]

    # -- optional per-stage CUDA-event timing on the launching stream (bench.py's live roofline numbers) --
    def enable_stage_timing(self, on=True):
        self._stage_events = {} if on else None

    class _Stage:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            ev = self.owner._stage_events
            if ev is not None:
                self.start = torch.cuda.Event(enable_timing=True)
                self.stop = torch.cuda.Event(enable_timing=True)
                self.start.record()

        def __exit__(self, *exc):
            ev = self.owner._stage_events
            if ev is not None:
                self.stop.record()
                ev.setdefault(self.name, []).append((self.start, self.stop))

    def _stage(self, name):
        return DistillTrainStep._Stage(self, name)

    def stage_times_ms(self):
        """Mean device time per stage over the recorded steps (call after torch.cuda.synchronize())."""
        out = {}
        for name, pairs in (self._stage_events or {}).items():
            out[name] = sum(a.elapsed_time(b) for a, b in pairs) / max(1, len(pairs))
        return out

    # ------------------------------------------------------------------------------------------------
    def _run(self, eeg_bct, teacher, tau_t, gather=None):
        """`gather` = (dataset, idx_dev): the batch is trials idx_dev of a resident dataset, fetched by the filter itself."""
        m = self.model
        cd = m.compute_dtype
        B = teacher.shape[0]
        with self._stage("filter"):
            if gather is not None:
                ds, idx_dev = gather
                mean, std = ds.norm_scalars()
                x_tbc = ops.sosfilt_gather(ds.eeg, idx_dev, ds.time_low, ds.time_high, self.sos, mean, std,
                                           out_layout="TBC", out_dtype=cd)
            else:
                x_tbc = ops.sosfilt(eeg_bct, self.sos, zero_phase=self.zero_phase, out_layout="TBC", out_dtype=cd)
        layers = m.lstm.layer_weights()
        act = ACT_RELU if m.include_top else ACT_NONE
        K = m.output.weight.shape[0]
        fused_head = (self.fused_head and self.center.numel() == K and m.output.bias is not None
                      and m.output.weight.data_ptr() % 16 == 0
                      and ops.head_dino_supported(B, m.output.weight.shape[1], K))
        with self._stage("encoder_fwd"):
            h_last, saved = encoder_fwd(x_tbc, layers, cd, training=True, cast_last=not fused_head)
        side = None
        with self._stage("head_loss"):
            if self.peer is not None:  # peers are done with the previous gradients; zero the centre-sum tail
                ops.dp_wait_done_zero(self.peer.flags, self.world, self._step_dev, self.batch_center)
            else:
                self.flat_g[self.n_param:].zero_()
            if fused_head:
                # Linear + loss + d_h in ONE kernel; the head's weight gradient (needs only d_pre and h_T) leaves the
                # critical path: it runs on a side stream next to the backward recurrence and is joined before Adam
                loss, d_hlast, d_pre = ops.head_dino_fwd_bwd(h_last, m.output.weight, m.output.bias, act, teacher, self.center,
                                                             self.loss.student_temp, tau_t, None)
                cur = torch.cuda.current_stream()
                if self._side is None:
                    self._side = torch.cuda.Stream(device=h_last.device)
                side = self._side
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    # centre statistics sum_b teacher (fixed summation order), also off the critical path
                    ops.colsum(teacher, out=self.batch_center, accumulate=True)
                    h32 = h_last if h_last.dtype == torch.float32 else ops.cast(h_last, torch.float32)
                    ops.gemm_f32(d_pre, h32, True, False, out=self.grad_of(m.output.weight), rowsum=self.grad_of(m.output.bias))
            else:
                emb, pre = linear_fwd(h_last, m.output.weight, m.output.bias, act)
                loss, d_emb, _ = ops.dino_loss_fwd_bwd(emb, teacher, self.center, self.loss.student_temp, tau_t,
                                                       DINO_SINGLE, batch_center=self.batch_center)
                d_hlast, _, _ = linear_bwd(h_last, m.output.weight, pre, d_emb, act, need_dx=True,
                                           dw_out=self.grad_of(m.output.weight), db_out=self.grad_of(m.output.bias))
        grads = [tuple(self.grad_of(w) for w in layer) for layer in layers]
        with self._stage("encoder_bwd"):
            encoder_bwd(d_hlast, layers, saved, grads, cd)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)  # the head's dW / db are in the gradient buffer
        if self.peer is not None:
            with self._stage("adam_center"):  # all-reduce + Adam + centre EMA in one kernel over peer memory
                pe = self.peer
                ops.dp_adam_step_peer(self.flat_p_opt, self.exp_avg, self.exp_avg_sq, pe.grad_ptrs, pe.flag_ptrs, self.world,
                                      pe.rank, self.center, self.loss.center_momentum, 1.0 / (B * self.world),
                                      self._step_dev, pe.ticket, self.lr, self.betas[0], self.betas[1], self.eps,
                                      self.weight_decay, self.decoupled, grad_scale=1.0 / self.world)
            return loss
        with self._stage("allreduce"):
            dp.allreduce_flat_(self.flat_g)  # gradients and centre statistics in one NCCL call
        with self._stage("adam_center"):
            ops.adam_step_graph(self.flat_p_opt, self.flat_g[:self.n_param], self.exp_avg, self.exp_avg_sq, self._step_dev,
                                self._adam_consts, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                self.decoupled, grad_scale=1.0 / self.world)
            ops.center_ema(self.center, self.batch_center, self.loss.center_momentum, 1.0 / (B * self.world))
        return loss

    def step(self, eeg_bct, teacher_feats, epoch=0):
        """eeg_bct: float32 [B, C, T] on the GPU (raw trials, stored layout); teacher_feats: float32 [B, K].
        Returns the loss as a 0-d device tensor (no host sync).  ALIASING: on the CUDA-graph path this is the graph's
        static output tensor -- the next step() overwrites it.  Consume it (add it to a running sum, .item(), .clone())
        before the next call; do not collect the returned tensors in a list."""
        tau_t = float(self.loss.teacher_temp_schedule[epoch])
        self.step_count += 1
        if not self.use_cuda_graph or self._stage_events is not None:
            return self._run(eeg_bct.contiguous(), teacher_feats.contiguous(), tau_t)
        return self._step_graphed(eeg_bct, teacher_feats, tau_t)

    def step_from_dataset(self, dataset, indices, teacher_feats, epoch=0):
        """One step on a batch of a GPU-RESIDENT dataset (dataset.DeviceEEGDataset): `indices` is a host int64 [B]
        (e.g. from dataset.epoch_batches), `teacher_feats` float32 [B, K] on the host (pinned) or the device.  The trials
        are fetched on the device (by the band-pass kernel's own loads when the shape allows, else by the gather
        kernel), so only the indices and the image features cross PCIe (8 B + 4 K bytes per trial instead of 4 C T) -- the replacement for DataLoader + `.to(device)`
        of LstmDistillFromDinoV2Train.py:358-363 when the .pth fits in HBM."""
        idx = dataset.check_indices(indices)
        B, K = idx.numel(), teacher_feats.shape[1]
        dev = dataset.eeg.device
        if self._idx_ring is None:
            # pinned index slots (a slot is rewritten only after its copy has completed) and TWO device-side input
            # slots (indices + image features), so the copies of step i run on a copy stream under step i - 1
            self._idx_ring = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._ds_slots = [{"idx": torch.zeros(B, dtype=torch.int64, device=dev),
                               "feats": torch.empty(B, K, dtype=torch.float32, device=dev),
                               "ready": torch.cuda.Event(), "consumed": torch.cuda.Event()} for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._idx_n = 0
        if self._ds_slots[0]["idx"].numel() != B:
            raise _lib.CsnError("step_from_dataset: batch size changed (%d -> %d)" % (self._ds_slots[0]["idx"].numel(), B))
        n = self._idx_n
        self._idx_n += 1
        h_idx, done = self._idx_ring[n % len(self._idx_ring)]
        slot = self._ds_slots[n % 2]
        done.synchronize()
        h_idx.copy_(idx)
        cur = torch.cuda.current_stream()
        on_host = not teacher_feats.is_cuda
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(slot["consumed"])  # the step that last read this slot has finished
            slot["idx"].copy_(h_idx, non_blocking=True)
            done.record(self._copy_stream)
            if on_host:
                slot["feats"].copy_(teacher_feats, non_blocking=True)
            slot["ready"].record(self._copy_stream)
        cur.wait_event(slot["ready"])
        if not on_host:  # produced on the caller's stream: copy it there
            slot["feats"].copy_(teacher_feats, non_blocking=True)
        try:
            if self.zero_phase or not dataset.fused_filter_ok():
                # general shapes: gather kernel into the static input buffer, then the ordinary step
                if self._static is None:
                    self._static = (torch.empty(B, dataset.C, dataset.samples, dtype=torch.float32, device=dev),
                                    torch.empty(B, K, dtype=torch.float32, device=dev))
                se, st = self._static
                dataset.gather_into(se, slot["idx"])
                st.copy_(slot["feats"], non_blocking=True)
                return self.step(se, st, epoch)
            # the filter's loads do the gather: the batch itself is never materialised
            tau_t = float(self.loss.teacher_temp_schedule[epoch])
            self.step_count += 1
            gather = (dataset, slot["idx"])
            if not self.use_cuda_graph or self._stage_events is not None or not self._warm:
                self._warm = True
                return self._run(None, slot["feats"], tau_t, gather=gather)
            gkey = (id(dataset), tau_t, n % 2)
            hit = self._ds_graphs.get(gkey)
            if hit is None:
                if len(self._ds_graphs) >= 8:  # epochs change tau: drop the stale captures
                    self._ds_graphs.clear()
                hit = self._ds_graphs[gkey] = self._capture(None, slot["feats"], tau_t, gather=gather)
            hit[0].replay()
            return hit[1]
        finally:
            slot["consumed"].record(cur)

    # CUDA-graph path.  The ~30 launches of a step are captured once and replayed with one submission.  Everything
    # that varies between steps lives on the device: the inputs are copied into static buffers (`input_buffers()`
    # exposes them so a loader can H2D straight into them), Adam's step count is a device counter
    # (csn_adam_step_graph).  The teacher temperature is a launch constant: the graph is re-captured when the
    # epoch changes it.  The first call runs eagerly (it also warms every lazily-initialised launch attribute).
    def input_buffers(self, like_eeg=None, like_feats=None):
        if self._static is None:
            if like_eeg is None:
                raise _lib.CsnError("input_buffers(): pass example tensors on the first call")
            self._static = (torch.empty_like(like_eeg), torch.empty_like(like_feats))
        return self._static

    def register_inputs(self, eeg_bct, teacher_feats):
        """Declare a RESIDENT (eeg, teacher) buffer pair that step() will be handed again and again (a loader's
        double buffer, a device-resident dataset shard).  Such a pair gets its own captured graph, so a replay reads
        the trials in place instead of first copying them into the graph's static input buffers (57.7 MB per cfg2
        step).  All graphs of a step object share one memory pool: they never run concurrently."""
        if not (eeg_bct.is_contiguous() and teacher_feats.is_contiguous()):
            raise _lib.CsnError("register_inputs: buffers must be contiguous")
        self._resident[(eeg_bct.data_ptr(), teacher_feats.data_ptr())] = (eeg_bct, teacher_feats)

    def _capture(self, eeg, teacher, tau_t, gather=None):
        torch.cuda.current_stream().synchronize()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self._pool):
            out = self._run(eeg, teacher, tau_t, gather=gather)
        return g, out

    def _step_graphed(self, eeg_bct, teacher, tau_t):
        key = (eeg_bct.data_ptr(), teacher.data_ptr())
        if key in self._resident and self._warm:
            gkey = key + (tau_t,)
            hit = self._resident_graphs.get(gkey)
            if hit is None:
                if len(self._resident_graphs) >= 64:  # epochs change tau: drop the stale captures
                    self._resident_graphs.clear()
                hit = self._resident_graphs[gkey] = self._capture(eeg_bct, teacher, tau_t)
            hit[0].replay()
            return hit[1]
        se, st = self.input_buffers(eeg_bct, teacher)
        if se.shape != eeg_bct.shape or st.shape != teacher.shape:
            raise _lib.CsnError("CUDA-graph step: batch shape changed (%s -> %s); build a new DistillTrainStep or pass "
                                "use_cuda_graph=False" % (tuple(se.shape), tuple(eeg_bct.shape)))
        if eeg_bct.data_ptr() != se.data_ptr():
            se.copy_(eeg_bct, non_blocking=True)
        if teacher.data_ptr() != st.data_ptr():
            st.copy_(teacher, non_blocking=True)
        if not self._warm:
            self._warm = True
            return self._run(se, st, tau_t)
        if self._graph is None or self._graph_tau != tau_t:
            g, out = self._capture(se, st, tau_t)
            self._graph, self._graph_tau, self._graph_out = g, tau_t, out
        self._graph.replay()
        return self._graph_out
