#!/usr/bin/env python
"""bench.py -- EEG trials/sec of the distillation train step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config[1] of BASELINE.json, the one the metric is quoted on): per GPU 256 raw trials of 128 channels x 440
samples -> fused 5-95 Hz order-4 Butterworth band-pass -> 1-layer LSTM (hidden 128, bf16 tcgen05 recurrence)
-> Linear(128, 384) -> DINO cross-entropy vs random 384-d teacher features (teacher-temperature warm-up, centre EMA)
-> BPTT -> Adam.  Weak scaling: per-GPU batch fixed, gradients + centre all-reduced over NCCL.  Synthetic data
(the reference's generator formula), random-init weights.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(batch_per_gpu=256, channels=128, samples=440, hidden=128, layers=1, feat=384, fs=1000.0, band=(5.0, 95.0),
           order=4, lr=1e-3, seed=43, n_resident_batches=8)
WORKLOAD = ("distill_step cfg2: 128ch x 440 samples, LSTM L1 H128, 384-d targets, fused 5-95 Hz band-pass, "
            "DINO CE + centre EMA, Adam")
METRIC = "eeg_trials_per_sec_distill_train_step"
UNIT = "trials/s"


def algorithmic_work(B, C, T, H, L, K):
    """Per-step algorithmic work (SURVEY.md section 8d / BASELINE.md section 4)."""
    I = C
    lstm_fwd_flop = 8.0 * T * H * sum((I if l == 0 else H) + H for l in range(L)) * B
    rec_fwd_flop = 8.0 * T * H * H * L * B  # recurrent part only (the persistent kernel's MMAs)
    return dict(filter_bytes=8.0 * C * T * B, loss_bytes=12.0 * K * B, lstm_fwd_flop=lstm_fwd_flop,
                lstm_train_flop=3.0 * lstm_fwd_flop, rec_fwd_flop=rec_fwd_flop, rec_bwd_flop=rec_fwd_flop)


def ncu_traffic(kernel_substrings):
    """Per-launch DRAM bytes (read + write) of the named kernels from the committed ncu --set full summary
    (profiles/traffic.json, written by scripts/summarize_profiles.py); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    ks = json.load(open(p))["kernels"]
    total = 0.0
    for sub in kernel_substrings:
        hit = [v for k, v in ks.items() if sub in k]
        if not hit:
            return None
        total += hit[0]["dram_read_bytes"] + hit[0]["dram_write_bytes"]
    return total


def ncu_tensor_pipe():
    """Active-SM tensor-pipe utilisation of the two recurrence kernels from the committed ncu --set full summary
    (sm__pipe_tensor_cycles_active / sm__pipe_tc_cycles_active, % of peak sustained on active SMs)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    d = json.load(open(p))
    out = {"source": d.get("source"), "tag": d.get("tag")}
    for k, v in d["kernels"].items():
        for short in ("lstm_fwd_tc_kernel", "lstm_bwd_tc_kernel"):
            if short in k and v.get("tensor_pipe_active_pct") is not None:
                out[short] = {"sm__pipe_tensor_cycles_active_pct": v["tensor_pipe_active_pct"],
                              "sm__pipe_tc_cycles_active_pct": v.get("tc_pipe_active_pct")}
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


def bind_to_gpu_numa_node(torch, device_index):
    """Multi-GPU runs: pin this rank's host thread to the CPUs NVML reports as local to its GPU BEFORE any pinned host
    buffer is allocated, so that the buffers the end-to-end leg copies from live on the GPU's own NUMA node (Linux
    allocates on the node of the touching thread).  With all ranks' pinned memory on one node the eight H2D streams of an
    8-GPU box share that node's memory controllers.  Returns the number of CPUs bound to (0: left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # no NVML / no affinity information: keep the default placement
        return 0


class ClockSampler:
    """SM clock + throttle reasons DURING the timed regions.  NVML is polled from a thread every few ms (the timed
    region of a default run is short; `nvidia-smi -lms` needs ~0.5 s to produce its first row); nvidia-smi is the
    fallback when pynvml cannot be initialised.  start()/pause() bracket each timed region, summary() reports."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, torch, device_index, period_s=0.004):
        import threading
        self.period = period_s
        self.samples = []          # (sm_mhz, reasons_bitmask, power_w)
        self.sm_max = None
        self.h = None
        self.nv = None
        self._on = threading.Event()
        self._quit = False
        self.backend = "none"
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.backend = "nvml"
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        except Exception:
            self.h = None
            self._smi_index = device_index
            self._smi = None
            self._smi_file = None

    def _reasons(self):
        nv = self.nv
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            fn = getattr(nv, name, None)
            if fn is not None:
                try:
                    return int(fn(self.h))
                except Exception:
                    continue
        return 0

    def _loop(self):
        nv = self.nv
        while not self._quit:
            self._on.wait()
            if self._quit:
                break
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((mhz, self._reasons(), pw))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self._on.set()
            return
        try:  # fallback: nvidia-smi loop (coarse)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            if self._smi is None:
                self._smi_file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(self._smi_index), "--query-gpu=" + q,
                                              "--format=csv,noheader,nounits", "-lms", "50"],
                                             stdout=self._smi_file, stderr=subprocess.DEVNULL)
                self.backend = "nvidia-smi"
        except Exception:
            self._smi = None

    def pause(self):
        if self.h is not None:
            self._on.clear()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "backend": self.backend}
        if self.h is not None:
            self._quit = True
            self._on.set()
            rows = list(self.samples)
            if rows:
                out["sm_mhz"] = statistics.median(r[0] for r in rows)
                out["sm_min_mhz"] = min(r[0] for r in rows)
                mask = 0
                for r in rows:
                    mask |= r[1]
                out["reasons"] = sorted(n for bit, n in self.REASONS.items() if mask & bit)
                pw = [r[2] for r in rows if r[2] is not None]
                if pw:
                    out["power_w_max"] = max(pw)
                out["samples"] = len(rows)
            return out
        if getattr(self, "_smi", None) is None:
            return out
        self._smi.terminate()
        try:
            self._smi.wait(timeout=5)
        except Exception:
            self._smi.kill()
        self._smi_file.flush()
        rows = [r.split(",") for r in open(self._smi_file.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self._smi_file.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        out["sm_mhz"] = statistics.median(sm) if sm else None
        try:
            out["sm_max_mhz"] = float(rows[0][2])
        except Exception:
            pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(rows)
        return out


def make_inputs(torch, n_batches, B, C, T, K, seed, device):
    """Synthetic Spampinato-shaped batches resident in HBM: N(0,1) + 0.5 sin(2 pi 40 t / fs)
    (utils/PerilsEEGDataset.py:140-147), teacher features N(0,1)."""
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(T, device=device, dtype=torch.float32) / CFG["fs"]
    wave = 0.5 * torch.sin(2 * torch.pi * 40.0 * t)
    eeg = [torch.randn(B, C, T, device=device, generator=g) + wave for _ in range(n_batches)]
    feats = [torch.randn(B, K, device=device, generator=g) for _ in range(n_batches)]
    return eeg, feats


def cpu_reference_step_rate(batch, steps, warmup, threads=None):
    """The reference PyTorch path on the host cores: scipy sosfilt -> restated Model on torch.nn.LSTM ->
    DINOLoss -> backward -> Adam (oracle/distill.py).  Returns (trials/s, ms/step, cores)."""
    cores = threads or os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all the host cores it can (set before torch's
    # thread pools are created, and again through the torch API in case they already exist)
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    import torch
    from oracle.distill import DistillStepOracle, synthetic_batch

    torch.set_num_threads(cores)
    o = DistillStepOracle(CFG["channels"], CFG["hidden"], CFG["layers"], CFG["feat"], include_top=False, nepochs=100,
                          lr=CFG["lr"], low_hz=CFG["band"][0], high_hz=CFG["band"][1], fs=CFG["fs"], order=CFG["order"],
                          seed=CFG["seed"])
    eeg, feats, _ = synthetic_batch(batch, CFG["channels"], CFG["samples"], CFG["feat"], seed=CFG["seed"])
    for _ in range(warmup):
        o.step(eeg, feats, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(eeg, feats, 0)
    dt = time.perf_counter() - t0
    return batch * steps / dt, 1e3 * dt / steps, cores


# ---------------------------------------------------------------------------------------------------------------------
# The other GPU configurations of BASELINE.json, measured in the same run at the same N (sub-records of the JSON line).
def _timed_steps(torch, dist, world, dev, fn, n, warm):
    """warm untimed + n timed calls of fn(i) bracketed by barrier + synchronize, CUDA events, max over ranks -> ms per call."""
    for i in range(warm):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(warm + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / n


def bench_cfg3(torch, dist, csn, dev, world, rank, steps, peaks):
    """configs[2]: LstmDistillation DINO step -- 64 trials per GPU of 96 ch x 495 samples, 2 x 300 + 4 x 200 crops, 4-layer
    LSTM (hidden 128) student / teacher, DINOHead K = 65536, reference multi-crop loss, two-shot gradient + centre exchange,
    fused clip + AdamW + EMA sweep (MultiCropDistillStep)."""
    import numpy as np
    from cerebralsignalnetworks_b200.schedules import cosine_scheduler
    B, C, T, H, L, K = 64, 96, 495, 128, 4, 65536
    torch.manual_seed(CFG["seed"])
    np.random.seed(CFG["seed"])  # same crop starts on every rank

    def make():
        return csn.MultiCropWrapper(csn.Model(C, H, L, H, include_top=False, compute_dtype=torch.bfloat16),
                                    csn.DINOHead(H, K, compute_dtype=torch.bfloat16)).to(dev)
    student, teacher = make(), make()
    crit = csn.DINOLoss(K, 6, 0.04, 0.04, 30, 100).to(dev)
    n_it = 4096
    lr = cosine_scheduler(5e-4 * B * world / 256.0, 1e-6, 1, n_it, warmup_epochs=0)
    wd = cosine_scheduler(0.04, 0.4, 1, n_it)
    mom = cosine_scheduler(0.996, 1.0, 1, n_it)
    step = csn.MultiCropDistillStep(student, teacher, crit, lr, wd, mom, clip_grad=3.0, freeze_last_layer=0, batch_size=B)
    g = torch.Generator(device=dev).manual_seed(31 + rank)
    eeg = [torch.randn(B, T, C, device=dev, generator=g) for _ in range(4)]
    n = max(3, min(steps, 15))
    ms = _timed_steps(torch, dist, world, dev, lambda i: step.step(eeg[i % 4], epoch=0), n, 3)
    # algorithmic work per trial: LSTM 4.88 GFLOP (student fwd + bwd, teacher fwd) + DINO head 0.87 GFLOP
    flop = (4.88e9 + 0.87e9) * B
    tf = flop / (ms * 1e-3) / 1e12
    peak = peaks["tflops_sustained"] or peaks["tflops"]
    n_param = step.opt.n_flat
    del step, student, teacher, crit, eeg
    torch.cuda.empty_cache()
    return {"workload": "cfg3 LstmDistillation DINO step: 96ch x 495, crops 2x300 + 4x200, LSTM L4 H128 student + teacher, "
                        "DINOHead K=65536, multi-crop loss, clip + AdamW + EMA", "batch_per_gpu": B, "global_batch": B * world,
            "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": n,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "note": "5.75 GFLOP algorithmic per trial (SURVEY 8d); three serial LSTM chains of 200-300 steps x 4 layers"},
            "exchange": {"kind": "two-shot reduce-scatter + all-gather over peer memory" if world > 1 else "none",
                         "floats": n_param + B * K, "bytes_per_rank_over_nvlink": (2.0 * (world - 1) / world * 4 * (n_param + B * K)) if world > 1 else 0}}


def bench_cfg4(torch, dist, csn, dev, world, rank, steps, peaks):
    """configs[3]: stacked 2-layer LSTM hidden 512, 768-d targets, global batch 1024 sharded over the ranks (N = 1: the
    128-trial shard of the 8-GPU run)."""
    B = 1024 // world if world > 1 else 128
    torch.manual_seed(CFG["seed"])
    model = csn.Model(128, 512, 2, 768, include_top=False, compute_dtype=torch.bfloat16).to(dev)
    crit = csn.DINOLoss(768, 1, 1.5, 0.22, 50, 100).to(dev)
    step = csn.DistillTrainStep(model, crit, lr=1e-3, sos=csn.EEGFilters(1000.0).sos(5.0, 95.0, 4))
    g = torch.Generator(device=dev).manual_seed(41 + rank)
    eeg = [torch.randn(B, 128, 440, device=dev, generator=g) for _ in range(3)]
    feats = [torch.randn(B, 768, device=dev, generator=g) for _ in range(3)]
    for e_, f_ in zip(eeg, feats):
        step.register_inputs(e_, f_)
    n = max(3, min(steps, 8))
    ms = _timed_steps(torch, dist, world, dev, lambda i: step.step(eeg[i % 3], feats[i % 3], epoch=0), n, 4)
    flop = 8.997e9 * B  # train step, L2 H512 I128 T440
    tf = flop / (ms * 1e-3) / 1e12
    peak = peaks["tflops_sustained"] or peaks["tflops"]
    ex = step.dp_exchange
    step._graph = None
    del step, model, crit, eeg, feats
    torch.cuda.empty_cache()
    return {"workload": "cfg4 distill step: 128ch x 440, stacked LSTM L2 H512, 768-d targets, band-pass, DINO CE, Adam",
            "batch_per_gpu": B, "global_batch": B * world, "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "steps": n, "dp_exchange": ex,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "note": "8.997 GFLOP algorithmic per trial (SURVEY 8d); persistent 16-CTA cluster recurrence (W_hh split over the cluster, "
                                 "resident in tensor memory; per-step DSMEM all-gather / reduce-scatter), one launch per layer and direction"}}


def bench_cfg5(torch, dist, csn, dev, world, rank, steps, peaks):
    """configs[4]: Perils-shaped trials (63 ch x 2000 samples) -> band-pass -> encoder inference -> exact top-5 (squared L2,
    what the reference's faiss IndexFlatL2 returns) over a 10 k gallery.  Replicas: every rank serves its own queries."""
    from cerebralsignalnetworks_b200 import retrieval
    B, C, T, H, D, NB, K = 256, 63, 2000, 128, 384, 10000, 5
    torch.manual_seed(CFG["seed"])
    model = csn.Model(C, H, 1, D, include_top=False, compute_dtype=torch.bfloat16).to(dev).eval()
    sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 4)
    g = torch.Generator(device=dev).manual_seed(51 + rank)
    eeg = [torch.randn(B, C, T, device=dev, generator=g) for _ in range(3)]
    gallery = torch.randn(NB, D, device=dev, generator=g)

    def query(i):
        emb = model.encode_trials(eeg[i % 3], sos=sos)
        return retrieval.topk_search(gallery, emb, K, retrieval.METRIC_L2)
    n = max(3, min(steps, 10))
    ms = _timed_steps(torch, dist, world, dev, query, n, 2)
    filt_bytes = 8.0 * C * T * B
    del model, eeg, gallery
    torch.cuda.empty_cache()
    return {"workload": "cfg5 retrieval: 63ch x 2000 trials -> band-pass -> LSTM L1 H128 encoder inference -> top-5 L2 over a "
                        "10k x 384 gallery", "batch_per_gpu": B, "value": world * B / (ms * 1e-3), "unit": "queries/s",
            "ms_per_batch": ms, "steps": n, "scaling": "replicas",
            "roofline": {"bound": "tensor (serial chain)", "note": "2000 dependent recurrence steps of 256 trials: latency bound; "
                         "filter algorithmic bytes %d per batch" % filt_bytes,
                         "achieved": 8.0 * T * H * (64 + H) * B / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "peak": peaks["tflops_sustained"] or peaks["tflops"],
                         "frac": 8.0 * T * H * (64 + H) * B / (ms * 1e-3) / 1e12 / (peaks["tflops_sustained"] or peaks["tflops"])}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # The cfg2 step itself (256 trials) when its K steps fit the host-time budget, else a bounded sample of it: the
    # largest of 128/64/32/16 trials per step that does (probed with one timed step each; CPU step time is linear in
    # the batch)
    budget_s = float(os.environ.get("CSN_REF_BUDGET_S", "150"))
    batch = 16
    for cand in (CFG["batch_per_gpu"], 128, 64, 32, 16):
        _, ms_probe, _ = cpu_reference_step_rate(cand, 1, 1)
        batch = cand
        if ms_probe * 1e-3 * (args.steps + min(args.warmup, 3)) <= budget_s:
            break
    rate, ms, cores = cpu_reference_step_rate(batch, args.steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": batch, "batch_per_step": batch, "same_batch_as_gpu_arm": batch == CFG["batch_per_gpu"]},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{batch}-trial steps of the cfg2 workload (torch-CPU LSTM + scipy sosfilt + DINO loss + Adam)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--batch_per_gpu", type=int, default=CFG["batch_per_gpu"])
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--configs", type=str, default="3,4,5", help="extra BASELINE.json configs measured as sub-records ('' = none)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus > 1 launch through torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(torch, local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.warmup < 3:
        args.warmup = 3

    B, C, T, H, L, K = args.batch_per_gpu, CFG["channels"], CFG["samples"], CFG["hidden"], CFG["layers"], CFG["feat"]
    dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    torch.manual_seed(CFG["seed"])  # identical initial weights on every rank (DDP broadcast equivalent)
    model = csn.Model(C, H, L, K, include_top=False, compute_dtype=dtype).to(dev)
    crit = csn.DINOLoss(K, 1, 1.5, 0.22, 50, 100).to(dev)
    sos = csn.EEGFilters(CFG["fs"]).sos(CFG["band"][0], CFG["band"][1], CFG["order"])
    step = csn.DistillTrainStep(model, crit, lr=CFG["lr"], sos=sos)

    NB = CFG["n_resident_batches"]  # 8 x 57.7 MB of raw trials per GPU: the rotation exceeds the 126 MB L2
    eeg, feats = make_inputs(torch, NB, B, C, T, K, CFG["seed"] + 1000 * rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for e_, f_ in zip(eeg, feats):  # resident batches are read in place by their own captured graph
        step.register_inputs(e_, f_)

    # ---------------- device-resident throughput (`value`) ----------------
    # The step runs as one replayed CUDA graph (DistillTrainStep default).  Warm-up covers the eager first call and
    # the capture.
    for i in range(max(args.warmup, NB + 1)):  # first call is eager; every resident batch gets its capture here
        step.step(eeg[i % NB], feats[i % NB], epoch=0)
    barrier()
    sampler = ClockSampler(torch, local) if rank == 0 else None
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step.step(eeg[i % NB], feats[i % NB], epoch=0)
    e1.record()
    barrier()
    if rank == 0:
        sampler.pause()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total)
    final_loss = float(loss)

    # ---------------- per-stage device times (eager launches, CUDA events on the launching stream) ----------------
    # graph replays cannot carry timing events, so the stage breakdown / live roofline numbers come from a few eager
    # steps of the same workload; kernel durations are identical, only the host gaps between launches differ.
    step.enable_stage_timing(True)
    step.step(eeg[0], feats[0], epoch=0)  # untimed: graph capture emptied the eager allocator cache; refill it first
    barrier()
    step.enable_stage_timing(True)        # reset the recorded events
    launches0 = _lib.launch_count()
    n_stage_steps = min(args.steps, 5)
    for i in range(n_stage_steps):
        torch.cuda._sleep(int(8e6))  # ~4 ms spin: the host enqueues the whole step behind it, so the CUDA-event deltas
        step.step(eeg[i % NB], feats[i % NB], epoch=0)  # below are device time only (no launch gaps)
    barrier()
    launches_per_step = (_lib.launch_count() - launches0) // n_stage_steps
    launches = launches_per_step * args.steps  # the timed region replays exactly these launches every step
    stages = step.stage_times_ms()
    step.enable_stage_timing(False)

    # DINO loss kernel at the cfg3 shape (multi-crop, K=65536, 64 trials): the HBM-bound case of SURVEY section 8d
    loss_gbs = None
    try:
        from cerebralsignalnetworks_b200 import ops as _ops
        Vs, Vt, Bl, Kl = 6, 2, 64, 65536
        g = torch.Generator(device=dev).manual_seed(7)
        sets = [(torch.randn(Vs, Bl, Kl, device=dev, generator=g), torch.randn(Vt, Bl, Kl, device=dev, generator=g)) for _ in range(3)]
        cen = torch.zeros(Bl * Kl, device=dev)
        for s_, t_ in sets:
            _ops.dino_loss_fwd_bwd(s_, t_, cen, 0.1, 0.04, _lib.DINO_MULTICROP_REF)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bc = torch.zeros(Bl * Kl, device=dev)
        outs = []
        l0.record()
        for r in range(9):
            s_, t_ = sets[r % 3]   # 3 x 134 MB of inputs rotate through the 126 MB L2
            outs.append(_ops.dino_loss_fwd_bwd(s_, t_, cen, 0.1, 0.04, _lib.DINO_MULTICROP_REF, batch_center=bc)[0])
        l1.record()
        torch.cuda.synchronize()
        loss_ms = l0.elapsed_time(l1) / 9
        loss_bytes = (Vs + Vt + Vs) * 4.0 * Kl * Bl  # student + teacher read, gradient written (3 670 016 B/trial)
        loss_gbs = loss_bytes / (loss_ms * 1e-3) / 1e9
        del sets, outs, bc, cen
    except Exception as exc:  # report, do not hide
        print("loss roofline probe failed:", exc, file=sys.stderr)

    # ---------------- end to end through the public API with HOST buffers (`e2e`) ----------------
    # pinned host trials -> H2D on a copy stream (three staging buffers, copies run back to back under the steps) -> step ->
    # D2H of the loss every step.
    NBUF = 3  # device staging buffers: the copy of batch i+1 never waits for the step that used its buffer last
    h_eeg = [e.cpu().pin_memory() for e in eeg[:2]]
    h_feat = [f.cpu().pin_memory() for f in feats[:2]]
    d_eeg = [torch.empty_like(eeg[0]) for _ in range(NBUF)]
    d_feat = [torch.empty_like(feats[0]) for _ in range(NBUF)]
    h_loss = torch.zeros((), dtype=torch.float32).pin_memory()
    for s_ in range(NBUF):  # the loader's staging buffers: read in place by the step
        step.register_inputs(d_eeg[s_], d_feat[s_])
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(NBUF)]
    consumed = [torch.cuda.Event() for _ in range(NBUF)]

    def stage_in(i):
        s = i % NBUF
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            d_eeg[s].copy_(h_eeg[i % 2], non_blocking=True)
            d_feat[s].copy_(h_feat[i % 2], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(NBUF):
            consumed[s].record()
        stage_in(0)
        if n > 1:
            stage_in(1)
        for i in range(n):
            s = i % NBUF
            if i + 2 < n:
                stage_in(i + 2)
            torch.cuda.current_stream().wait_event(ready[s])
            l = step.step(d_eeg[s], d_feat[s], epoch=0)
            consumed[s].record()
            h_loss.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_loop(2 * NBUF)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    clocks = sampler.summary() if rank == 0 else None
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms)

    # ---------------- same step fed from a GPU-RESIDENT dataset (`e2e_resident_dataset`) ----------------
    # The .pth of the reference (a few GB) fits in HBM many times over: DeviceEEGDataset holds the raw trials on the
    # device, a step receives host INDICES + host image features, the batch is gathered on the device.  Reported NEXT
    # TO `e2e` (which stays the strict host-buffer path): this is the framework's answer to the PCIe bound above.
    ds_ms, n_ds = None, 8 * B
    ds = None
    try:
        from cerebralsignalnetworks_b200.dataset import DeviceEEGDataset
        g = torch.Generator(device=dev).manual_seed(11 + rank)
        raw = torch.randn(n_ds, C, T + 60, device=dev, generator=g)
        ds = DeviceEEGDataset.from_tensor(raw, time_low=20, time_high=20 + T, device=dev)
        del raw
        order = torch.Generator().manual_seed(5)
        batches = list(ds.epoch_batches(B, shuffle=True, generator=order))
    except Exception as exc:  # report, do not hide
        print("resident-dataset leg: setup failed:", repr(exc), file=sys.stderr)
        ds = None
    ds_ok = torch.tensor([1.0 if ds is not None else 0.0], device=dev)
    if world > 1:  # every rank runs the leg or none does (its steps contain the gradient exchange)
        dist.all_reduce(ds_ok, op=dist.ReduceOp.MIN)
    if float(ds_ok) > 0:
        def ds_loop(n):
            for i in range(n):
                l = step.step_from_dataset(ds, batches[i % len(batches)], h_feat[i % 2], epoch=0)
                h_loss.copy_(l, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ds_loop(max(3, args.warmup))
        barrier()
        t0 = time.perf_counter()
        ds_loop(args.steps)
        barrier()
        ds_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev)
        if world > 1:
            dist.all_reduce(ds_ms, op=dist.ReduceOp.MAX)
        ds_ms = float(ds_ms)

    # ---------------- the other configurations, same run, same N (sub-records) ----------------
    extra = {}
    peaks_ = measured_peaks()
    step._graph = None
    for tag, fn in (("3", bench_cfg3), ("4", bench_cfg4), ("5", bench_cfg5)):
        if tag not in [c.strip() for c in args.configs.split(",") if c.strip()]:
            continue
        try:
            extra["cfg" + tag] = fn(torch, dist, csn, dev, world, rank, args.steps, peaks_)
        except Exception as exc:  # report, do not hide (deterministic across ranks: nobody is left in a collective)
            extra["cfg" + tag] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            print("cfg%s sub-bench failed: %r" % (tag, exc), file=sys.stderr)

    def finish():
        # Multi-rank teardown: every rank meets at a barrier, then leaves without tearing NCCL down.  (Destroying the
        # process group while captured graphs still hold NCCL kernels hung rank teardown for minutes on the GPU box;
        # the OS reclaims the communicator.)
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            step._graph = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)
        return 0

    if rank != 0:
        return finish()

    work = algorithmic_work(B, C, T, H, L, K)
    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    # dominant stage: the encoder (fwd+bwd); its tensor-core roofline uses the LSTM's algorithmic flops
    enc_ms = stages.get("encoder_fwd", 0.0) + stages.get("encoder_bwd", 0.0)
    achieved_tflops = work["lstm_train_flop"] / (enc_ms * 1e-3) / 1e12 if enc_ms > 0 else None
    peak_tf = peaks["tflops_sustained"] or peaks["tflops"]
    filt_ms = stages.get("filter", 0.0)
    filt_gbs = work["filter_bytes"] / (filt_ms * 1e-3) / 1e9 if filt_ms > 0 else None
    phys_filter_bytes = (4.0 + (2.0 if dtype == torch.bfloat16 else 4.0)) * C * T * B
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if dtype == torch.bfloat16 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2_policy": f"rotating {NB} resident input batches ({NB * B * C * T * 4 / 1e6:.0f} MB) > 126 MB L2"},
        "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": B * C * T * 4 + B * K * 4, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                "host_numa": ("rank bound to the %d CPUs local to its GPU before the pinned buffers were allocated" % numa_cpus)
                             if numa_cpus else "default placement"},
        "e2e_resident_dataset": None if ds_ms is None else {
            "value": world * B * args.steps / (ds_ms * 1e-3), "unit": UNIT, "ms_per_step": ds_ms / args.steps,
            "h2d_bytes_per_step": B * 8 + B * K * 4, "d2h_bytes_per_step": 4, "dataset_trials_per_gpu": n_ds,
            "note": "DeviceEEGDataset: raw trials uploaded once, per step host indices + host image features -> device gather -> step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "LSTM encoder fwd+bwd (lstm_fwd_tc_kernel: 128 recurrence CTAs + 20 input-projection server CTAs; "
                               "lstm_bwd_tc_kernel: 128 recurrence CTAs + 20 dW consumer CTAs; bptt_finalize_kernel)",
                     "achieved": achieved_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tflops / peak_tf) if achieved_tflops else None,
                     "traffic": ncu_traffic(["lstm_fwd_tc_kernel", "lstm_bwd_tc_kernel", "bptt_finalize_kernel"]),
                     "traffic_note": "DRAM read+write bytes per step of the encoder kernels (forward launch incl. the fp16 input projection its server CTAs write, backward launch incl. the dW its consumer CTAs fold, the fixed-order finalize), ncu --set full, profiles/traffic.json",
                     "algorithmic_flop_per_step": work["lstm_train_flop"], "stage_ms": enc_ms,
                     "peak_source": peaks["source"] + " (sustained: kernel timed inside the step)",
                     "tensor_pipe_active_sm": ncu_tensor_pipe()},
        "roofline_filter": {"bound": "hbm", "kernel": "sosfilt_warp_kernel", "achieved": filt_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": (filt_gbs / peaks["hbm_gbs"]) if filt_gbs else None, "traffic": ncu_traffic(["sosfilt_warp_kernel"]),
                            "algorithmic_bytes": work["filter_bytes"], "algorithmic_note": "SURVEY 8(d): 8*C*T per trial (fp32 in + fp32 out)",
                            # the timed variant writes the encoder's bf16 [T,B,C] input: 4 + 2 bytes per sample really move
                            "physical_bytes": phys_filter_bytes,
                            "achieved_physical": (phys_filter_bytes / (filt_ms * 1e-3) / 1e9) if filt_ms > 0 else None,
                            "frac_physical": (phys_filter_bytes / (filt_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if filt_ms > 0 else None,
                            "peak_source": peaks["source"]},
        "roofline_loss": {"bound": "hbm", "kernel": "dino_loss_staged_kernel (cfg3 shape: 6 student + 2 teacher views, 64 trials, K=65536)",
                          "achieved": loss_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": (loss_gbs / peaks["hbm_gbs"]) if loss_gbs else None, "traffic": ncu_traffic(["dino_loss_staged_kernel"]),
                          "peak_source": peaks["source"]},
        "stages_ms": stages,
        "configs": extra,
        "step_submission": "cuda_graph_replay",
        "dp_exchange": step.dp_exchange,
        "loss": final_loss,
    }
    if not args.no_cpu_baseline and world == 1:  # (N > 1: the N = 1 line carries it; ranks must not idle behind a CPU loop)
        rate, ms, cores = cpu_reference_step_rate(16, 10, 3)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                                "sample": "10 steps of the cfg1 workload (batch 16, same model) on the host cores: torch-CPU LSTM + "
                                          "scipy sosfilt + DINO loss + Adam"}
    print(json.dumps(line), flush=True)
    return finish()


if __name__ == "__main__":
    sys.exit(main())
