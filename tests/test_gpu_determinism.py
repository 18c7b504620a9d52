"""Run-to-run determinism: the library holds no floating-point atomics (every cross-CTA sum is a fixed-order fold), so
the same inputs must give the same BITS on every run -- like the reference's CPU reductions (loss mean / centre sum of
LstmDistillFromDinoV2Train.py:78,95-105; autograd's weight / bias gradient sums behind :374).  Each case runs a kernel
(or the whole step) several times and compares with torch.equal."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPEATS = 4


def _same(results):
    first = results[0]
    for other in results[1:]:
        for a, b in zip(first, other):
            assert torch.equal(a, b)


@pytest.mark.parametrize("B,I,K", [(256, 128, 384), (600, 128, 384), (33, 64, 100)])
def test_fused_head_bits(B, I, K):
    from cerebralsignalnetworks_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + K)
    h = (torch.randn(B, I, device="cuda", generator=g) * 0.5).bfloat16()
    w = torch.randn(K, I, device="cuda", generator=g) * 0.2
    b = torch.randn(K, device="cuda", generator=g) * 0.1
    teacher = torch.randn(B, K, device="cuda", generator=g)
    center = torch.randn(K, device="cuda", generator=g) * 0.1
    runs = []
    for _ in range(REPEATS):
        bc = torch.zeros(K, device="cuda")
        loss, d_h, d_pre = ops.head_dino_fwd_bwd(h, w, b, 0, teacher, center, 0.1, 0.07, bc)
        runs.append((loss.clone(), d_h, d_pre, bc))
    _same(runs)


@pytest.mark.parametrize("mode,Vs,Vt,B,K", [(0, 1, 1, 256, 384), (0, 1, 1, 300, 768), (2, 4, 2, 64, 4096), (0, 1, 1, 96, 8192),
                                            (1, 6, 2, 16, 65536), (2, 3, 2, 5, 1000)])
def test_dino_loss_bits(mode, Vs, Vt, B, K):
    from cerebralsignalnetworks_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K + B)
    s = torch.randn(Vs, B, K, device="cuda", generator=g)
    t = torch.randn(Vt, B, K, device="cuda", generator=g)
    c = torch.randn(K, device="cuda", generator=g) * 0.1
    runs = []
    for _ in range(REPEATS):
        loss, ds, bc = ops.dino_loss_fwd_bwd(s, t, c, 0.1, 0.05, mode)
        runs.append((loss.clone(), ds, bc))
    _same(runs)
    if mode != 1:  # shared centre: the column sum over (view, row) against float64, within the fp32 bound of the sum
        exact = t.double().sum((0, 1)).cpu().numpy()
        bound = Vt * B * 2.0 ** -24 * t.abs().max().item()
        assert np.abs(runs[0][2].cpu().numpy().astype(np.float64) - exact).max() <= bound


@pytest.mark.parametrize("M,N", [(256, 384), (5, 7), (128, 65536), (20000, 512), (600, 40)])
def test_colsum_bits_and_bound(M, N):
    from cerebralsignalnetworks_b200 import ops
    x = torch.randn(M, N, device="cuda")
    outs = [ops.colsum(x) for _ in range(REPEATS)]
    _same([(o,) for o in outs])
    exact = x.double().sum(0).cpu().numpy()
    assert np.abs(outs[0].cpu().numpy().astype(np.float64) - exact).max() <= M * 2.0 ** -24 * x.abs().max().item()
    acc = torch.ones(N, device="cuda")
    ops.colsum(x, out=acc, accumulate=True)
    np.testing.assert_allclose(acc.cpu().numpy(), exact + 1.0, rtol=0, atol=(M + 1) * 2.0 ** -24 * (x.abs().max().item() + 1))


@pytest.mark.parametrize("T,B,I,H", [(40, 256, 128, 128), (24, 37, 96, 96), (12, 130, 128, 512)])
def test_lstm_bptt_bits(T, B, I, H):
    """Weight and bias gradients of the tcgen05 paths (persistent recurrence: per-CTA bias partials + split-K slabs;
    large-hidden path: split-K slabs) are bit-identical from run to run."""
    from cerebralsignalnetworks_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T + B)
    k = 1.0 / np.sqrt(H)
    w = [(torch.rand(4 * H, I, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) * k,
         (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k]
    x = torch.randn(T, B, I, device="cuda", generator=g).bfloat16()
    d_hlast = torch.randn(B, H, device="cuda", generator=g)
    runs = []
    for _ in range(REPEATS):
        h_seq, reserve, ws = ops.lstm_layer_fwd(x, *w, torch.bfloat16, True)
        grads = tuple(torch.full_like(t, float("nan")) for t in w)
        ops.lstm_layer_bwd(x, w[0], w[1], h_seq, reserve, ws, None, d_hlast, grads, False, torch.bfloat16)
        runs.append((h_seq,) + grads)
    _same(runs)


def test_split_k_gemm_bits_and_accumulate():
    from cerebralsignalnetworks_b200 import ops
    a = torch.randn(4096, 256, device="cuda").bfloat16()
    b = torch.randn(4096, 128, device="cuda").bfloat16()
    outs = [ops.gemm_bf16(a, b, True, False, split_k=7) for _ in range(REPEATS)]
    _same([(o,) for o in outs])
    ref = a.float().t() @ b.float()
    assert (outs[0] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    acc = torch.ones_like(outs[0])
    ops.gemm_bf16(a, b, True, False, out=acc, accumulate=True, split_k=7)
    assert (acc - 1 - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    acc = torch.ones_like(outs[0])
    ops.gemm_bf16(a, b, True, False, out=acc, accumulate=True, split_k=1)
    assert (acc - 1 - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_train_step_bits(dtype):
    """Three fused train steps (filter -> encoder -> head + loss -> BPTT -> Adam) from the same initial state, twice:
    loss, centre and every parameter agree bit for bit."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos
    B, C, T, H, D = 64, 32, 48, 32, 96
    g = torch.Generator().manual_seed(5)
    eeg = torch.randn(3, B, C, T, generator=g).cuda()
    feats = torch.randn(3, B, D, generator=g).cuda()
    outs = []
    for _ in range(2):
        torch.manual_seed(7)
        model = csn.Model(C, H, 1, D, include_top=False, compute_dtype=dtype).cuda()
        crit = csn.DINOLoss(D, 1, 1.5, 0.22, 5, 10).cuda()
        step = csn.DistillTrainStep(model, crit, lr=1e-3, sos=design_bandpass_sos(5.0, 95.0, 1000.0, 4))
        losses = [step.step(eeg[i], feats[i], epoch=1).clone() for i in range(3)]
        torch.cuda.synchronize()
        outs.append(tuple(losses) + (crit.center.clone(),) + tuple(p.detach().clone() for p in model.parameters()))
    _same(outs)
