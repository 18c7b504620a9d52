"""GEMM kernels through the C-ABI: fp32 SIMT (tight) and tcgen05 bf16 (vs fp32 matmul of the bf16-rounded operands)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from cerebralsignalnetworks_b200 import ops
    return ops


def _mk(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(70, 45, 33), (128, 64, 16), (1, 7, 300), (257, 129, 65)])
def test_gemm_f32(ta, tb, M, N, K):
    ops = _ops()
    a = _mk((K, M) if ta else (M, K), 1)
    b = _mk((N, K) if tb else (K, N), 2)
    bias = _mk((N,), 3)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + bias.double()
    out = ops.gemm_f32(a.cuda(), b.cuda(), ta, tb, bias=bias.cuda()).cpu()
    np.testing.assert_allclose(out.numpy(), ref.float().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 384, 128), (384, 128, 256), (256, 128, 384), (20, 68, 4), (36, 200, 512), (16, 64, 36)])
def test_gemm_f32_resident_strip(ta, tb, M, N, K):
    """Short-contraction shapes (the projection head and its gradients) take the whole-K-resident kernel
    (16-byte aligned strips); the row sums of op(A) ride along (bias gradient of dW = dY^T X)."""
    ops = _ops()
    a = _mk((K, M) if ta else (M, K), 11)
    b = _mk((N, K) if tb else (K, N), 12)
    bias = _mk((N,), 13)
    opa = (a.t() if ta else a).double()
    ref = opa @ (b.t() if tb else b).double() + bias.double()
    rs = torch.full((M,), float("nan"), device="cuda")
    out = ops.gemm_f32(a.cuda(), b.cuda(), ta, tb, bias=bias.cuda(), rowsum=rs).cpu()
    np.testing.assert_allclose(out.numpy(), ref.float().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(rs.cpu().numpy(), opa.sum(1).float().numpy(), rtol=1e-4, atol=1e-4)
    # accumulate into C (beta) and activation epilogue
    c0 = _mk((M, N), 14)
    out2 = ops.gemm_f32(a.cuda(), b.cuda(), ta, tb, out=c0.clone().cuda(), alpha=0.5, beta=2.0).cpu()
    ref2 = 0.5 * (opa @ (b.t() if tb else b).double()) + 2.0 * c0.double()
    np.testing.assert_allclose(out2.numpy(), ref2.float().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("a_mn,b_mn", [(0, False), (1, False), (0, True), (1, True), (2, False), (2, True)])
@pytest.mark.parametrize("N,K", [(16, 16), (16, 128), (64, 64), (256, 32)])
def test_umma_tile_descriptors(a_mn, b_mn, N, K):
    """One tcgen05.mma chain on thread-staged canonical (no-swizzle) operands: the layout the recurrence uses."""
    ops = _ops()
    a = _mk((128, K), 4).bfloat16()
    b = _mk((N, K), 5).bfloat16()
    ref = a.float() @ b.float().t()
    d = ops.dbg_umma_tile(a.cuda(), b.cuda(), a_mn, b_mn).cpu()
    np.testing.assert_allclose(d.numpy(), ref.numpy(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("ta,tb", [(False, True), (True, False), (False, False), (True, True)])
@pytest.mark.parametrize("M,N,K,split", [(128, 128, 64, 1), (256, 384, 128, 1), (200, 72, 136, 1), (512, 128, 1024, 4),
                                         (130, 40, 72, 2), (1000, 512, 128, 1)])
def test_gemm_bf16_tc(ta, tb, M, N, K, split):
    ops = _ops()
    # TMA needs a 16-byte row pitch: stored leading dims must be multiples of 8 elements
    if ((M if ta else K) % 8) or ((K if tb else N) % 8):
        pytest.skip("leading dimension not a multiple of 8 (rejected by the ABI, see test below)")
    a = _mk((K, M) if ta else (M, K), 6).bfloat16()
    b = _mk((N, K) if tb else (K, N), 7).bfloat16()
    bias = _mk((N,), 8)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float()) + bias
    out = ops.gemm_bf16(a.cuda(), b.cuda(), ta, tb, bias=bias.cuda(), split_k=split).cpu()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("ta,tb", [(False, True), (True, False), (False, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(5000, 1024, 128), (5003, 1000, 72), (130, 40960, 256), (40000, 128, 512)])
def test_gemm_bf16_persistent_tiles(ta, tb, M, N, K):
    """Products with hundreds of output tiles and a short contraction walk the tiles with persistent CTAs (double-buffered
    accumulators in tensor memory): ragged edges in both directions, every operand layout, fp32 and bf16 output."""
    ops = _ops()
    if ((M if ta else K) % 8) or ((K if tb else N) % 8):
        pytest.skip("leading dimension not a multiple of 8 (rejected by the ABI)")
    a = _mk((K, M) if ta else (M, K), 16).bfloat16()
    b = _mk((N, K) if tb else (K, N), 17).bfloat16()
    bias = _mk((N,), 18)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float()) + bias
    out = ops.gemm_bf16(a.cuda(), b.cuda(), ta, tb, bias=bias.cuda()).cpu()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (err, scale)
    ob = ops.gemm_bf16(a.cuda(), b.cuda(), ta, tb, out_dtype=torch.bfloat16).cpu().float()
    assert (ob - (ref - bias)).abs().max().item() <= 1e-2 * scale + 1e-2


def test_gemm_bf16_accumulate_and_bf16_out():
    ops = _ops()
    a = _mk((256, 192), 9).bfloat16().cuda()
    b = _mk((128, 192), 10).bfloat16().cuda()
    base = _mk((256, 128), 11).cuda()
    out = base.clone()
    ops.gemm_bf16(a, b, False, True, out=out, accumulate=True)
    ref = base + a.float() @ b.float().t()
    assert torch.allclose(out, ref, rtol=2e-3, atol=2e-3)
    ob = ops.gemm_bf16(a, b, False, True, out_dtype=torch.bfloat16)
    assert ob.dtype == torch.bfloat16 and torch.allclose(ob.float(), a.float() @ b.float().t(), rtol=2e-2, atol=2e-2)


def test_gemm_bf16_rejects_unaligned_pitch():
    from cerebralsignalnetworks_b200 import _lib
    ops = _ops()
    with pytest.raises(_lib.CsnError):
        ops.gemm_bf16(torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda"),
                      torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda"), False, True)
