"""Tensor-level wrappers over the C-ABI.  torch is used for device memory and streams only; every
arithmetic op here is a kernel of libcsn_b200.so launched on torch's current stream."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, F32, call

_LAYOUTS = {"BCT": _lib.LAYOUT_BCT, "BTC": _lib.LAYOUT_BTC, "TBC": _lib.LAYOUT_TBC}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _dt(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise _lib.CsnError("unsupported dtype %s (float32 / bfloat16 only)" % dtype)


def _chk(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.CsnError("%s must live on a CUDA device (no CPU fallback)" % name)
    if not t.is_contiguous():
        raise _lib.CsnError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise _lib.CsnError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def device_info():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    call("csn_device_info", C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


# ------------------------------------------------------------------------------------------------ filter
def sosfilt(x, sos, zero_phase=False, out_layout="BCT", out_dtype=torch.float32, out=None):
    """x: float32 [B, C, T] on the GPU; sos: float64 [n_sections, 6] (host).  Returns y in `out_layout`."""
    _chk(x, torch.float32, "x")
    if x.dim() != 3:
        raise _lib.CsnError("sosfilt expects [B, C, T], got shape %s" % (tuple(x.shape),))
    B, Cc, T = x.shape
    sos = np.ascontiguousarray(np.asarray(sos, dtype=np.float64))
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise _lib.CsnError("sos must have shape [n_sections, 6]")
    shape = {"BCT": (B, Cc, T), "BTC": (B, T, Cc), "TBC": (T, B, Cc)}[out_layout]
    y = out if out is not None else torch.empty(shape, dtype=out_dtype, device=x.device)
    _chk(y, out_dtype, "out")
    call("csn_sosfilt_f32", _p(x), _p(y), sos.ctypes.data_as(C.POINTER(C.c_double)), sos.shape[0], B, Cc, T,
         1 if zero_phase else 0, _LAYOUTS[out_layout], _dt(out_dtype), _stream())
    return y


def sosfilt_gather(src_nct, idx, time_low, time_high, sos, mean=0.0, std=1.0, out_layout="TBC", out_dtype=torch.float32):
    """Causal band-pass of the trials `idx` (device int64 [B], validated by the caller) of a resident [N, C, T_raw]
    array, cropped to [time_low, time_high) and optionally normalised -- gather fused into the filter's loads."""
    _chk(src_nct, torch.float32, "src")
    N, Cc, T_raw = src_nct.shape
    B, T = idx.numel(), time_high - time_low
    sos = np.ascontiguousarray(np.asarray(sos, dtype=np.float64))
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise _lib.CsnError("sos must have shape [n_sections, 6]")
    shape = {"BCT": (B, Cc, T), "BTC": (B, T, Cc), "TBC": (T, B, Cc)}[out_layout]
    y = torch.empty(shape, dtype=out_dtype, device=src_nct.device)
    call("csn_sosfilt_gather_f32", _p(src_nct), _p(idx), N, Cc, T_raw, int(time_low), int(time_high), float(mean), float(std),
         _p(y), sos.ctypes.data_as(C.POINTER(C.c_double)), sos.shape[0], B, _LAYOUTS[out_layout], _dt(out_dtype), _stream())
    return y


def btc_to_tbc(x, out_dtype=torch.float32):
    _chk(x, torch.float32, "x")
    B, T, Cc = x.shape
    y = torch.empty((T, B, Cc), dtype=out_dtype, device=x.device)
    call("csn_btc_to_tbc", _p(x), _p(y), B, T, Cc, _dt(out_dtype), _stream())
    return y


def cast(x, dtype):
    _chk(x, name="x")
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    call("csn_cast", _p(x), _dt(x.dtype), _p(y), _dt(dtype), x.numel(), _stream())
    return y


# ------------------------------------------------------------------------------------------------ dense
def gemm_f32(a, b, trans_a=False, trans_b=False, bias=None, act=ACT_NONE, out=None, alpha=1.0, beta=0.0, rowsum=None):
    """out[M,N] = alpha * op(a) @ op(b) + beta*out (+bias)(act).  op(a) [M,K], op(b) [K,N].
    rowsum: optional [M] tensor that receives the row sums of op(a) from the same pass (bias gradient of dW = dY^T X)."""
    _chk(a, torch.float32, "a"); _chk(b, torch.float32, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    K2, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != K2:
        raise _lib.CsnError("gemm_f32: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _chk(out, torch.float32, "out")
    if rowsum is not None:
        _chk(rowsum, torch.float32, "rowsum")
        if rowsum.numel() != M:
            raise _lib.CsnError("gemm_f32: rowsum must have %d elements" % M)
        call("csn_gemm_f32_rowsum", int(trans_a), int(trans_b), M, N, K, float(alpha), _p(a), a.shape[1], _p(b),
             b.shape[1], float(beta), _p(out), N, _p(bias), int(act), _p(rowsum), 0, _stream())
        return out
    call("csn_gemm_f32", int(trans_a), int(trans_b), M, N, K, float(alpha), _p(a), a.shape[1], _p(b), b.shape[1],
         float(beta), _p(out), N, _p(bias), int(act), _stream())
    return out


def gemm_bf16(a, b, trans_a=False, trans_b=False, bias=None, out=None, out_dtype=torch.float32, accumulate=False,
              split_k=1):
    """tcgen05 GEMM; a, b bfloat16.  Same op() convention as gemm_f32.  split_k > 1: partial slabs in a scratch
    tensor, added in split order by a second kernel (no atomics)."""
    _chk(a, torch.bfloat16, "a"); _chk(b, torch.bfloat16, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    K2, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != K2:
        raise _lib.CsnError("gemm_bf16: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _chk(out, name="out")
    ws = torch.empty((int(split_k), M, N), dtype=torch.float32, device=a.device) if split_k > 1 else None
    call("csn_gemm_bf16_tc", int(trans_a), int(trans_b), M, N, K, _p(a), a.shape[1], _p(b), b.shape[1], _p(out), N,
         _dt(out.dtype), _p(bias), int(accumulate), int(split_k), _p(ws), _stream())
    return out


def colsum(x, out=None, accumulate=False):
    _chk(x, torch.float32, "x")
    M, N = x.shape
    if out is None:
        out = torch.empty((N,), dtype=torch.float32, device=x.device)
    call("csn_colsum_f32", _p(x), _p(out), M, N, N, int(accumulate), _stream())
    return out


def scale_(x, scale_dev=None, scale_host=1.0):
    """x *= scale_dev[0] * scale_host in place (scale_dev: 0-d device tensor or None)."""
    _chk(x, torch.float32, "x")
    call("csn_scale_f32", _p(x), x.numel(), _p(scale_dev), float(scale_host), _stream())
    return x


def act_fwd(x, act):
    _chk(x, torch.float32, "x")
    y = torch.empty_like(x)
    call("csn_act_fwd", _p(x), _p(y), x.numel(), int(act), _stream())
    return y


def act_bwd(x, dy, act):
    _chk(x, torch.float32, "x"); _chk(dy, torch.float32, "dy")
    dx = torch.empty_like(x)
    call("csn_act_bwd", _p(x), _p(dy), _p(dx), x.numel(), int(act), _stream())
    return dx


def l2norm_fwd(x):
    _chk(x, torch.float32, "x")
    M, N = x.shape
    y = torch.empty_like(x)
    inv = torch.empty((M,), dtype=torch.float32, device=x.device)
    call("csn_l2norm_fwd", _p(x), _p(y), _p(inv), M, N, _stream())
    return y, inv


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, eps, momentum, training):
    """nn.BatchNorm1d on x [M, N] float32.  Returns (y, save_mean, save_rstd); the running statistics move in place."""
    _chk(x, torch.float32, "x")
    M, N = x.shape
    y = torch.empty_like(x)
    sm = torch.empty(N, dtype=torch.float32, device=x.device)
    sr = torch.empty(N, dtype=torch.float32, device=x.device)
    call("csn_batchnorm_fwd", _p(x), _p(gamma) if gamma is not None else None, _p(beta) if beta is not None else None,
         _p(running_mean) if running_mean is not None else None, _p(running_var) if running_var is not None else None,
         _p(y), _p(sm), _p(sr), M, N, float(eps), float(momentum), 1 if training else 0, _stream())
    return y, sm, sr


def batchnorm_bwd(x, dy, gamma, save_mean, save_rstd, training, need_dx=True, need_affine=True):
    _chk(x, torch.float32, "x"); _chk(dy, torch.float32, "dy")
    M, N = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dg = torch.empty(N, dtype=torch.float32, device=x.device) if need_affine else None
    db = torch.empty(N, dtype=torch.float32, device=x.device) if need_affine else None
    call("csn_batchnorm_bwd", _p(x), _p(dy), _p(gamma) if gamma is not None else None, _p(save_mean), _p(save_rstd),
         _p(dx) if dx is not None else None, _p(dg) if dg is not None else None, _p(db) if db is not None else None, M, N,
         1 if training else 0, _stream())
    return dx, dg, db


def l2norm_bwd(y, inv, dy):
    M, N = y.shape
    dx = torch.empty_like(y)
    call("csn_l2norm_bwd", _p(y), _p(inv), _p(_chk(dy, torch.float32, "dy")), _p(dx), M, N, _stream())
    return dx


def weight_norm_fwd(v, g):
    _chk(v, torch.float32, "v")
    N, K = v.shape
    w = torch.empty_like(v)
    inv = torch.empty((N,), dtype=torch.float32, device=v.device)
    call("csn_weight_norm_fwd", _p(v), _p(_chk(g.reshape(-1), torch.float32, "g")), _p(w), _p(inv), N, K, _stream())
    return w, inv


def weight_norm_bwd(v, g, inv, dw, need_dg=False):
    N, K = v.shape
    dv = torch.empty_like(v)
    dg = torch.empty((N,), dtype=torch.float32, device=v.device) if need_dg else None
    call("csn_weight_norm_bwd", _p(v), _p(g.reshape(-1)), _p(inv), _p(_chk(dw, torch.float32, "dw")), _p(dv), _p(dg),
         N, K, _stream())
    return dv, dg


# ------------------------------------------------------------------------------------------------ LSTM
def lstm_layer_bytes(T, B, I, H, compute_dtype):
    r, w = C.c_size_t(), C.c_size_t()
    call("csn_lstm_layer_bytes", T, B, I, H, _dt(compute_dtype), C.byref(r), C.byref(w))
    return r.value, w.value


def set_lstm_cta_budget(max_ctas):
    """Cap the SMs one persistent recurrence launch may occupy (0 = all): lets independent recurrences issued on
    different CUDA streams (crops of different length, the teacher pass) run side by side."""
    call("csn_lstm_set_cta_budget", int(max_ctas))


def lstm_layer_fwd(x, w_ih, w_hh, b_ih, b_hh, compute_dtype, training=True):
    """x [T,B,I] in compute_dtype -> (h_seq [T,B,H] compute_dtype, reserve, workspace)."""
    _chk(x, compute_dtype, "x")
    T, B, I = x.shape
    H = w_hh.shape[1]
    for n, w in (("w_ih", w_ih), ("w_hh", w_hh), ("b_ih", b_ih), ("b_hh", b_hh)):
        _chk(w, torch.float32, n)
    rb, wb = lstm_layer_bytes(T, B, I, H, compute_dtype)
    reserve = torch.empty((rb,), dtype=torch.uint8, device=x.device)
    workspace = torch.empty((wb,), dtype=torch.uint8, device=x.device)
    h_seq = torch.empty((T, B, H), dtype=compute_dtype, device=x.device)
    call("csn_lstm_layer_fwd", _p(x), _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh), _p(h_seq), _p(reserve), _p(workspace),
         T, B, I, H, _dt(compute_dtype), int(training), _stream())
    return h_seq, reserve, workspace


def lstm_layer_bwd(x, w_ih, w_hh, h_seq, reserve, workspace, d_hseq, d_hlast, grads, need_dx, compute_dtype,
                   accumulate=False):
    """grads = (dw_ih, dw_hh, db_ih, db_hh) fp32 tensors written (or accumulated into).  Returns dx or None."""
    T, B, I = x.shape
    H = w_hh.shape[1]
    dw_ih, dw_hh, db_ih, db_hh = grads
    dx = torch.empty((T, B, I), dtype=torch.float32, device=x.device) if need_dx else None
    call("csn_lstm_layer_bwd", _p(x), _p(w_ih), _p(w_hh), _p(h_seq), _p(reserve), _p(d_hseq), _p(d_hlast), _p(dw_ih),
         _p(dw_hh), _p(db_ih), _p(db_hh), _p(dx), _p(workspace), T, B, I, H, _dt(compute_dtype), int(accumulate),
         _stream())
    return dx


# ------------------------------------------------------------------------------------------------ loss / optimiser
def loss_workspace(rows, device):
    """Scratch for the fixed-order cross-CTA loss fold of the loss kernels (csn_loss_workspace_bytes)."""
    n = C.c_size_t()
    call("csn_loss_workspace_bytes", int(rows), C.byref(n))
    return torch.empty((n.value,), dtype=torch.uint8, device=device)


def dino_loss_fwd_bwd(student, teacher, center, student_temp, teacher_temp, mode, grad_scale=1.0, batch_center=None):
    """student [Vs,B,K] / teacher [Vt,B,K] fp32 (2-D inputs are treated as one view).
    Returns (loss scalar tensor, d_student, batch_center)."""
    s3 = student if student.dim() == 3 else student.unsqueeze(0)
    t3 = teacher if teacher.dim() == 3 else teacher.unsqueeze(0)
    _chk(s3, torch.float32, "student"); _chk(t3, torch.float32, "teacher"); _chk(center, torch.float32, "center")
    Vs, B, K = s3.shape
    Vt = t3.shape[0]
    if t3.shape[1:] != (B, K):
        raise _lib.CsnError("teacher shape %s does not match student %s" % (tuple(t3.shape), tuple(s3.shape)))
    center_rows = center.numel() // K
    loss = torch.empty((), dtype=torch.float32, device=s3.device)
    d_student = torch.empty_like(s3)
    if batch_center is None:
        n = B * K if mode == _lib.DINO_MULTICROP_REF else K
        batch_center = torch.zeros((n,), dtype=torch.float32, device=s3.device)
    call("csn_dino_loss_fwd_bwd", _p(s3), _p(t3), _p(center), center_rows, float(student_temp), float(teacher_temp),
         _p(loss), _p(d_student), _p(batch_center), Vs, Vt, B, K, int(mode), float(grad_scale),
         _p(loss_workspace(B, s3.device)), _stream())
    return loss, d_student.view(student.shape), batch_center


def head_dino_supported(B, I, K):
    """Whether csn_head_dino_fwd_bwd serves this head shape (else: gemm_f32 + dino_loss_fwd_bwd + gemm_f32)."""
    return bool(_lib.load().csn_head_dino_supported(int(B), int(I), int(K)))


def head_dino_fwd_bwd(h_last, w, bias, act, teacher, center, student_temp, teacher_temp, batch_center, grad_scale=1.0):
    """Projection head + single-view DINO loss, forward and backward in one kernel.  h_last [B, I] fp32 or bf16 (the
    recurrence's own output), w [K, I], teacher [B, K], center [K]; batch_center [K] is ACCUMULATED into (a fixed-order
    column sum of the teacher after the kernel), or None to skip it (run `colsum(teacher, out, accumulate=True)` wherever
    it overlaps best).  Returns (loss scalar tensor, d_hlast [B, I] fp32, d_pre [B, K] fp32)."""
    _chk(h_last, None, "h_last"); _chk(w, torch.float32, "w"); _chk(teacher, torch.float32, "teacher")
    _chk(center, torch.float32, "center")
    if batch_center is not None:
        _chk(batch_center, torch.float32, "batch_center")
    B, I = h_last.shape
    K = w.shape[0]
    if (w.shape[1] != I or tuple(teacher.shape) != (B, K) or center.numel() != K
            or (batch_center is not None and batch_center.numel() != K)):
        raise _lib.CsnError("head_dino_fwd_bwd: shapes do not match (h %s, w %s, teacher %s, center %d)"
                            % (tuple(h_last.shape), tuple(w.shape), tuple(teacher.shape), center.numel()))
    if not head_dino_supported(B, I, K):
        raise _lib.CsnError("head_dino_fwd_bwd: shape B=%d I=%d K=%d is not served by the fused head kernel" % (B, I, K))
    loss = torch.empty((), dtype=torch.float32, device=w.device)
    d_hlast = torch.empty((B, I), dtype=torch.float32, device=w.device)
    d_pre = torch.empty((B, K), dtype=torch.float32, device=w.device)
    call("csn_head_dino_fwd_bwd", _p(h_last), _dt(h_last.dtype), _p(w), _p(bias), int(act), _p(teacher), _p(center),
         float(student_temp), float(teacher_temp), _p(loss), _p(d_hlast), _p(d_pre), _p(batch_center), B, I, K,
         float(grad_scale), _p(loss_workspace(B, w.device)), _stream())
    return loss, d_hlast, d_pre


def center_ema(center, batch_center, momentum, scale):
    call("csn_center_ema", _p(center), _p(batch_center), center.numel(), float(momentum), float(scale), _stream())
    return center


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              decoupled=False, step=1, grad_scale=1.0):
    for n, t in (("params", params), ("grads", grads), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _chk(t, torch.float32, n)
    call("csn_adam_step", _p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(decoupled), int(step), float(grad_scale), _stream())


def adam_step_graph(params, grads, exp_avg, exp_avg_sq, step_counter, consts, lr, beta1=0.9, beta2=0.999, eps=1e-8,
                    weight_decay=0.0, decoupled=False, grad_scale=1.0):
    """Adam with the step count on the device (int32 tensor, incremented by the call): CUDA-graph replay safe."""
    call("csn_adam_step_graph", _p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), float(lr),
         float(beta1), float(beta2), float(eps), float(weight_decay), int(decoupled), _p(step_counter), _p(consts),
         float(grad_scale), _stream())


def dp_wait_done_zero(flags_local, world, step_counter, zero_tensor):
    """Wait (on the stream) until every peer has finished reading this rank's previous gradients, then zero
    `zero_tensor` (the centre-sum tail of the exchanged buffer)."""
    call("csn_dp_wait_done_zero", _p(flags_local), int(world), _p(step_counter), _p(zero_tensor),
         0 if zero_tensor is None else zero_tensor.numel(), _stream())


def dp_adam_step_peer(params, exp_avg, exp_avg_sq, grad_ptr_array, flag_ptr_array, world, rank, center,
                      center_momentum, center_scale, step_counter, ticket, lr, beta1=0.9, beta2=0.999, eps=1e-8,
                      weight_decay=0.0, decoupled=False, grad_scale=1.0):
    """Fused peer all-reduce + Adam + centre EMA (see include/csn_b200.h).  grad_ptr_array / flag_ptr_array are
    ctypes arrays of `world` device pointers (rank order)."""
    for n, t in (("params", params), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _chk(t, torch.float32, n)
    call("csn_dp_adam_step_peer", _p(params), _p(exp_avg), _p(exp_avg_sq), params.numel(), grad_ptr_array,
         flag_ptr_array, int(world), int(rank), _p(center), 0 if center is None else center.numel(),
         float(center_momentum), float(center_scale), _p(step_counter), _p(ticket), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(decoupled), float(grad_scale), _stream())


def ema_update_(dst, src, momentum):
    """dst = momentum * dst + (1 - momentum) * src, in place (flat fp32 buffers)."""
    _chk(dst, torch.float32, "dst"); _chk(src, torch.float32, "src")
    call("csn_ema_update", _p(dst), _p(src), dst.numel(), float(momentum), _stream())
    return dst


def dbg_umma_tile(a, b, a_mn=False, b_mn=False):
    """a [128,K], b [N,K] bf16 -> a @ b^T fp32 [128,N] through one tcgen05.mma chain (bring-up test hook)."""
    _chk(a, torch.bfloat16, "a"); _chk(b, torch.bfloat16, "b")
    N, K = b.shape
    d = torch.empty((128, N), dtype=torch.float32, device=a.device)
    call("csn_dbg_umma_tile", _p(a), _p(b), _p(d), N, K, int(a_mn), int(b_mn), _stream())
    return d
