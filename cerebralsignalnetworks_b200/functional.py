"""autograd bridges: each Function's forward/backward is a sequence of libcsn_b200 kernels (ops.py).
No torch arithmetic on the data path; torch only owns the tensors."""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU


# ---------------------------------------------------------------------------------------------------- encoder
def encoder_fwd(x_tbc, layer_weights, compute_dtype, training=True, cast_last=True):
    """Stacked LSTM over time-major input.  layer_weights: list of (w_ih, w_hh, b_ih, b_hh).
    Returns (h_last fp32 [B,H], saved) where saved is what encoder_bwd needs.  cast_last=False hands back the last
    hidden state in the recurrence's own dtype (a view of h_seq) for consumers that convert on load."""
    saved = []
    inp = x_tbc
    for (w_ih, w_hh, b_ih, b_hh) in layer_weights:
        h_seq, reserve, workspace = ops.lstm_layer_fwd(inp, w_ih, w_hh, b_ih, b_hh, compute_dtype, training)
        saved.append((inp, h_seq, reserve, workspace))
        inp = h_seq
    h_last = inp[-1]
    if not cast_last:
        return h_last.contiguous(), saved
    if h_last.dtype != torch.float32:
        h_last = ops.cast(h_last.contiguous(), torch.float32)
    else:
        h_last = h_last.contiguous()
    return h_last, saved


def encoder_bwd(d_hlast, layer_weights, saved, grads, compute_dtype, accumulate=False):
    """BPTT through the stack, top layer first.  grads: list of (dw_ih, dw_hh, db_ih, db_hh) to write."""
    d_hseq = None
    d_last = d_hlast.contiguous()
    for l in range(len(layer_weights) - 1, -1, -1):
        w_ih, w_hh, _, _ = layer_weights[l]
        inp, h_seq, reserve, workspace = saved[l]
        dx = ops.lstm_layer_bwd(inp, w_ih, w_hh, h_seq, reserve, workspace, d_hseq, d_last, grads[l],
                                need_dx=(l > 0), compute_dtype=compute_dtype, accumulate=accumulate)
        d_hseq, d_last = dx, None
    return None


class LSTMEncoderFunction(torch.autograd.Function):
    """x [T,B,I] (compute dtype) + flat list of per-layer weights -> h_last [B,H] fp32."""

    @staticmethod
    def forward(ctx, x_tbc, compute_dtype, training, *weights):
        layers = [tuple(weights[i:i + 4]) for i in range(0, len(weights), 4)]
        h_last, saved = encoder_fwd(x_tbc, layers, compute_dtype, training)
        ctx.layers, ctx.saved, ctx.compute_dtype = layers, saved, compute_dtype
        return h_last

    @staticmethod
    def backward(ctx, d_hlast):
        grads = [tuple(torch.empty_like(w) for w in layer) for layer in ctx.layers]
        encoder_bwd(d_hlast, ctx.layers, ctx.saved, grads, ctx.compute_dtype)
        flat = [g for layer in grads for g in layer]
        return (None, None, None) + tuple(flat)


# ---------------------------------------------------------------------------------------------------- dense
def linear_fwd(x, w, b, act=ACT_NONE, compute_dtype=torch.float32):
    """y = act(x @ w^T + b).  Returns (y, pre) where pre is the pre-activation (same tensor when act is NONE)."""
    if compute_dtype == torch.bfloat16:
        pre = ops.gemm_bf16(ops.cast(x, torch.bfloat16), ops.cast(w, torch.bfloat16), False, True, bias=b)
    else:
        pre = ops.gemm_f32(x, w, False, True, bias=b)
    y = pre if act == ACT_NONE else ops.act_fwd(pre, act)
    return y, pre


def linear_bwd(x, w, pre, dy, act=ACT_NONE, compute_dtype=torch.float32, need_dx=True, has_bias=True,
               dw_out=None, db_out=None):
    """Returns (dx, dw, db)."""
    dpre = dy.contiguous() if act == ACT_NONE else ops.act_bwd(pre, dy.contiguous(), act)
    if compute_dtype == torch.bfloat16:
        dpre_b = ops.cast(dpre, torch.bfloat16)
        x_b = ops.cast(x, torch.bfloat16)
        M = x.shape[0]
        dw = ops.gemm_bf16(dpre_b, x_b, True, False, out=dw_out, split_k=max(1, min(16, M // 256)))
        dx = None
        if need_dx:
            # dX = dY W contracts over the OUTPUT features: for the K = 65536 DINO head that is 1024 k-iterations on the
            # six 128 x 128 tiles of a [384, 256] result.  Split the contraction so the product fills the machine (partial
            # slabs, added in split order: deterministic).
            n_out, n_in = w.shape
            tiles = ((M + 127) // 128) * ((n_in + 127) // 128)
            split = max(1, min(32, 148 // max(1, tiles), n_out // 512))
            dx = ops.gemm_bf16(dpre_b, ops.cast(w, torch.bfloat16), False, False, split_k=split)
    else:
        # fp32: the bias gradient rides on the dW product (row sums of dY^T from the same shared-memory strip)
        db = None
        if has_bias:
            db = db_out if db_out is not None else torch.empty((dpre.shape[1],), dtype=torch.float32, device=dpre.device)
        dw = ops.gemm_f32(dpre, x, True, False, out=dw_out, rowsum=db)
        dx = ops.gemm_f32(dpre, w, False, False) if need_dx else None
        return dx, dw, db
    db = ops.colsum(dpre, out=db_out) if has_bias else None
    return dx, dw, db


class LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act, compute_dtype):
        x = x.contiguous()
        y, pre = linear_fwd(x, w, b, act, compute_dtype)
        ctx.save_for_backward(x, w, pre)
        ctx.act, ctx.compute_dtype, ctx.has_bias = act, compute_dtype, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, pre = ctx.saved_tensors
        dx, dw, db = linear_bwd(x, w, pre, dy, ctx.act, ctx.compute_dtype, need_dx=ctx.needs_input_grad[0],
                                has_bias=ctx.has_bias)
        return dx, dw, db, None, None


class ActFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.act = act
        return ops.act_fwd(x, act)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.act_bwd(x, dy.contiguous(), ctx.act), None


class BatchNormFunction(torch.autograd.Function):
    """nn.BatchNorm1d on [M, N] -- the use_bn layers of DINOHead (LstmDistillation.py:72-80)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, training):
        x = x.contiguous()
        y, sm, sr = ops.batchnorm_fwd(x, gamma, beta, running_mean, running_var, eps, momentum, training)
        ctx.save_for_backward(x, gamma, sm, sr)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, sm, sr = ctx.saved_tensors
        dx, dg, db = ops.batchnorm_bwd(x, dy.contiguous(), gamma, sm, sr, ctx.training, need_dx=ctx.needs_input_grad[0],
                                       need_affine=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        return dx, dg, db, None, None, None, None, None


class L2NormFunction(torch.autograd.Function):
    """F.normalize(x, dim=-1, p=2) -- LstmDistillation.py:97."""

    @staticmethod
    def forward(ctx, x):
        y, inv = ops.l2norm_fwd(x.contiguous())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(y, inv, dy.contiguous())


class WeightNormFunction(torch.autograd.Function):
    """w = g * v / ||v||_row -- nn.utils.weight_norm(dim=0), LstmDistillation.py:86."""

    @staticmethod
    def forward(ctx, v, g):
        w, inv = ops.weight_norm_fwd(v.contiguous(), g.contiguous())
        ctx.save_for_backward(v, g, inv)
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g, inv = ctx.saved_tensors
        dv, dg = ops.weight_norm_bwd(v, g, inv, dw.contiguous(), need_dg=ctx.needs_input_grad[1])
        if dg is not None:
            dg = dg.view(g.shape)
        return dv, dg


# ---------------------------------------------------------------------------------------------------- loss
class DINOLossFunction(torch.autograd.Function):
    """Loss and dLoss/dstudent come out of ONE kernel; backward just scales by the incoming gradient."""

    @staticmethod
    def forward(ctx, student, teacher, center, student_temp, teacher_temp, mode, stats_out, batch_center=None):
        loss, d_student, bc = ops.dino_loss_fwd_bwd(student.contiguous(), teacher.contiguous(), center.contiguous(),
                                                    student_temp, teacher_temp, mode, batch_center=batch_center)
        stats_out.append(bc)
        ctx.save_for_backward(d_student)
        ctx.needs_clone = False  # single backward pass per forward (retain_graph callers set this)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (d_student,) = ctx.saved_tensors
        # d_student is owned by this node; scale in place with the (scalar) upstream gradient
        g = d_student.clone() if ctx.needs_clone else d_student
        ops.scale_(g, dloss.contiguous().to(torch.float32))
        return g, None, None, None, None, None, None, None
