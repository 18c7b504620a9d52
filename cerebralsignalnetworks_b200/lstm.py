"""Drop-in `models.lstm.Model` -- the LSTM EEG encoder.

The reference imports `from models.lstm import Model` (LstmDistillFromDinoV2Train.py:5,
LstmDistillation.py:5, LstmDistillFromDinoV2TrainSpampinato.py:5, LstmDistillFromDinoV2Eval.py:5) but does not
ship the file; the interface below follows its call sites:
    Model(input_size=96, lstm_size=96, lstm_layers=2, output_size=features_length, include_top=True)   (Train.py:323)
    Model(128, 128, 4, features_length, include_top=False)                                             (Spampinato.py:368)
    lstm_output, cls_pred = model(eeg)  with eeg [B, T, C] float32                                     (Train.py:365)
    MultiCropWrapper assigns .fc / .head = nn.Identity()                                               (LstmDistillation.py:40)
    state_dict keys load after stripping "backbone."                                                   (Eval.py:309-313)
and the body of the in-tree analogues (LSTMDistillRetreival.py:85-110, LSTMDistill.py:112-142).
The arithmetic runs in libcsn_b200 (persistent tcgen05 recurrence in bf16 mode, SIMT fp32 in parity mode).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ACT_NONE, ACT_RELU
from .functional import ActFunction, LinearFunction, LSTMEncoderFunction


class LSTMStack(nn.Module):
    """Parameter holder with torch.nn.LSTM's names (weight_ih_l{k}, weight_hh_l{k}, bias_ih_l{k}, bias_hh_l{k})
    and its default initialisation (drawn through a throw-away nn.LSTM so a seeded construction yields
    bit-identical initial weights to the torch reference)."""

    def __init__(self, input_size, hidden_size, num_layers):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        proto = nn.LSTM(input_size, hidden_size, num_layers=num_layers, batch_first=True)
        for name, p in proto.named_parameters():
            self.register_parameter(name, nn.Parameter(p.detach().clone()))

    def layer_weights(self):
        return [(getattr(self, f"weight_ih_l{k}"), getattr(self, f"weight_hh_l{k}"),
                 getattr(self, f"bias_ih_l{k}"), getattr(self, f"bias_hh_l{k}")) for k in range(self.num_layers)]


class Linear(nn.Module):
    """nn.Linear-compatible parameters (weight [out, in], bias [out]); forward on libcsn_b200 GEMMs."""

    def __init__(self, in_features, out_features, bias=True, compute_dtype=torch.float32):
        super().__init__()
        proto = nn.Linear(in_features, out_features, bias=bias)
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(proto.weight.detach().clone())
        self.bias = nn.Parameter(proto.bias.detach().clone()) if bias else None
        self.compute_dtype = compute_dtype

    def forward(self, x, act=ACT_NONE):
        return LinearFunction.apply(x, self.weight, self.bias, act, self.compute_dtype)


class Model(nn.Module):
    def __init__(self, input_size=128, lstm_size=128, lstm_layers=1, output_size=128, include_top=True,
                 n_classes=40, compute_dtype=torch.bfloat16):
        super().__init__()
        self.input_size, self.lstm_size, self.lstm_layers = input_size, lstm_size, lstm_layers
        self.output_size, self.include_top = output_size, include_top
        self.compute_dtype = compute_dtype
        self.lstm = LSTMStack(input_size, lstm_size, lstm_layers)
        self.output = Linear(lstm_size, output_size)
        if include_top:
            self.classifier = Linear(output_size, n_classes)
        self.fc = nn.Identity()
        self.head = nn.Identity()

    # -- time-major entry: what the fused band-pass kernel emits ([T, B, C] in compute dtype) --
    def forward_time_major(self, x_tbc):
        flat = [w for layer in self.lstm.layer_weights() for w in layer]
        h_last = LSTMEncoderFunction.apply(x_tbc, self.compute_dtype, self.training or torch.is_grad_enabled(), *flat)
        feat = self.output(h_last)
        if self.include_top:
            cls = self.classifier(feat)
            return ActFunction.apply(feat, ACT_RELU), cls
        return feat

    @torch.no_grad()
    def encode_trials(self, eeg_bct, sos=None, zero_phase=False):
        """Inference from RAW stored-layout trials [B, C, T] float32 (BASELINE.json config 5: 63-channel, 2000-sample
        Perils trials): fused band-pass -> [T, B, C] -> LSTM stack (no BPTT reserve) -> projection.  Returns the
        embeddings [B, output_size] (the features LstmDistillFromDinoV2Eval.py:333-380 hands to the faiss search).
        The bf16 tensor-core path needs 16-byte input rows: a channel count that is not a multiple of 8 is padded with
        zero channels (and W_ih with zero columns) -- zero inputs through zero weights, the result is unchanged."""
        from .functional import encoder_fwd, linear_fwd
        _lib.require_gpu()
        B, Cc, T = eeg_bct.shape
        if Cc != self.input_size:
            raise ValueError("Model expects [B, %d, T] trials, got %s" % (self.input_size, tuple(eeg_bct.shape)))
        x = eeg_bct.contiguous().float()
        layers = self.lstm.layer_weights()
        pad = (-Cc) % 8 if self.compute_dtype == torch.bfloat16 else 0
        if pad:
            x = torch.nn.functional.pad(x, (0, 0, 0, pad))  # [B, C + pad, T], zero channels
            w_ih0 = torch.nn.functional.pad(layers[0][0].detach(), (0, pad))
            layers = [(w_ih0,) + tuple(layers[0][1:])] + layers[1:]
        if sos is not None:
            x_tbc = ops.sosfilt(x, sos, zero_phase=zero_phase, out_layout="TBC", out_dtype=self.compute_dtype)
        else:
            x_tbc = ops.btc_to_tbc(x.transpose(1, 2).contiguous(), self.compute_dtype)
        h_last, _ = encoder_fwd(x_tbc, [tuple(w.detach() for w in l) for l in layers], self.compute_dtype, training=False)
        feat, _ = linear_fwd(h_last, self.output.weight.detach(), self.output.bias.detach(), ACT_NONE)
        return feat

    def forward(self, x):
        """x: float32 [B, T, C] (batch_first, as the reference DataLoader yields it)."""
        if x.dim() != 3 or x.shape[-1] != self.input_size:
            raise ValueError("Model expects [B, T, %d], got %s" % (self.input_size, tuple(x.shape)))
        _lib.require_gpu()
        x = x.contiguous().float()
        return self.forward_time_major(ops.btc_to_tbc(x, self.compute_dtype))
