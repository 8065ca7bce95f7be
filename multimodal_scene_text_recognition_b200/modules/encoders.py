"""Sequence encoder behind the reference's module interface (modules/encoders.py:10-30).

``BidirectionalLSTM`` keeps ``self.rnn = nn.LSTM(...)`` and ``self.linear = nn.Linear(...)`` as
parameter containers, so the state-dict keys (``rnn.weight_ih_l0[_reverse]`` ...) are the
reference's; neither container's own forward is ever called.  The work is three launches:

1. input projection of both directions  x [W_ih_fwd; W_ih_bwd]^T + (b_ih + b_hh)  - GEMM
   (``mstr_conv_gemm_bf16`` on tcgen05, or ``mstr_conv_gemm_f32`` in parity mode);
2. ``mstr_bilstm_recurrent_bf16`` (bf16 mode: h_t W_hh^T on tcgen05, c_t in TMEM) or
   ``mstr_bilstm_recurrent`` (fp32 parity mode, SIMT) - ONE persistent launch for all T steps of both
   directions;
3. Linear(2H -> out) + bias - GEMM.

The transformer encoder of the reference (encoders.py:68-243) is mirrored in modules/transformer.py and re-exported
here, like ``Oscar_Bert`` (encoders.py:34-66, a BERT-base encoder).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from .._pack import act_dtype, cached, require_inference, wrap_weight


class BidirectionalLSTM(nn.Module):

    def __init__(self, input_size, hidden_size, output_size):
        super(BidirectionalLSTM, self).__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, bidirectional=True, batch_first=True)
        self.linear = nn.Linear(hidden_size * 2, output_size)

    def _packed(self, precision: str):
        r = self.rnn
        src = [r.weight_ih_l0, r.weight_ih_l0_reverse, r.weight_hh_l0, r.weight_hh_l0_reverse,
               r.bias_ih_l0, r.bias_hh_l0, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse,
               self.linear.weight, self.linear.bias]

        def build():
            dt = act_dtype(precision)
            w_ih = torch.cat([r.weight_ih_l0, r.weight_ih_l0_reverse], 0).detach().float()      # [2*4H, in]
            bias = torch.cat([r.bias_ih_l0 + r.bias_hh_l0,
                              r.bias_ih_l0_reverse + r.bias_hh_l0_reverse], 0).detach().float()  # [2*4H]
            w_hh = torch.stack([r.weight_hh_l0, r.weight_hh_l0_reverse], 0).detach().float()     # [2, 4H, H]
            return {"w_ih": wrap_weight(w_ih, precision), "bias": bias.contiguous(), "w_hh": w_hh.contiguous(),
                    "w_hh_tc": ops.pack_w_hh_bf16(w_hh) if precision == 'bf16' else None,
                    "w_lin": wrap_weight(self.linear.weight.detach().float(), precision),
                    "b_lin": self.linear.bias.detach().float().contiguous()}
        return cached(self, ("bilstm", precision), src, build)

    def forward_features(self, col_feats, want_f32: bool = True, want_act: bool = False):
        """col_feats [B, T, in] (f32, or already in the activation dtype) -> contextual feature
        [B, T, out]: f32 and / or activation-dtype copies."""
        require_inference(col_feats)
        precision = config.PRECISION
        dt = act_dtype(precision)
        pk = self._packed(precision)
        x = ops.cast(col_feats.contiguous(), dt)
        B, T, _ = x.shape
        hd = self.rnn.hidden_size
        if precision == 'bf16':
            # the input projection of both directions; with GATES_X_BF16 it is stored (and read back by the recurrence)
            # as bf16: half the HBM round trip of the stage
            half = bool(config.RNN_TENSOR_CORES and config.GATES_X_BF16)
            gx = ops.linear(x, pk["w_ih"], pk["bias"], want_f32=not half, want_bf16=half)    # [B,T,2*4H]
            if config.RNN_TENSOR_CORES:
                rec = ops.bilstm_recurrent_bf16(gx.view(B, T, 2, 4 * hd), pk["w_hh_tc"])
            else:
                rec = ops.bilstm_recurrent(gx.view(B, T, 2, 4 * hd), pk["w_hh"], want_f32=False, want_bf16=True)
            return ops.linear(rec, pk["w_lin"], pk["b_lin"], want_f32=want_f32, want_bf16=want_act)
        gx = ops.linear(x, pk["w_ih"], pk["bias"])
        rec = ops.bilstm_recurrent(gx.view(B, T, 2, 4 * hd), pk["w_hh"], want_f32=True, want_bf16=False)
        out = ops.linear(rec, pk["w_lin"], pk["b_lin"])
        return (out, out) if (want_f32 and want_act) else out

    def forward(self, col_feats):
        """visual feature [B, T, input_size] -> contextual feature [B, T, output_size] f32
        (encoders.py:17-30)."""
        return self.forward_features(col_feats, want_f32=True, want_act=False)


# the transformer / BERT encoder families live in modules/transformer.py; the reference defines these names in this module
from .transformer import TF_Encoder, Oscar_Bert, PositionalEncoding, TransformerEncoderLayer, MLP  # noqa: E402,F401
