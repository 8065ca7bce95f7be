"""Attention decoder behind the reference's module interface (modules/decoders.py:10-101).

``Attention`` / ``AttentionCell`` keep ``i2h``, ``h2h``, ``score``, ``rnn = nn.LSTMCell`` and
``generator`` as parameter containers (same state-dict keys).  ``Attention.forward`` is two
launches: the loop-invariant ``i2h(batch_H)`` GEMM (the reference recomputes it every step,
decoders.py:81) and ONE persistent ``mstr_attn_decode`` launch that runs all
``config.MAX_TEXT_LENGTH + 1`` steps - Bahdanau attention, LSTMCell, generator, argmax feedback -
with optional early exit on ``[s]``.

``semantics`` is accepted and ignored, as in the reference (AttentionCell.forward never reads
it, decoders.py:79-101).

The transformer / linear decoders (decoders.py:103-516) are mirrored in modules/transformer.py and re-exported
here.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from .._pack import act_dtype, cached, require_inference, wrap_weight


def _cell_pack(cell: "AttentionCell", generator: nn.Linear, precision: str):
    hd = cell.hidden_size
    src = [cell.i2h.weight, cell.h2h.weight, cell.h2h.bias, cell.score.weight, cell.rnn.weight_ih,
           cell.rnn.weight_hh, cell.rnn.bias_ih, cell.rnn.bias_hh, generator.weight, generator.bias]

    def build():
        f = lambda t: t.detach().float().contiguous()
        w_ih = cell.rnn.weight_ih.detach().float()
        pk = {
            "w_i2h": wrap_weight(f(cell.i2h.weight), precision),
            "w_h2h": f(cell.h2h.weight), "b_h2h": f(cell.h2h.bias),
            "w_score": f(cell.score.weight).view(-1).contiguous(),
            "w_ctx": w_ih[:, :hd].contiguous(),                 # context part of LSTMCell.weight_ih
            "w_tok": w_ih[:, hd:].t().contiguous(),             # one-hot part, transposed: a row gather
            "w_hh": f(cell.rnn.weight_hh),
            "bias": (cell.rnn.bias_ih.detach().float() + cell.rnn.bias_hh.detach().float()).contiguous(),
            "w_gen": f(generator.weight), "b_gen": f(generator.bias),
        }
        if precision == 'bf16':   # operands of the tensor-core decode (mstr_attn_decode_bf16)
            head = torch.zeros((384, hd), device=w_ih.device, dtype=torch.float32)
            head[:hd] = cell.h2h.weight.detach().float()
            head[hd:hd + generator.weight.shape[0]] = generator.weight.detach().float()
            pk["w_head"] = head.to(torch.bfloat16).contiguous()
            pk["w_cat"] = ops.pack_gate_rows_bf16(torch.cat([w_ih[:, :hd], cell.rnn.weight_hh.detach().float()], 1))
        return pk
    return cached(cell, ("attn", precision, id(generator)), src, build)


def _project(batch_H: torch.Tensor, pk: dict, precision: str) -> torch.Tensor:
    """batch_H_proj = i2h(batch_H) (decoders.py:81), f32 [B, T, H]."""
    if precision == 'bf16':
        return ops.linear(ops.cast(batch_H, torch.bfloat16), pk["w_i2h"], None, want_f32=True, want_bf16=False)
    return ops.linear(batch_H, pk["w_i2h"], None)


class Attention(nn.Module):

    def __init__(self, input_size, hidden_size, num_classes):
        super(Attention, self).__init__()
        self.attention_cell = AttentionCell(input_size, hidden_size, num_classes)
        self.hidden_size = hidden_size
        self.num_classes = num_classes
        self.generator = nn.Linear(hidden_size, num_classes)

    def _char_to_onehot(self, input_char, onehot_dim=38):
        """decoders.py:19-24; kept for callers, not used by the fused kernel (a one-hot times
        W_ih is a row gather)."""
        input_char = input_char.unsqueeze(1)
        one_hot = torch.zeros(input_char.size(0), onehot_dim, device=input_char.device)
        return one_hot.scatter_(1, input_char, 1)

    def decode(self, encoder_output, text, is_train, enc_act=None, want_alpha: bool = False):
        """Full result of the fused decode: dict(logits, ids, lengths[, alpha])."""
        require_inference(encoder_output)
        precision = config.PRECISION
        num_steps = config.MAX_TEXT_LENGTH + 1          # decoders.py:34, read at call time
        pk = _cell_pack(self.attention_cell, self.generator, precision)
        tokens = None
        if is_train:
            if text is None:
                raise ValueError("is_train=True needs the teacher-forcing text (decoders.py:41-50)")
            tokens = text.to(device=encoder_output.device, dtype=torch.int64).contiguous()
            if config.CHECK_INPUTS and tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) >= self.num_classes):
                raise IndexError("teacher-forcing text holds a token outside [0, %d) (decoders.py:23 scatter_ would raise)"
                                 % self.num_classes)   # costs a device->host sync: off on the hot path (the kernels map it to 0)
        if (precision == 'bf16' and config.RNN_TENSOR_CORES and self.hidden_size == 256
                and encoder_output.shape[1] <= 96 and self.num_classes <= 128):
            enc_b = enc_act if enc_act is not None else ops.cast(encoder_output.contiguous(), torch.bfloat16)
            proj_b = ops.linear(enc_b, pk["w_i2h"], None, want_f32=False, want_bf16=True)
            return ops.attn_decode_bf16(enc_b, proj_b, pk, num_steps, tokens,
                                        early_exit=(not is_train) and bool(config.DECODER_EARLY_EXIT),
                                        want_alpha=want_alpha)
        enc = ops.cast(encoder_output.contiguous(), torch.float32)
        if precision == 'bf16' and enc_act is not None:
            proj = ops.linear(enc_act, pk["w_i2h"], None, want_f32=True, want_bf16=False)
        else:
            proj = _project(enc, pk, precision)
        if is_train:
            return ops.attn_decode(enc, proj, pk, num_steps, tokens, early_exit=False, want_alpha=want_alpha)
        return ops.attn_decode(enc, proj, pk, num_steps, None, early_exit=bool(config.DECODER_EARLY_EXIT),
                               want_alpha=want_alpha)

    def forward(self, encoder_output, text, semantics, is_train):
        """encoder_output [B, T, C] -> logits [B, num_steps, num_classes] (decoders.py:26-64)."""
        return self.decode(encoder_output, text, is_train)["logits"]


class AttentionCell(nn.Module):

    def __init__(self, input_size, hidden_size, num_embeddings):
        super(AttentionCell, self).__init__()
        self.i2h = nn.Linear(input_size, hidden_size, bias=False)
        self.h2h = nn.Linear(hidden_size, hidden_size)
        self.score = nn.Linear(hidden_size, 1, bias=False)
        self.rnn = nn.LSTMCell(input_size + num_embeddings, hidden_size)
        self.hidden_size = hidden_size
        self.num_embeddings = num_embeddings

    def forward(self, prev_hidden, batch_H, char_onehots, semantics):
        """One decoder step (decoders.py:79-101) -> ((h, c), alpha [B, T, 1]).

        Runs the fused kernel for a single step; ``char_onehots`` must be one-hot rows, which is
        all the reference ever passes (decoders.py:43,57)."""
        require_inference(batch_H)
        precision = config.PRECISION
        enc = ops.cast(batch_H.contiguous(), torch.float32)
        gen = self.__dict__.get("_step_generator")
        if gen is None:   # the step API has no generator; logits of a zero generator are discarded
            gen = nn.Linear(self.hidden_size, self.num_embeddings).to(enc.device)
            object.__setattr__(self, "_step_generator", gen)
        pk = _cell_pack(self, gen, precision)
        proj = _project(enc, pk, precision)
        tokens = char_onehots.argmax(dim=1, keepdim=True).to(torch.int64).contiguous()
        res = ops.attn_decode(enc, proj, pk, 1, tokens, early_exit=False,
                              h0=prev_hidden[0].contiguous().float(), c0=prev_hidden[1].contiguous().float(),
                              want_alpha=True, want_state=True)
        return (res["h"], res["c"]), res["alpha"].permute(0, 2, 1)


# the transformer / linear decoders live in modules/transformer.py; the reference defines these names in this module
from .transformer import (TF_Decoder, Linear_Decoder, TransformerDecoder, TransformerDecoderLayer,  # noqa: E402,F401
                          PositionalEncoding, MLP)
