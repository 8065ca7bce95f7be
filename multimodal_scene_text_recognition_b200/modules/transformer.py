"""Transformer encoder / decoder family behind the reference's module interface.

Mirrors ``TF_Encoder`` (modules/encoders.py:68-143), ``TF_Decoder`` (modules/decoders.py:103-275) and
``Linear_Decoder`` (:278-296) with every fusion switch of ``config`` at its default (False): the reference's own
default configuration, ``ENCODER = DECODER = 'Transformer'`` (config.py:27-28).  The helper classes keep the
reference's names, constructor arguments, attribute names and construction order, so the state-dict keys
(``encoder.encoder_layer.*``, ``encoder.encoder.layers.{i}.*``, ``decoder.decoder.layers.{i}.self_attn.in_proj_weight``
...) and the weights a given seed produces are the reference's.  ``nn.MultiheadAttention`` / ``nn.Linear`` /
``nn.LayerNorm`` / ``nn.Embedding`` are parameter containers only: their own forward is never called.

Work per layer: the GEMMs (in_proj, out_proj, linear1 + ReLU, linear2) run on the conv-GEMM kernels with bias, ReLU
and the residual in the epilogue (``ops.linear``); between them ``ops.mha`` / ``ops.mha_step`` (attention core),
``ops.layer_norm``, ``ops.add_pos_encoding`` / ``ops.embed_tokens`` and ``ops.argmax_tokens`` (csrc/transformer.cu).

Two things differ from the reference's code, neither in the numbers it defines:
* the greedy loop keeps a K/V cache.  The reference re-runs the decoder over positions 0..t at every step
  (decoders.py:250-279); under its causal mask the rows of earlier positions do not change, so the cache yields the
  same ``t_output[t]``;
* ``nn.TransformerEncoder.forward`` of the installed torch (2.11) passes ``is_causal=`` to the layer, which the
  reference's own ``TransformerEncoderLayer.forward`` (encoders.py:185) does not accept: the reference's encoder
  raises TypeError here.  What it computes under the torch it was written for is layers in order, then ``norm`` -
  that is what ``TF_Encoder`` does (and what the golden fixtures drive the reference's layer objects with).

``Oscar_Bert`` (encoders.py:34-66) is mirrored too (BERT-base over the column features).

Of the semantic-fusion switches (config.py:40-53) ``PRE_ENCODER_MLP`` and ``PRE_DECODER_MLP`` are built - they are what
makes the object-tag vectors of ``Linear_Embedding`` reach the prediction (``relevant_semantic`` / ``combine_with_semantic``
below) -; ``CLS_DECODER_INIT`` and ``POST_DECODER_MLP`` too.  ``OSCAR_ENCODER`` and the ``MULTIHEAD_*`` switches raise
NotImplementedError at construction: they raise in the reference itself (tensor-shape errors in its forward).  Eval mode only (dropout is the identity).
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import config, ops
from .._pack import act_dtype, cached, require_inference, wrap_weight


def _no_fusion(*flags):
    on = [f for f in flags if getattr(config, f, False)]
    if on:
        raise NotImplementedError("config.%s: the semantic-fusion experiments of the reference are not implemented "
                                  "(SURVEY.md 8 f4 covers the default configuration)" % ", config.".join(on))


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


class PositionalEncoding(nn.Module):
    """encoders.py:146-165 / decoders.py:300-317: sinusoidal table registered as the buffer ``pe`` [max_len, 1, d]."""

    def __init__(self, d_model, dropout=0.1, max_len=26):
        super(PositionalEncoding, self).__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0).transpose(0, 1))

    def table(self) -> torch.Tensor:
        """[max_len, d] f32 view of the buffer."""
        return self.pe.detach().float().view(self.pe.shape[0], self.pe.shape[2])

    def forward(self, x):
        """x [T, B, d] (the reference's layout) -> x + pe[:T]."""
        require_inference(x)
        xb = x.permute(1, 0, 2).contiguous()
        return ops.add_pos_encoding(xb, self.table().contiguous(), xb.dtype).permute(1, 0, 2)


class MLP(nn.Module):
    """encoders.py:200-243 / decoders.py:473-516: ``num_layers`` Linear layers (ReLU and a p = 0 dropout between them)
    registered as ``layers.fc{i}``, weights N(0, sqrt(2 / fan_in)), biases 0.  A parameter container: the fusion code
    below runs the layers on the GEMM kernels."""

    def __init__(self, input_size, hidden_size, num_classes, num_layers=1, dropout_p=0.0):
        super(MLP, self).__init__()
        self.input_size, self.hidden_size, self.num_classes = input_size, hidden_size, num_classes
        layers = []
        for i in range(num_layers):
            idim = input_size if i == 0 else hidden_size
            odim = num_classes if i == num_layers - 1 else hidden_size
            fc = nn.Linear(idim, odim)
            fc.weight.data.normal_(0.0, math.sqrt(2. / idim))
            fc.bias.data.fill_(0)
            layers.append(('fc' + str(i), fc))
            if i != num_layers - 1:
                layers.append(('relu' + str(i), nn.ReLU()))
                layers.append(('dropout' + str(i), nn.Dropout(p=dropout_p)))
        self.layers = nn.Sequential(OrderedDict(layers))

    def params_to_train(self):
        return self.layers.parameters()

    def linears(self):
        return [m for m in self.layers if isinstance(m, nn.Linear)]

    def forward(self, x):
        """x [..., input_size] -> [..., num_classes] on the GEMM kernels (activation dtype of config.PRECISION)."""
        require_inference(x)
        precision = config.PRECISION
        h = ops.cast(x.contiguous(), act_dtype(precision))
        lins = self.linears()
        for i, fc in enumerate(lins):
            w, b = cached(fc, ("mlp_fc", precision), [fc.weight, fc.bias],
                          lambda fc=fc: (wrap_weight(_f(fc.weight), precision), _f(fc.bias)))
            h = ops.linear(h, w, b, relu=i != len(lins) - 1)
        return h


def _pad_w(w: torch.Tensor, pad: int) -> torch.Tensor:
    """Weight [N, K] zero-padded to multiples of ``pad`` in both dimensions (pad = 1: unchanged)."""
    n, k = w.shape
    N, K = -(-n // pad) * pad, -(-k // pad) * pad
    if (N, K) == (n, k):
        return w.contiguous()
    out = torch.zeros((N, K), device=w.device, dtype=w.dtype)
    out[:n, :k] = w
    return out


def _pad_v(v: torch.Tensor, pad: int) -> torch.Tensor:
    n = v.shape[0]
    N = -(-n // pad) * pad
    if N == n:
        return v.contiguous()
    out = torch.zeros((N,), device=v.device, dtype=v.dtype)
    out[:n] = v
    return out


def _relevance_pack(mlp: MLP, feat_dim: int, precision: str, pad: int = 1):
    """The 3-layer relevance MLP on cat(column feature [feat_dim], object vector): first layer split into its column
    and its object half, last layer (-> 1 score) as a vector.  ``pad``: every dimension zero-padded to a multiple of it
    (the 97-wide layers of POST_DECODER_MLP on GEMM tiles of 128; padded channels stay exactly zero through the ReLUs)."""
    fc0, fc1, fc2 = mlp.linears()
    if fc2.out_features != 1:
        raise NotImplementedError("the relevance MLP must end in one score")

    def build():
        w0 = _f(fc0.weight)
        return {"w0_feat": wrap_weight(_pad_w(w0[:, :feat_dim], pad), precision), "b0": _pad_v(_f(fc0.bias), pad),
                "w0_sem": wrap_weight(_pad_w(w0[:, feat_dim:], pad), precision),
                "w1": wrap_weight(_pad_w(_f(fc1.weight), pad), precision), "b1": _pad_v(_f(fc1.bias), pad),
                "w2": _pad_v(_f(fc2.weight).view(-1), pad), "b2": float(fc2.bias.detach().float().item())}
    return cached(mlp, ("relevance", precision, feat_dim, pad), [p for fc in (fc0, fc1, fc2) for p in (fc.weight, fc.bias)], build)


def relevant_semantic(feats: torch.Tensor, sem_vec: torch.Tensor, mlp: MLP, feat_dim: int = 0, pad: int = 1) -> torch.Tensor:
    """get_relevant_semantic (encoders.py:89-111, decoders.py:147-170): feats [B, T, F] (activation dtype), sem_vec
    [B, O, E] f32 -> [B, T, E]: for every (column, object) pair ``mlp(cat(column, object))`` is a score, soft-maxed over
    the objects, weighting the sum of the object vectors.  The reference materialises cat(...) for all T x O pairs; the
    first layer is linear, so its column half runs once per column and its object half once per object (two GEMMs) and
    ``pair_add_relu`` forms the T x O hidden rows; the middle layer is one GEMM over them, the last layer + softmax +
    weighted sum one kernel.  ``feat_dim`` / ``pad``: the un-padded feature width when feats / sem_vec arrive zero-padded."""
    precision = config.PRECISION
    dt = act_dtype(precision)
    pk = _relevance_pack(mlp, feat_dim or feats.shape[-1], precision, pad)
    sem32 = sem_vec.contiguous().float()
    a = ops.linear(feats, pk["w0_feat"], pk["b0"])                                  # [B, T, Hd]
    s_ = ops.linear(ops.cast(sem32, dt), pk["w0_sem"], None)                        # [B, O, Hd]
    h = ops.pair_add_relu(a, s_)                                                    # [B, T, O, Hd]
    h = ops.linear(h, pk["w1"], pk["b1"], relu=True)
    return ops.relevance_reduce(h, pk["w2"], pk["b2"], sem32)                       # [B, T, E]


def _combine_pack(mlp: MLP, feat_dim: int, precision: str, pad: int = 1):
    lins = mlp.linears()

    def build():
        w0 = _f(lins[0].weight)
        rest = [(wrap_weight(_pad_w(_f(fc.weight), pad), precision), _pad_v(_f(fc.bias), pad)) for fc in lins[1:]]
        return {"w0_feat": wrap_weight(_pad_w(w0[:, :feat_dim], pad), precision), "b0": _pad_v(_f(lins[0].bias), pad),
                "w0_sem": wrap_weight(_pad_w(w0[:, feat_dim:], pad), precision), "rest": rest}
    return cached(mlp, ("combine", precision, feat_dim, pad), [p for fc in lins for p in (fc.weight, fc.bias)], build)


def combine_with_semantic(feats: torch.Tensor, rel: torch.Tensor, mlp: MLP, feat_dim: int = 0, pad: int = 1,
                          want_f32: bool = False) -> torch.Tensor:
    """feats + combine_mlp(cat(feats, rel)) (encoders.py:131-134, decoders.py:220-222, :289-295) without materialising
    the concatenation: the first layer as two GEMMs (the second adds the first in its epilogue), the residual ``feats`` in
    the epilogue of the last (``want_f32``: that last result in f32 whatever the activation dtype)."""
    pk = _combine_pack(mlp, feat_dim or feats.shape[-1], config.PRECISION, pad)
    h = ops.linear(feats, pk["w0_feat"], pk["b0"])
    h = ops.linear(rel, pk["w0_sem"], None, residual=h, relu=True)
    for i, (w, b) in enumerate(pk["rest"]):
        last = i == len(pk["rest"]) - 1
        if last and want_f32 and config.PRECISION == 'bf16':
            h = ops.linear(h, w, b, residual=feats, want_f32=True, want_bf16=False)
        else:
            h = ops.linear(h, w, b, relu=not last, residual=feats if last else None)
    return h


def _mha_pack(attn: nn.MultiheadAttention, precision: str):
    E = attn.embed_dim
    w, b = _f(attn.in_proj_weight), _f(attn.in_proj_bias)
    return {"w_in": wrap_weight(w, precision), "b_in": b,
            "w_q": wrap_weight(w[:E].contiguous(), precision), "b_q": b[:E].contiguous(),
            "w_kv": wrap_weight(w[E:].contiguous(), precision), "b_kv": b[E:].contiguous(),
            "w_out": wrap_weight(_f(attn.out_proj.weight), precision), "b_out": _f(attn.out_proj.bias)}


def _mha_sources(attn: nn.MultiheadAttention):
    return [attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.weight, attn.out_proj.bias]


def _ffn_pack(layer, precision: str):
    return {"w1": wrap_weight(_f(layer.linear1.weight), precision), "b1": _f(layer.linear1.bias),
            "w2": wrap_weight(_f(layer.linear2.weight), precision), "b2": _f(layer.linear2.bias)}


def _ln(norm: nn.LayerNorm):
    return (_f(norm.weight), _f(norm.bias), float(norm.eps))


class TransformerEncoderLayer(nn.Module):
    """encoders.py:167-197.  NOT the textbook post-norm layer: the residual stream itself is normalised and the
    sub-layer output is added AFTER the norm (``src = norm1(src); src = src + src2``, :188-189, :191-192)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu"):
        super(TransformerEncoderLayer, self).__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = torch.nn.functional.relu

    def _packed(self, precision: str):
        src = _mha_sources(self.self_attn) + [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                                              self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias]

        def build():
            pk = {"sa": _mha_pack(self.self_attn, precision), "n1": _ln(self.norm1), "n2": _ln(self.norm2)}
            pk.update(_ffn_pack(self, precision))
            return pk
        return cached(self, ("tf_enc_layer", precision), src, build)

    def forward_btd(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, T, d] in the activation dtype -> same."""
        pk = self._packed(config.PRECISION)
        sa = pk["sa"]
        E = x.shape[-1]
        qkv = ops.linear(x, sa["w_in"], sa["b_in"])                                   # [B,T,3E]
        att = ops.mha(qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:], self.self_attn.num_heads)
        n = ops.layer_norm(x, *pk["n1"])                                              # src = norm1(src)
        x = ops.linear(att, sa["w_out"], sa["b_out"], residual=n)                     # src = src + src2
        h = ops.linear(x, pk["w1"], pk["b1"], relu=True)
        n = ops.layer_norm(x, *pk["n2"])
        return ops.linear(h, pk["w2"], pk["b2"], residual=n)

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        """src [T, B, d] like the reference (masks are not supported: the reference never passes one)."""
        if src_mask is not None or src_key_padding_mask is not None:
            raise NotImplementedError("attention masks are not used by the reference's encoder")
        require_inference(src)
        x = ops.cast(src.permute(1, 0, 2).contiguous(), act_dtype(config.PRECISION))
        return self.forward_btd(x).permute(1, 0, 2)


class _LayerStack(nn.Module):
    """Container with the attribute names of nn.TransformerEncoder / the reference's TransformerDecoder
    (decoders.py:321-346): ``layers`` (deep copies of one initialised layer, so all start equal) and ``norm``."""

    def __init__(self, layer, num_layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm


class TF_Encoder(nn.Module):
    """Transformer encoder: col feats [B, T, HIDDEN_DIM] -> [B, T, HIDDEN_DIM] (encoders.py:68-143)."""

    def __init__(self):
        super(TF_Encoder, self).__init__()
        self.pos_encoder = PositionalEncoding(config.HIDDEN_DIM)
        self.encoder_layer = TransformerEncoderLayer(d_model=config.HIDDEN_DIM, nhead=8, dim_feedforward=2048, dropout=0.1)
        self.layer_norm = nn.LayerNorm(config.HIDDEN_DIM)
        self.encoder = _LayerStack(self.encoder_layer, num_layers=6, norm=self.layer_norm)
        self.pre_encoder_mlp = bool(config.PRE_ENCODER_MLP)
        if self.pre_encoder_mlp:
            # encoders.py:83-93: relevance scores of the objects per column, and the map back to the hidden size
            self.sem_relevence_mlp = MLP(input_size=(config.HIDDEN_DIM + config.EMBED_DIM), hidden_size=config.HIDDEN_DIM,
                                         num_classes=1, num_layers=3)
            self.combine_mlp = MLP(input_size=(config.HIDDEN_DIM + config.EMBED_DIM), hidden_size=config.HIDDEN_DIM,
                                   num_classes=config.HIDDEN_DIM, num_layers=3)
            self.emb_to_hid = nn.Linear(config.EMBED_DIM, config.HIDDEN_DIM, bias=False)   # never used by the reference either

    def get_relevant_semantic(self, feats, sem_vec, is_train=False):
        return relevant_semantic(ops.cast(feats.contiguous(), act_dtype(config.PRECISION)), sem_vec, self.sem_relevence_mlp)

    def forward_features(self, col_feats: torch.Tensor, semantics=None) -> torch.Tensor:
        """[B, T, d] (f32 or activation dtype) -> [B, T, d] in the activation dtype."""
        require_inference(col_feats)
        if self.training:
            raise RuntimeError("TF_Encoder is in train mode: dropout is not implemented - call model.eval()")
        dt = act_dtype(config.PRECISION)
        col_feats = col_feats.contiguous()
        if self.pre_encoder_mlp:
            # encoders.py:128-134: input = col_feats + combine_mlp(cat(col_feats, relevant semantics))
            if semantics is None:
                raise ValueError("config.PRE_ENCODER_MLP needs the semantic vectors")
            feats = ops.cast(col_feats, dt)
            col_feats = combine_with_semantic(feats, relevant_semantic(feats, semantics, self.sem_relevence_mlp), self.combine_mlp)
        x = ops.add_pos_encoding(col_feats, self.pos_encoder.table().contiguous(), dt)
        for layer in self.encoder.layers:
            x = layer.forward_btd(x)
        return ops.layer_norm(x, *_ln(self.encoder.norm))

    def forward(self, col_feats, semantics, is_train=False):
        return ops.cast(self.forward_features(col_feats, semantics), torch.float32)


class TransformerDecoderLayer(nn.Module):
    """decoders.py:348-466 with the MULTIHEAD_* switches off: the textbook post-norm decoder layer.  ``semantic_to_emb``
    is created (and saved in checkpoints) by the reference but never used."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super(TransformerDecoderLayer, self).__init__()
        _no_fusion('MULTIHEAD_PRE_TARGET', 'MULTIHEAD_PRE_MEMORY', 'MULTIHEAD_POST_MEMORY')
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.semantic_to_emb = nn.Linear(config.HIDDEN_DIM, config.EMBED_DIM)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.activation = torch.nn.functional.relu

    def _packed(self, precision: str):
        src = (_mha_sources(self.self_attn) + _mha_sources(self.multihead_attn) +
               [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias] +
               [p for n in (self.norm1, self.norm2, self.norm3) for p in (n.weight, n.bias)])

        def build():
            pk = {"sa": _mha_pack(self.self_attn, precision), "ca": _mha_pack(self.multihead_attn, precision),
                  "n1": _ln(self.norm1), "n2": _ln(self.norm2), "n3": _ln(self.norm3)}
            pk.update(_ffn_pack(self, precision))
            return pk
        return cached(self, ("tf_dec_layer", precision), src, build)

    def memory_kv(self, memory: torch.Tensor) -> torch.Tensor:
        """k | v projections of the memory for this layer's multihead_attn: [B, T, 2E] (computed once per forward)."""
        ca = self._packed(config.PRECISION)["ca"]
        return ops.linear(memory, ca["w_kv"], ca["b_kv"])

    def _tail(self, x, att, pk, mem_att):
        """From the self-attention core's output to the layer's output; ``mem_att(q)`` is the memory attention core."""
        sa, ca = pk["sa"], pk["ca"]
        x = ops.layer_norm(ops.linear(att, sa["w_out"], sa["b_out"], residual=x), *pk["n1"])      # :441-444
        q = ops.linear(x, ca["w_q"], ca["b_q"])
        x = ops.layer_norm(ops.linear(mem_att(q), ca["w_out"], ca["b_out"], residual=x), *pk["n2"])   # :450-453
        h = ops.linear(x, pk["w1"], pk["b1"], relu=True)
        return ops.layer_norm(ops.linear(h, pk["w2"], pk["b2"], residual=x), *pk["n3"])           # :461-463

    def forward_sequence(self, x: torch.Tensor, mem_kv: torch.Tensor) -> torch.Tensor:
        """Whole target sequence under the causal mask (teacher forcing): x [B, L, E] -> [B, L, E]."""
        pk = self._packed(config.PRECISION)
        E, H = x.shape[-1], self.self_attn.num_heads
        qkv = ops.linear(x, pk["sa"]["w_in"], pk["sa"]["b_in"])
        att = ops.mha(qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:], H, causal=True)
        return self._tail(x, att, pk, lambda q: ops.mha(q, mem_kv[..., :E], mem_kv[..., E:], H))

    def forward_step(self, x: torch.Tensor, pos: int, k_cache: torch.Tensor, v_cache: torch.Tensor,
                     mem_kv: torch.Tensor) -> torch.Tensor:
        """One target position: x [B, E]; this position's k / v rows are appended to the caches [B, Lmax, E]."""
        pk = self._packed(config.PRECISION)
        E, H = x.shape[-1], self.self_attn.num_heads
        qkv = ops.linear(x, pk["sa"]["w_in"], pk["sa"]["b_in"])                      # [B, 3E]
        att = ops.mha_step(qkv[:, :E], k_cache, v_cache, pos + 1, H, k_new=qkv[:, E:2 * E], v_new=qkv[:, 2 * E:])
        return self._tail(x, att, pk,
                          lambda q: ops.mha_step(q, mem_kv[..., :E], mem_kv[..., E:], mem_kv.shape[1], H))

    def forward(self, tgt, memory, semantics, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, is_train=False):
        """tgt [L, B, E], memory [T, B, E] like the reference; ``tgt_mask`` given means the causal mask (the only
        mask the reference builds, decoders.py:142-145)."""
        require_inference(tgt, memory)
        dt = act_dtype(config.PRECISION)
        x = ops.cast(tgt.permute(1, 0, 2).contiguous(), dt)
        mem_kv = self.memory_kv(ops.cast(memory.permute(1, 0, 2).contiguous(), dt))
        if tgt_mask is None:
            raise NotImplementedError("the reference always passes the square subsequent mask")
        return self.forward_sequence(x, mem_kv).permute(1, 0, 2)


class TransformerDecoder(_LayerStack):
    """decoders.py:321-346."""


class TF_Decoder(nn.Module):
    """Transformer decoder: encoder output [B, T, HIDDEN_DIM] -> logits [B, steps, num_classes] (decoders.py:103-275):
    greedy over MAX_TEXT_LENGTH steps at inference, teacher-forced over MAX_TEXT_LENGTH + 1 positions in training
    mode."""

    def __init__(self, num_classes):
        super(TF_Decoder, self).__init__()
        self.decoder_layer = TransformerDecoderLayer(d_model=config.EMBED_DIM, nhead=8, dim_feedforward=2048, dropout=0.1)
        self.layer_norm = nn.LayerNorm(config.EMBED_DIM)
        self.decoder = TransformerDecoder(self.decoder_layer, num_layers=6, norm=self.layer_norm)
        self.pos_encoder = PositionalEncoding(config.EMBED_DIM)
        self.hid_to_emb = nn.Linear(config.HIDDEN_DIM, config.EMBED_DIM)
        self.emb = nn.Embedding(num_classes, config.EMBED_DIM)
        self.emb_to_classes = nn.Linear(config.EMBED_DIM, num_classes)
        self.num_classes = num_classes
        self.pre_decoder_mlp = bool(config.PRE_DECODER_MLP)
        if self.pre_decoder_mlp:      # decoders.py:122-124
            self.relevant_mlp = MLP(input_size=(config.EMBED_DIM * 2), hidden_size=config.EMBED_DIM, num_classes=1, num_layers=3)
            self.combine_mlp = MLP(input_size=(config.EMBED_DIM * 2), hidden_size=config.EMBED_DIM, num_classes=config.EMBED_DIM,
                                   num_layers=2)
        # decoders.py:126-130.  get_semantic_cls (:196-201) soft-maxes the relevant semantics over the columns and then SUMS
        # over the same axis: every entry is 1 whatever sem_cls_mlp computes, so the first target position's embedding is
        # replaced by a vector of ones (the MLP is created for the state dict and never evaluated).
        self.sem_cls = bool(config.CLS_DECODER_INIT)
        if self.sem_cls:
            self.sem_cls_mlp = MLP(input_size=(config.EMBED_DIM * 2), hidden_size=config.EMBED_DIM, num_classes=1, num_layers=3)
        self.post_decoder = bool(config.POST_DECODER_MLP)
        if self.post_decoder:         # decoders.py:132-135 (attribute names as in the reference, typo included)
            self.post_decoder_mlp = MLP(input_size=(num_classes * 2), hidden_size=num_classes, num_classes=1, num_layers=3)
            self.post_deocer_combine_mlp = MLP(input_size=(num_classes * 2), hidden_size=num_classes, num_classes=num_classes,
                                               num_layers=3)
            self.sem_to_classes = nn.Linear(config.EMBED_DIM, num_classes)

    def _generate_square_subsequent_mask(self, sz):
        mask = (torch.triu(torch.ones(sz, sz)) == 1).transpose(0, 1)
        return mask.float().masked_fill(mask == 0, float('-inf')).masked_fill(mask == 1, float(0.0))

    def _packed(self, precision: str):
        src = [self.hid_to_emb.weight, self.hid_to_emb.bias, self.emb.weight, self.emb_to_classes.weight,
               self.emb_to_classes.bias]

        def build():
            C, E = self.emb_to_classes.weight.shape
            Cp = (C + 127) // 128 * 128                   # GEMM tiles want a multiple of 128 output columns
            w = torch.zeros((Cp, E), device=self.emb.weight.device, dtype=torch.float32)
            b = torch.zeros((Cp,), device=self.emb.weight.device, dtype=torch.float32)
            w[:C] = self.emb_to_classes.weight.detach().float()
            b[:C] = self.emb_to_classes.bias.detach().float()
            return {"w_mem": wrap_weight(_f(self.hid_to_emb.weight), precision), "b_mem": _f(self.hid_to_emb.bias),
                    "emb": _f(self.emb.weight), "w_cls": wrap_weight(w, precision), "b_cls": b}
        return cached(self, ("tf_dec", precision), src, build)

    def _classes(self, x, pk):
        """emb_to_classes on the normalised decoder output: f32 [..., Cpad]."""
        if config.PRECISION == 'bf16':
            return ops.linear(x, pk["w_cls"], pk["b_cls"], want_f32=True, want_bf16=False)
        return ops.linear(x, pk["w_cls"], pk["b_cls"])

    def get_relevant_semantic(self, feats, sem_vec, mlp, is_train=False):
        return relevant_semantic(ops.cast(feats.contiguous(), act_dtype(config.PRECISION)), sem_vec, mlp)

    def decode(self, encoder_output, text, is_train, enc_act=None, semantics=None):
        """dict(logits [B, steps, C] f32, ids [B, steps] i32, lengths [B] i32)."""
        require_inference(encoder_output)
        if self.training:
            raise RuntimeError("TF_Decoder is in train mode: dropout is not implemented - call model.eval()")
        precision = config.PRECISION
        dt = act_dtype(precision)
        pk = self._packed(precision)
        enc = enc_act if (enc_act is not None and enc_act.dtype == dt) else ops.cast(encoder_output.contiguous(), dt)
        B, T, _ = enc.shape
        dev = enc.device
        C, E = self.num_classes, self.emb.weight.shape[1]
        pe = self.pos_encoder.table().contiguous()
        memory = ops.linear(enc, pk["w_mem"], pk["b_mem"])                            # hid_to_emb: [B, T, E]
        if self.pre_decoder_mlp:
            # decoders.py:218-222: memory = memory + combine_mlp(cat(memory, relevant semantics))
            if semantics is None:
                raise ValueError("config.PRE_DECODER_MLP needs the semantic vectors")
            memory = combine_with_semantic(memory, relevant_semantic(memory, semantics, self.relevant_mlp), self.combine_mlp)
        layers = list(self.decoder.layers)
        mem_kv = [layer.memory_kv(memory) for layer in layers]
        final = _ln(self.decoder.norm)
        if is_train:
            # teacher forcing (decoders.py:228-248): all MAX_TEXT_LENGTH + 1 target positions under the causal mask
            if text is None:
                raise ValueError("is_train=True needs the teacher-forcing text (decoders.py:230-233)")
            L = config.MAX_TEXT_LENGTH + 1
            tokens = text.to(device=dev)[:, :L].to(torch.int32).contiguous()
            if tokens.shape[1] != L:
                raise ValueError("text must hold MAX_TEXT_LENGTH + 1 = %d target positions" % L)
            x = ops.embed_tokens(tokens, pk["emb"], pe, 0, dt)
            if self.sem_cls:
                x[:, 0, :] = self._cls_embedding(B, pe, dt, dev)[:, 0, :]
            for layer, kv in zip(layers, mem_kv):
                x = layer.forward_sequence(x, kv)
            cls = self._classes(ops.layer_norm(x, *final), pk)                        # [B, L, Cpad] f32
            logits = torch.empty((B, L, C), device=dev, dtype=torch.float32)
            ids = torch.empty((B, L), device=dev, dtype=torch.int32)
            ops.argmax_tokens(cls.view(B * L, -1), C, logits.view(B * L, C), ids.view(B * L), None)
            return self._post(logits, ids, semantics)
        # greedy (decoders.py:250-279): position t reads targets[t] (targets[0] = 0 = [GO]) and yields targets[t+1]
        steps = config.MAX_TEXT_LENGTH
        logits = torch.empty((B, steps, C), device=dev, dtype=torch.float32)
        ids = torch.empty((B, steps), device=dev, dtype=torch.int32)
        tok = torch.zeros((B, 1), device=dev, dtype=torch.int32)
        k_cache = [torch.empty((B, steps, E), device=dev, dtype=dt) for _ in layers]
        v_cache = [torch.empty((B, steps, E), device=dev, dtype=dt) for _ in layers]
        for t in range(steps):
            x = (self._cls_embedding(B, pe, dt, dev) if (t == 0 and self.sem_cls)
                 else ops.embed_tokens(tok, pk["emb"], pe, t, dt)).view(B, E)
            for i, layer in enumerate(layers):
                x = layer.forward_step(x, t, k_cache[i], v_cache[i], mem_kv[i])
            cls = self._classes(ops.layer_norm(x, *final), pk)                        # [B, Cpad] f32
            ops.argmax_tokens(cls, C, logits[:, t], ids[:, t], tok.view(B))
        return self._post(logits, ids, semantics)

    def _cls_embedding(self, B, pe, dt, dev):
        """CLS_DECODER_INIT: ones (see __init__) + pe[0], [B, 1, E] in the activation dtype."""
        return ops.add_pos_encoding(torch.ones((B, 1, pe.shape[1]), device=dev, dtype=torch.float32), pe, dt)

    def _post(self, logits, ids, semantics):
        """POST_DECODER_MLP (decoders.py:284-295): output = output + post_deocer_combine_mlp(cat(output, relevant semantics))
        with the semantics mapped to class space by sem_to_classes.  The greedy feedback inside the loop used the logits
        BEFORE this step (decoders.py:276-279), as here; the returned ids are the argmax of the final logits.  The
        97-wide layers run on GEMM tiles of 128 with zero padding."""
        if not self.post_decoder:
            return {"logits": logits, "ids": ids, "lengths": _lengths(ids)}
        if semantics is None:
            raise ValueError("config.POST_DECODER_MLP needs the semantic vectors")
        precision = config.PRECISION
        dt = act_dtype(precision)
        B, S, C = logits.shape
        pad = 128
        Cp = -(-C // pad) * pad
        stc = self.sem_to_classes
        w_s, b_s = cached(stc, ("sem_to_classes", precision), [stc.weight, stc.bias],
                          lambda: (wrap_weight(_pad_w(_f(stc.weight), pad), precision), _pad_v(_f(stc.bias), pad)))
        sem = ops.cast(semantics.contiguous().float(), dt)
        sem_c = ops.linear(sem, w_s, b_s, want_f32=True, want_bf16=False) if precision == 'bf16' else ops.linear(sem, w_s, b_s)
        feats32 = torch.zeros((B, S, Cp), device=logits.device, dtype=torch.float32)
        feats32[..., :C] = logits
        feats = ops.cast(feats32, dt)
        rel = relevant_semantic(feats, sem_c, self.post_decoder_mlp, feat_dim=C, pad=pad)       # [B, S, Cp]
        out = combine_with_semantic(feats, rel, self.post_deocer_combine_mlp, feat_dim=C, pad=pad, want_f32=True)
        final = torch.empty((B, S, C), device=logits.device, dtype=torch.float32)
        ids2 = torch.empty((B, S), device=logits.device, dtype=torch.int32)
        ops.argmax_tokens(ops.cast(out, torch.float32).view(B * S, Cp), C, final.view(B * S, C), ids2.view(B * S), None)
        return {"logits": final, "ids": ids2, "lengths": _lengths(ids2)}

    def forward(self, encoder_output, text, semantics, is_train):
        return self.decode(encoder_output, text, is_train, semantics=semantics)["logits"]


class Linear_Decoder(nn.Module):
    """One nn.Linear over the encoder output (decoders.py:278-296): [B, T, HIDDEN_DIM] -> [B, T, num_classes]."""

    def __init__(self, num_classes):
        super(Linear_Decoder, self).__init__()
        self.linear_decoder = nn.Linear(config.HIDDEN_DIM, num_classes)
        self.num_classes = num_classes
        self.init_weights()

    def init_weights(self):
        initrange = 0.1
        self.linear_decoder.bias.data.zero_()
        self.linear_decoder.weight.data.uniform_(-initrange, initrange)

    def decode(self, encoder_output, text, is_train, enc_act=None, semantics=None):
        require_inference(encoder_output)
        precision = config.PRECISION
        dt = act_dtype(precision)
        lin = self.linear_decoder

        def build():
            C, K = lin.weight.shape
            Cp = (C + 127) // 128 * 128
            w = torch.zeros((Cp, K), device=lin.weight.device, dtype=torch.float32)
            b = torch.zeros((Cp,), device=lin.weight.device, dtype=torch.float32)
            w[:C] = lin.weight.detach().float()
            b[:C] = lin.bias.detach().float()
            return {"w": wrap_weight(w, precision), "b": b}
        pk = cached(self, ("linear_dec", precision), [lin.weight, lin.bias], build)
        enc = enc_act if (enc_act is not None and enc_act.dtype == dt) else ops.cast(encoder_output.contiguous(), dt)
        B, T, _ = enc.shape
        C = self.num_classes
        cls = ops.linear(enc, pk["w"], pk["b"], want_f32=True, want_bf16=False) if precision == 'bf16' \
            else ops.linear(enc, pk["w"], pk["b"])
        logits = torch.empty((B, T, C), device=enc.device, dtype=torch.float32)
        ids = torch.empty((B, T), device=enc.device, dtype=torch.int32)
        ops.argmax_tokens(cls.view(B * T, -1), C, logits.view(B * T, C), ids.view(B * T), None)
        return {"logits": logits, "ids": ids, "lengths": _lengths(ids)}

    def forward(self, encoder_output, text, semantics, is_train):
        return self.decode(encoder_output, text, is_train)["logits"]


class Oscar_Bert(nn.Module):
    """BERT-base encoder over the column features (encoders.py:34-66): ``hid_to_bert`` (512 -> 768), a randomly
    initialised ``transformers.BertModel(BertConfig())`` fed through ``inputs_embeds`` (12 post-norm layers, 12 heads x
    64, exact-erf GELU feed-forward of 3072, LayerNorm eps 1e-12), ``bert_to_hid`` (768 -> 512).  ``bert_model`` is kept
    as the parameter container (203 state-dict tensors, the reference's names; its word-embedding table and pooler are
    never used, as in the reference); its forward is not called.  ``config.OSCAR_ENCODER`` (concatenating the semantic
    vectors with segment ids, "No pretrained models yet", config.py:46) is one of the fusion switches: it raises."""

    def __init__(self):
        super(Oscar_Bert, self).__init__()
        _no_fusion('OSCAR_ENCODER')
        from transformers import BertConfig, BertModel      # what the reference imports (encoders.py:32)
        self.bert_config = BertConfig()
        self.bert_model = BertModel(self.bert_config)
        self.hid_to_bert = nn.Linear(512, 768)
        self.bert_to_hid = nn.Linear(768, 512)

    def _packed(self, precision: str):
        bm = self.bert_model
        src = [p for p in bm.encoder.parameters()] + [bm.embeddings.position_embeddings.weight,
                                                      bm.embeddings.token_type_embeddings.weight,
                                                      bm.embeddings.LayerNorm.weight, bm.embeddings.LayerNorm.bias,
                                                      self.hid_to_bert.weight, self.hid_to_bert.bias,
                                                      self.bert_to_hid.weight, self.bert_to_hid.bias]

        def build():
            emb = bm.embeddings
            layers = []
            for lyr in bm.encoder.layer:
                a = lyr.attention
                w_qkv = torch.cat([_f(a.self.query.weight), _f(a.self.key.weight), _f(a.self.value.weight)], 0)
                b_qkv = torch.cat([_f(a.self.query.bias), _f(a.self.key.bias), _f(a.self.value.bias)], 0)
                layers.append({
                    "w_qkv": wrap_weight(w_qkv, precision), "b_qkv": b_qkv.contiguous(),
                    "w_o": wrap_weight(_f(a.output.dense.weight), precision), "b_o": _f(a.output.dense.bias),
                    "n1": _ln(a.output.LayerNorm),
                    "w1": wrap_weight(_f(lyr.intermediate.dense.weight), precision), "b1": _f(lyr.intermediate.dense.bias),
                    "w2": wrap_weight(_f(lyr.output.dense.weight), precision), "b2": _f(lyr.output.dense.bias),
                    "n2": _ln(lyr.output.LayerNorm)})
            # inputs_embeds + token_type_embeddings[0] + position_embeddings[t] (BertEmbeddings.forward with
            # token_type_ids = None -> zeros, position ids 0..T-1): one table, one fp32 add on the host side
            table = (_f(emb.position_embeddings.weight) + _f(emb.token_type_embeddings.weight)[0][None, :]).contiguous()
            return {"layers": layers, "table": table, "n0": _ln(emb.LayerNorm),
                    "w_in": wrap_weight(_f(self.hid_to_bert.weight), precision), "b_in": _f(self.hid_to_bert.bias),
                    "w_out": wrap_weight(_f(self.bert_to_hid.weight), precision), "b_out": _f(self.bert_to_hid.bias)}
        return cached(self, ("oscar", precision), src, build)

    def forward_features(self, col_feats: torch.Tensor) -> torch.Tensor:
        """[B, T, 512] (f32 or activation dtype) -> [B, T, 512] in the activation dtype."""
        require_inference(col_feats)
        if self.training:
            raise RuntimeError("Oscar_Bert is in train mode: dropout is not implemented - call model.eval()")
        cfg = self.bert_config
        if cfg.hidden_act != "gelu" or getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise NotImplementedError("only the default BertConfig (gelu, absolute positions) is implemented")
        precision = config.PRECISION
        dt = act_dtype(precision)
        pk = self._packed(precision)
        heads = cfg.num_attention_heads
        x = ops.linear(ops.cast(col_feats.contiguous(), dt), pk["w_in"], pk["b_in"])          # hid_to_bert
        E = x.shape[-1]
        x = ops.layer_norm(ops.add_pos_encoding(x, pk["table"], dt), *pk["n0"])               # BertEmbeddings
        for lp in pk["layers"]:
            qkv = ops.linear(x, lp["w_qkv"], lp["b_qkv"])
            att = ops.mha(qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:], heads)
            x = ops.layer_norm(ops.linear(att, lp["w_o"], lp["b_o"], residual=x), *lp["n1"])  # BertSelfOutput
            h = ops.gelu_(ops.linear(x, lp["w1"], lp["b1"]))                                  # BertIntermediate
            x = ops.layer_norm(ops.linear(h, lp["w2"], lp["b2"], residual=x), *lp["n2"])      # BertOutput
        return ops.linear(x, pk["w_out"], pk["b_out"])                                        # bert_to_hid

    def forward(self, col_feats, semantics, is_train):
        return ops.cast(self.forward_features(col_feats), torch.float32)


def _lengths(ids: torch.Tensor) -> torch.Tensor:
    """Index of the first [s] (id 1) per crop, or the number of steps (what training_functions.py:188-189 cuts at)."""
    stop = ids == 1
    first = torch.argmax(stop.to(torch.int32), dim=1)
    return torch.where(stop.any(dim=1), first, torch.full_like(first, ids.shape[1])).to(torch.int32)
