"""CPU ORACLE (test infrastructure, NOT a product path) for the transformer encoder / decoder family.

A numpy restatement of the reference's ``TF_Encoder`` (modules/encoders.py:68-197), ``TF_Decoder``
(modules/decoders.py:103-275, :321-466) and ``Linear_Decoder`` (:278-296) with every fusion switch of config.py:40-53
at its default (False), in eval mode (dropout = identity).  The arithmetic lives in PyTorch ATen (nn.MultiheadAttention,
nn.LayerNorm, nn.Linear, nn.Embedding; third-party, un-vendored; the image has torch 2.11.0), so the operator semantics
restated here are ATen's: q is scaled by 1/sqrt(head dim) before q k^T, softmax over the keys, LayerNorm with the biased
variance and eps 1e-5.

The greedy loop is restated the way the reference runs it - the whole target prefix 0..t goes through the decoder at
every step under the square subsequent mask (decoders.py:250-279), no K/V cache - so that the CUDA path's cache is
checked against the recomputation, not against itself.

PINNING: ``tests/golden/make_golden_tf.py`` (build container) runs the reference's OWN module objects on the seeded
synthetic weights of ``synth.make_state_dict_transformer`` and stores their outputs in ``tests/golden/cfg_tf.npz``;
``tests/test_oracle_golden.py`` checks this file against them.  One caveat, stated there too: the reference's
``TF_Encoder.forward`` itself raises TypeError under torch 2.11 (``nn.TransformerEncoder.forward`` passes
``is_causal=`` to the reference's layer class, which does not take it), so the fixture drives the reference's
``pos_encoder``, its six ``TransformerEncoderLayer`` objects and its ``norm`` from a three-line loop - what
``nn.TransformerEncoder.forward`` did in the torch the reference was written for.  ``TF_Decoder.forward`` and
``Linear_Decoder.forward`` run unmodified.

Only ``tests/`` may import this module.  The product path (the CUDA extension) never does.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

LN_EPS = 1e-5  # nn.LayerNorm default (encoders.py:78,176-177; decoders.py:112,364-366)


def linear(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def layer_norm(x, w, b):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(LN_EPS)) * w + b


def softmax(x, axis):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def multihead_attention(q_in, k_in, v_in, sd, prefix, heads, mask=None):
    """nn.MultiheadAttention.forward(query, key, value, attn_mask)[0] in eval mode, batch-first here:
    q_in [B, Lq, E], k_in / v_in [B, Lk, E]; ``mask`` [Lq, Lk] additive (0 / -inf)."""
    dt = q_in.dtype
    w, b = sd[prefix + "in_proj_weight"].astype(dt), sd[prefix + "in_proj_bias"].astype(dt)
    E = q_in.shape[-1]
    dh = E // heads
    q = linear(q_in, w[:E], b[:E]) * dt.type(1.0 / np.sqrt(dh))
    k = linear(k_in, w[E:2 * E], b[E:2 * E])
    v = linear(v_in, w[2 * E:], b[2 * E:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.reshape(B, Lq, heads, dh).transpose(0, 2, 1, 3)
    k = k.reshape(B, Lk, heads, dh).transpose(0, 2, 1, 3)
    v = v.reshape(B, Lk, heads, dh).transpose(0, 2, 1, 3)
    s = q @ k.transpose(0, 1, 3, 2)                         # [B, heads, Lq, Lk]
    if mask is not None:
        s = s + mask.astype(dt)
    o = (softmax(s, axis=-1) @ v).transpose(0, 2, 1, 3).reshape(B, Lq, E)
    return linear(o, sd[prefix + "out_proj.weight"].astype(dt), sd[prefix + "out_proj.bias"].astype(dt))


def _ln(x, sd, prefix):
    return layer_norm(x, sd[prefix + "weight"].astype(x.dtype), sd[prefix + "bias"].astype(x.dtype))


def _ffn(x, sd, prefix):
    dt = x.dtype
    h = np.maximum(linear(x, sd[prefix + "linear1.weight"].astype(dt), sd[prefix + "linear1.bias"].astype(dt)), 0)
    return linear(h, sd[prefix + "linear2.weight"].astype(dt), sd[prefix + "linear2.bias"].astype(dt))


def encoder_layer(src, sd, prefix, heads=8):
    """TransformerEncoderLayer.forward (encoders.py:185-193): the sub-layer output is added AFTER the norm of the
    residual stream - ``src = norm1(src); src = src + src2``."""
    src2 = multihead_attention(src, src, src, sd, prefix + "self_attn.", heads)
    src = _ln(src, sd, prefix + "norm1.") + src2
    src2 = _ffn(src, sd, prefix)
    return _ln(src, sd, prefix + "norm2.") + src2


def tf_encoder(col_feats, sd, dt=np.float32, layers=6, prefix="encoder.", semantics=None):
    """TF_Encoder.forward (encoders.py:127-143): [PRE_ENCODER_MLP fusion when ``semantics`` is given,] positional
    encoding, the layers, the norm."""
    x = col_feats.astype(dt)
    if semantics is not None:
        x = pre_mlp_fusion(x, semantics, sd, prefix + "sem_relevence_mlp.", prefix + "combine_mlp.")
    pe = sd[prefix + "pos_encoder.pe"].astype(dt)[:, 0, :]
    x = x + pe[None, :x.shape[1], :]
    for i in range(layers):
        x = encoder_layer(x, sd, "%sencoder.layers.%d." % (prefix, i))
    return _ln(x, sd, prefix + "encoder.norm.")


def decoder_layer(tgt, memory, sd, prefix, mask, heads=8):
    """TransformerDecoderLayer.forward (decoders.py:433-466) with the MULTIHEAD_* switches off."""
    tgt = _ln(tgt + multihead_attention(tgt, tgt, tgt, sd, prefix + "self_attn.", heads, mask), sd, prefix + "norm1.")
    tgt = _ln(tgt + multihead_attention(tgt, memory, memory, sd, prefix + "multihead_attn.", heads), sd, prefix + "norm2.")
    return _ln(tgt + _ffn(tgt, sd, prefix), sd, prefix + "norm3.")


def _causal_mask(n):
    """_generate_square_subsequent_mask (decoders.py:142-145)."""
    m = np.zeros((n, n), np.float64)
    m[np.triu_indices(n, 1)] = -np.inf
    return m


def _decoder_stack(x, memory, sd, prefix, layers):
    mask = _causal_mask(x.shape[1])
    for i in range(layers):
        x = decoder_layer(x, memory, sd, "%sdecoder.layers.%d." % (prefix, i), mask)
    return _ln(x, sd, prefix + "decoder.norm.")


def tf_decoder(enc, sd, max_len=25, text: Optional[np.ndarray] = None, dt=np.float32, layers=6, prefix="decoder.",
               semantics=None, cls_init=False, post_semantics=None):
    """TF_Decoder.forward (decoders.py:213-296 of the class body, :228-279 the two branches).
    text None: greedy, logits [B, max_len, C]; text [B, >= max_len + 1]: teacher forcing, logits [B, max_len + 1, C]."""
    enc = enc.astype(dt)
    memory = linear(enc, sd[prefix + "hid_to_emb.weight"].astype(dt), sd[prefix + "hid_to_emb.bias"].astype(dt))
    if semantics is not None:            # PRE_DECODER_MLP (decoders.py:218-222)
        memory = pre_mlp_fusion(memory, semantics, sd, prefix + "relevant_mlp.", prefix + "combine_mlp.")
    emb = sd[prefix + "emb.weight"].astype(dt)
    pe = sd[prefix + "pos_encoder.pe"].astype(dt)[:, 0, :]
    w_c, b_c = sd[prefix + "emb_to_classes.weight"].astype(dt), sd[prefix + "emb_to_classes.bias"].astype(dt)
    B = enc.shape[0]
    def embed(tokens):
        e = emb[tokens].copy()
        if cls_init:
            # CLS_DECODER_INIT (decoders.py:196-201,236-238,260-262): softmax over the columns of the relevant semantics,
            # then the SUM over the same axis - a vector of ones whatever sem_cls_mlp computes - replaces position 0
            rel = get_relevant_semantic(memory, post_or_sem.astype(dt), sd, prefix + "sem_cls_mlp.")
            e[:, 0] = softmax(rel, axis=1).sum(axis=1)
        return e + pe[None, :tokens.shape[1]]

    def post(out):
        if post_semantics is None:
            return out
        # POST_DECODER_MLP (decoders.py:284-295)
        sem_c = linear(post_semantics.astype(dt), sd[prefix + "sem_to_classes.weight"].astype(dt), sd[prefix + "sem_to_classes.bias"].astype(dt))
        return pre_mlp_fusion(out, sem_c, sd, prefix + "post_decoder_mlp.", prefix + "post_deocer_combine_mlp.")
    post_or_sem = post_semantics if post_semantics is not None else semantics
    if text is not None:
        L = max_len + 1
        return post(linear(_decoder_stack(embed(text[:, :L]), memory, sd, prefix, layers), w_c, b_c))
    targets = np.zeros((B, max_len + 1), np.int64)
    out = np.zeros((B, max_len, w_c.shape[0]), dt)
    for t in range(max_len):
        x = embed(targets[:, :t + 1])                         # the reference embeds all positions and slices [:t+1]
        y = linear(_decoder_stack(x, memory, sd, prefix, layers), w_c, b_c)
        targets[:, t + 1] = y[:, -1].argmax(axis=1)
        out[:, t] = y[:, t]
    return post(out)


def linear_decoder(enc, sd, dt=np.float32, prefix="decoder."):
    """Linear_Decoder.forward (decoders.py:293-295)."""
    return linear(enc.astype(dt), sd[prefix + "linear_decoder.weight"].astype(dt), sd[prefix + "linear_decoder.bias"].astype(dt))


def gelu(x):
    """transformers' "gelu": x * 0.5 * (1 + erf(x / sqrt(2)))."""
    from scipy.special import erf
    return (x * 0.5 * (1.0 + erf(x / np.sqrt(2.0)))).astype(x.dtype)


def oscar_bert(col_feats, sd, dt=np.float32, prefix="encoder.", heads=12, layers=12, eps=1e-12):
    """Oscar_Bert.forward with OSCAR_ENCODER off (encoders.py:47-66): hid_to_bert, BertModel(inputs_embeds=...) -
    transformers' BertEmbeddings (inputs + token_type_embeddings[0] + position_embeddings[:T], LayerNorm), twelve
    post-norm BertLayers (separate query / key / value Linears, exact GELU) - and bert_to_hid on the last hidden state."""
    g = lambda k: sd[prefix + k].astype(dt)

    def ln(x, name):
        mu = x.mean(-1, keepdims=True)
        var = ((x - mu) ** 2).mean(-1, keepdims=True)
        return (x - mu) / np.sqrt(var + dt(eps)) * g(name + ".weight") + g(name + ".bias")
    dt = np.dtype(dt).type
    x = linear(col_feats.astype(dt), g("hid_to_bert.weight"), g("hid_to_bert.bias"))
    B, T, E = x.shape
    dh = E // heads
    x = x + g("bert_model.embeddings.token_type_embeddings.weight")[0] + g("bert_model.embeddings.position_embeddings.weight")[None, :T]
    x = ln(x, "bert_model.embeddings.LayerNorm")
    for i in range(layers):
        p = "bert_model.encoder.layer.%d." % i
        q = linear(x, g(p + "attention.self.query.weight"), g(p + "attention.self.query.bias"))
        k = linear(x, g(p + "attention.self.key.weight"), g(p + "attention.self.key.bias"))
        v = linear(x, g(p + "attention.self.value.weight"), g(p + "attention.self.value.bias"))
        split = lambda t: t.reshape(B, T, heads, dh).transpose(0, 2, 1, 3)
        s = (split(q) @ split(k).transpose(0, 1, 3, 2)) * dt(1.0 / np.sqrt(dh))
        ctx = (softmax(s, -1) @ split(v)).transpose(0, 2, 1, 3).reshape(B, T, E)
        x = ln(linear(ctx, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias")) + x, p + "attention.output.LayerNorm")
        h = gelu(linear(x, g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")))
        x = ln(linear(h, g(p + "output.dense.weight"), g(p + "output.dense.bias")) + x, p + "output.LayerNorm")
    return linear(x, g("bert_to_hid.weight"), g("bert_to_hid.bias"))


def mlp(x, sd, prefix):
    """MLP.forward (encoders.py:238-243): Linear layers ``layers.fc{i}`` with ReLU between them."""
    i = 0
    while (prefix + "layers.fc%d.weight" % (i + 1)) in sd:
        x = np.maximum(linear(x, sd[prefix + "layers.fc%d.weight" % i].astype(x.dtype), sd[prefix + "layers.fc%d.bias" % i].astype(x.dtype)), 0)
        i += 1
    return linear(x, sd[prefix + "layers.fc%d.weight" % i].astype(x.dtype), sd[prefix + "layers.fc%d.bias" % i].astype(x.dtype))


def get_relevant_semantic(feats, sem_vec, sd, mlp_prefix):
    """encoders.py:89-111 / decoders.py:147-170: the concatenation of every (column, object) pair through the MLP, softmax
    over the objects, weighted sum of the object vectors.  feats [B, T, F], sem_vec [B, O, E] -> [B, T, E]."""
    B, T, _ = feats.shape
    O = sem_vec.shape[1]
    sem_seq = np.broadcast_to(sem_vec[:, None], (B, T, O, sem_vec.shape[2]))
    col_seq = np.broadcast_to(feats[:, :, None], (B, T, O, feats.shape[2]))
    scores = softmax(mlp(np.concatenate([col_seq, sem_seq], axis=3), sd, mlp_prefix), axis=2)
    return (sem_seq * scores).sum(axis=2)


def pre_mlp_fusion(feats, sem_vec, sd, relevance_prefix, combine_prefix):
    """feats + combine_mlp(cat(feats, relevant semantics)) (encoders.py:128-134, decoders.py:218-222)."""
    rel = get_relevant_semantic(feats, sem_vec.astype(feats.dtype), sd, relevance_prefix)
    return feats + mlp(np.concatenate([feats, rel], axis=2), sd, combine_prefix)
