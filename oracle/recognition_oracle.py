"""CPU ORACLE (test infrastructure, NOT a product path) for the word-crop recognition forward.

A numpy restatement of the reference's LSTM/LSTM inference path
``Model.forward(input, text, overlap, scene, ious, is_train=False)`` (model.py:85-107):
TPS rectification -> ResNet -> column features -> 2x BidirectionalLSTM -> Attention greedy loop,
plus ``Linear_Embedding`` (semantic_vectors.py:33-49).  Every function cites the reference lines
it follows.  The arithmetic itself lives in PyTorch ATen (third-party, un-vendored, version not
pinned by the reference; the image has torch 2.11.0), so the operator semantics restated here are
ATen's: eval-mode BatchNorm, -inf padded max-pool, ``grid_sample(bilinear, border,
align_corners=True)``, LSTM gate order i,f,g,o.

PINNING: the reference has no tests, golden vectors or fixtures.  This oracle
is pinned against outputs of the *reference's own modules* executed in the build container:
``tests/golden/make_golden.py`` imports /root/reference, loads the seeded synthetic weights of
``multimodal_scene_text_recognition_b200.synth`` and stores stage tensors in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product path (the CUDA extension) never does.

``dtype``   np.float32 reproduces the reference precision, np.float64 gives the "true" value.
``quant``   None, or "bf16": insert round-to-nearest-even bf16 rounding at exactly the points
            where the CUDA bf16 throughput mode stores/reads bf16 (documented in DESIGN.md 4),
            so the GPU path can be compared with a quantisation-matched oracle as well.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

BN_EPS = 1e-5  # nn.BatchNorm2d default, feature_extraction.py:20 / transformation.py:50


# --------------------------------------------------------------------------------------
# rounding helpers
# --------------------------------------------------------------------------------------
def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round fp32 values to the nearest bf16 (ties to even), returned as fp32."""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).reshape(a32.shape)
    # NaN/Inf keep their class under this rounding for all finite inputs we produce
    return out.astype(a.dtype) if a.dtype != np.float32 else out


def _q(a: np.ndarray, quant: Optional[str]) -> np.ndarray:
    return round_bf16(a) if quant == "bf16" else a


# --------------------------------------------------------------------------------------
# primitive operators (ATen semantics)
# --------------------------------------------------------------------------------------
def conv2d(x: np.ndarray, w: np.ndarray, stride=(1, 1), pad=(0, 0)) -> np.ndarray:
    """NCHW cross-correlation without bias (nn.Conv2d(..., bias=False)); zero padding.

    im2col + GEMM, the same decomposition the reference runs with cuDNN disabled (model.py:6).
    """
    B, C, H, W = x.shape
    O, C2, kh, kw = w.shape
    assert C == C2
    sh, sw = stride
    ph, pw = pad
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)))
    Ho = (H + 2 * ph - kh) // sh + 1
    Wo = (W + 2 * pw - kw) // sw + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(2, 3))  # B,C,H',W',kh,kw
    win = win[:, :, ::sh, ::sw][:, :, :Ho, :Wo]
    cols = win.transpose(0, 2, 3, 1, 4, 5).reshape(B * Ho * Wo, C * kh * kw)
    out = cols @ w.reshape(O, C * kh * kw).T
    return out.reshape(B, Ho, Wo, O).transpose(0, 3, 1, 2)


def batchnorm_eval(x, weight, bias, mean, var):
    """Eval-mode BatchNorm2d: (x - mean) / sqrt(var + eps) * weight + bias  (SURVEY App. B)."""
    dt = x.dtype
    scale = (weight.astype(dt) / np.sqrt(var.astype(dt) + dt.type(BN_EPS)))
    shift = bias.astype(dt) - mean.astype(dt) * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def relu(x):
    return np.maximum(x, 0)


def maxpool2d(x, kernel=(2, 2), stride=(2, 2), pad=(0, 0)):
    """nn.MaxPool2d; padding is -inf (feature_extraction.py:78 maxpool3 relies on it)."""
    kh, kw = kernel
    sh, sw = stride
    ph, pw = pad
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)), constant_values=-np.inf)
    H, W = xp.shape[2:]
    Ho = (H - kh) // sh + 1
    Wo = (W - kw) // sw + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(2, 3))
    win = win[:, :, ::sh, ::sw][:, :, :Ho, :Wo]
    return win.max(axis=(4, 5))


def linear(x, w, b=None):
    y = x @ w.T.astype(x.dtype)
    return y if b is None else y + b.astype(x.dtype)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def grid_sample_bilinear_border(img: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """F.grid_sample(img, grid, mode='bilinear', padding_mode='border', align_corners=True)
    as called at transformation.py:33-34.  img [B,C,H,W], grid [B,Ho,Wo,2] (x, y) in [-1,1]."""
    B, C, H, W = img.shape
    dt = img.dtype
    gx = grid[..., 0].astype(dt)
    gy = grid[..., 1].astype(dt)
    ix = (gx + 1) / 2 * (W - 1)
    iy = (gy + 1) / 2 * (H - 1)
    ix = np.clip(ix, 0, W - 1)
    iy = np.clip(iy, 0, H - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    # ATen forms the corner weights from the neighbour coordinates ((x0+1) - ix, ix - x0, ...)
    wx1 = ix - x0
    wy1 = iy - y0
    wx0 = (x0 + 1) - ix
    wy0 = (y0 + 1) - iy
    x0i = x0.astype(np.int64)
    y0i = y0.astype(np.int64)
    x1i = x0i + 1
    y1i = y0i + 1
    # taps that fall outside contribute zero (only possible at the exact upper border,
    # where their weight is zero anyway); clamp the index for the gather
    vx1 = (x1i <= W - 1)
    vy1 = (y1i <= H - 1)
    x1c = np.minimum(x1i, W - 1)
    y1c = np.minimum(y1i, H - 1)
    bidx = np.arange(B)[:, None, None]
    out = np.empty((B, C) + gx.shape[1:], dtype=dt)
    for c in range(C):
        p = img[:, c]
        v00 = p[bidx, y0i, x0i]
        v01 = p[bidx, y0i, x1c] * vx1
        v10 = p[bidx, y1c, x0i] * vy1
        v11 = p[bidx, y1c, x1c] * (vx1 & vy1)
        out[:, c] = v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1)
    return out


# --------------------------------------------------------------------------------------
# TPS rectification (modules/transformation.py)
# --------------------------------------------------------------------------------------
def _conv_bn_relu(x, sd, conv_name, bn_name, dt, stride=(1, 1), pad=(1, 1), quant=None,
                  quant_weight=True, act=True):
    w = sd[conv_name + ".weight"].astype(dt)
    if quant and quant_weight:
        w = _q(w, quant)
    y = conv2d(x, w, stride, pad)
    y = batchnorm_eval(y, sd[bn_name + ".weight"], sd[bn_name + ".bias"],
                       sd[bn_name + ".running_mean"], sd[bn_name + ".running_var"])
    return relu(y) if act else y


def localization_network(x, sd, dt=np.float32, quant=None):
    """LocalizationNetwork.forward (transformation.py:74-82): conv stack :48-58, fc1 :60, fc2 :61.
    Returns C' [B, F, 2]."""
    p = "Transformation.LocalizationNetwork."
    h = x.astype(dt)
    for i in range(4):
        # conv.0 has C_in = I_channel_num and runs in fp32 SIMT on the GPU: weights not quantised
        h = _conv_bn_relu(h, sd, f"{p}conv.{4 * i}", f"{p}conv.{4 * i + 1}", dt, quant=quant,
                          quant_weight=(i > 0))
        if i < 3:
            h = maxpool2d(h)
            h = _q(h, quant)           # stored bf16 (NHWC) between LocNet stages
        else:
            h = _q(h, quant)           # conv.12 output stored bf16, then averaged in fp32
            h = h.mean(axis=(2, 3))    # AdaptiveAvgPool2d(1) + view(B, -1)
    h = relu(linear(h, sd[p + "localization_fc1.0.weight"], sd[p + "localization_fc1.0.bias"]))
    c = linear(h, sd[p + "localization_fc2.weight"], sd[p + "localization_fc2.bias"])
    F = c.shape[1] // 2
    return c.reshape(-1, F, 2)


def build_P_prime(c_prime, sd):
    """GridGenerator.build_P_prime (transformation.py:154-163): T = inv_delta_C [C';0], P' = P_hat T."""
    dt = c_prime.dtype
    inv_delta_C = sd["Transformation.GridGenerator.inv_delta_C"].astype(dt)
    P_hat = sd["Transformation.GridGenerator.P_hat"].astype(dt)
    B = c_prime.shape[0]
    cz = np.concatenate([c_prime, np.zeros((B, 3, 2), dt)], axis=1)
    T = np.einsum("ij,bjk->bik", inv_delta_C, cz)
    return np.einsum("nj,bjk->bnk", P_hat, T)


def tps_forward(x, sd, H, W, dt=np.float32, quant=None):
    """TPS_SpatialTransformerNetwork.forward (transformation.py:28-38)."""
    c_prime = localization_network(x, sd, dt, quant)
    grid = build_P_prime(c_prime, sd).reshape(-1, H, W, 2)
    return c_prime, grid, grid_sample_bilinear_border(x.astype(dt), grid)


# --------------------------------------------------------------------------------------
# ResNet feature extractor (modules/feature_extraction.py)
# --------------------------------------------------------------------------------------
def _bn_fold(sd, name, dt):
    scale = sd[name + ".weight"].astype(dt) / np.sqrt(sd[name + ".running_var"].astype(dt) + dt.type(BN_EPS))
    return scale, sd[name + ".bias"].astype(dt) - sd[name + ".running_mean"].astype(dt) * scale


def _basic_block(x, sd, prefix, dt, quant, trace=None):
    """BasicBlock.forward (feature_extraction.py:32-47)."""
    y = _basic_block_body(x, sd, prefix, dt, quant, trace)
    if trace is not None:
        trace.append((prefix.split("ConvNet.")[-1], y))
    return y


def _basic_block_body(x, sd, prefix, dt, quant, trace):
    out = _q(_conv_bn_relu(x, sd, prefix + ".conv1", prefix + ".bn1", dt, quant=quant), quant)
    if trace is not None:
        trace.append((prefix.split("ConvNet.")[-1] + ".conv1", out))
    if (prefix + ".downsample.0.weight") in sd and quant:
        # CUDA bf16 mode: conv2/bn2 + downsample conv/bn + add as ONE GEMM whose bf16 weights carry the folded
        # BatchNorm scales (both branches share the fp32 accumulator; no intermediate rounding of the residual)
        s2, b2 = _bn_fold(sd, prefix + ".bn2", dt)
        sdn, bd = _bn_fold(sd, prefix + ".downsample.1", dt)
        w2 = _q(sd[prefix + ".conv2.weight"].astype(dt) * s2[:, None, None, None], quant)
        wd = _q(sd[prefix + ".downsample.0.weight"].astype(dt) * sdn[:, None, None, None], quant)
        y = conv2d(out, w2, (1, 1), (1, 1)) + conv2d(x, wd, (1, 1), (0, 0)) + (b2 + bd)[None, :, None, None]
        return _q(relu(y), quant)
    out = _conv_bn_relu(out, sd, prefix + ".conv2", prefix + ".bn2", dt, quant=quant, act=False)
    if (prefix + ".downsample.0.weight") in sd:
        res = _conv_bn_relu(x, sd, prefix + ".downsample.0", prefix + ".downsample.1", dt,
                            pad=(0, 0), quant=quant, act=False)
    else:
        res = x
    return _q(relu(out + res), quant)


def resnet_forward(x, sd, dt=np.float32, quant=None, taps: Optional[dict] = None, trace: Optional[list] = None):
    """ResNet.forward (feature_extraction.py:109-150). x [B,C,H,W] -> [B,512,H',W'].
    ``trace``: list that receives (layer name, output) of every conv group / block / pool in execution order
    (the per-layer parity tests compare them with the CUDA path's launches one to one)."""
    n = "FeatureExtraction.ConvNet."

    def tr(name, v):
        if trace is not None:
            trace.append((name, v))
        return v
    # conv0_1: C_in = input_channel, fp32 SIMT on the GPU (weights not quantised), output bf16
    h = tr("conv0_1", _q(_conv_bn_relu(x.astype(dt), sd, n + "conv0_1", n + "bn0_1", dt, quant=quant,
                                       quant_weight=False), quant))
    h = tr("conv0_2", _q(_conv_bn_relu(h, sd, n + "conv0_2", n + "bn0_2", dt, quant=quant), quant))
    h = tr("maxpool1", maxpool2d(h))
    h = _basic_block(h, sd, n + "layer1.0", dt, quant, trace)
    h = tr("conv1", _q(_conv_bn_relu(h, sd, n + "conv1", n + "bn1", dt, quant=quant), quant))
    h = tr("maxpool2", maxpool2d(h))
    for b in range(2):
        h = _basic_block(h, sd, f"{n}layer2.{b}", dt, quant, trace)
    h = tr("conv2", _q(_conv_bn_relu(h, sd, n + "conv2", n + "bn2", dt, quant=quant), quant))
    h = tr("maxpool3", maxpool2d(h, (2, 2), (2, 1), (0, 1)))
    if taps is not None:
        taps["after_maxpool3"] = h
    for b in range(5):
        h = _basic_block(h, sd, f"{n}layer3.{b}", dt, quant, trace)
    h = tr("conv3", _q(_conv_bn_relu(h, sd, n + "conv3", n + "bn3", dt, quant=quant), quant))
    for b in range(3):
        h = _basic_block(h, sd, f"{n}layer4.{b}", dt, quant, trace)
    if taps is not None:
        taps["after_layer4"] = h
    h = tr("conv4_1", _q(_conv_bn_relu(h, sd, n + "conv4_1", n + "bn4_1", dt, stride=(2, 1), pad=(0, 1),
                                       quant=quant), quant))
    h = tr("conv4_2", _q(_conv_bn_relu(h, sd, n + "conv4_2", n + "bn4_2", dt, pad=(0, 0), quant=quant), quant))
    return h


def column_features(feat):
    """model.py:92-93: permute(0,3,1,2) -> AdaptiveAvgPool2d((None,1)) -> squeeze(3): [B,W,C]."""
    return feat.transpose(0, 3, 1, 2).mean(axis=3)


# --------------------------------------------------------------------------------------
# semantic vectors (modules/semantic_vectors.py:33-49)
# --------------------------------------------------------------------------------------
def linear_embedding(overlap, scene, ious, table, mode="overlap", dt=np.float32):
    """Linear_Embedding.forward: ids arrive as float and are cast with .long()."""
    table = table.astype(dt)
    if mode == "overlap":
        return table[overlap.astype(np.int64)]
    if mode == "scene":
        w = softmax(ious.astype(dt), axis=1)
        return table[scene.astype(np.int64)] * w[..., None]
    raise ValueError("SEMANTIC_VECTOR %r is not runnable in the reference either" % mode)


# --------------------------------------------------------------------------------------
# BiLSTM encoder (modules/encoders.py:10-30)
# --------------------------------------------------------------------------------------
def _lstm_direction(gx, w_hh, reverse, quant=None):
    """One direction of nn.LSTM given precomputed input gates gx [B,T,4H] (bias included).
    ``quant='bf16'``: W_hh and the fed-back h_t are rounded to bf16 (the tensor-core recurrence of
    the CUDA bf16 mode keeps h_t in bf16; the cell state and accumulation stay fp32)."""
    B, T, G = gx.shape
    Hd = G // 4
    h = np.zeros((B, Hd), gx.dtype)
    c = np.zeros((B, Hd), gx.dtype)
    out = np.empty((B, T, Hd), gx.dtype)
    w_hh_t = _q(w_hh.astype(gx.dtype), quant).T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = gx[:, t] + _q(h, quant) @ w_hh_t
        i = sigmoid(g[:, :Hd]); f = sigmoid(g[:, Hd:2 * Hd])
        gg = np.tanh(g[:, 2 * Hd:3 * Hd]); o = sigmoid(g[:, 3 * Hd:])
        c = f * c + i * gg
        h = o * np.tanh(c)
        out[:, t] = h
    return out


def bidirectional_lstm(x, sd, prefix, dt=np.float32, quant=None, gates_bf16=True):
    """BidirectionalLSTM.forward (encoders.py:17-30): nn.LSTM(bidirectional, batch_first) + Linear."""
    x = x.astype(dt)
    outs = []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih = _q(sd[f"{prefix}.rnn.weight_ih_l0{sfx}"].astype(dt), quant)
        bias = (sd[f"{prefix}.rnn.bias_ih_l0{sfx}"].astype(dt) + sd[f"{prefix}.rnn.bias_hh_l0{sfx}"].astype(dt))
        gx = _q(x, quant) @ w_ih.T + bias
        if quant and gates_bf16:
            gx = _q(gx, quant)       # config.GATES_X_BF16: the input projection is stored as bf16
        outs.append(_lstm_direction(gx, sd[f"{prefix}.rnn.weight_hh_l0{sfx}"], rev, quant))
    rec = _q(np.concatenate(outs, axis=2), quant)   # [fwd | bwd], stored bf16 for the Linear GEMM
    w = _q(sd[prefix + ".linear.weight"].astype(dt), quant)
    return rec @ w.T + sd[prefix + ".linear.bias"].astype(dt)


# --------------------------------------------------------------------------------------
# attention decoder (modules/decoders.py:10-101)
# --------------------------------------------------------------------------------------
def attention_decode(enc, sd, steps, num_classes, text=None, dt=np.float32, quant=None,
                     return_alpha=False):
    """Attention.forward (decoders.py:26-64) with AttentionCell.forward (:79-101).

    ``text is None`` -> greedy branch (:52-62, is_train=False): token_0 = [GO] = 0, then argmax
    feedback (np.argmax returns the lowest index on ties, like torch.max on CPU).
    ``text`` given   -> teacher-forced branch (:41-50, is_train=True) using text[:, i].
    The ``semantics`` argument of the reference is accepted and ignored there (:79-101 never
    reads it), so it does not appear here.
    Returns logits [B, steps, num_classes] (the reference calls them probs; no softmax applied).
    """
    p = "decoder.attention_cell."
    enc = enc.astype(dt)
    B, T, Hd = enc.shape
    w_i2h = _q(sd[p + "i2h.weight"].astype(dt), quant)
    # quant='bf16' restates the tensor-core decode of the CUDA bf16 mode: enc, proj, h and ctx are bf16
    # operands of the GEMMs (weights bf16 as well), accumulation / cell state / logits stay fp32
    enc = _q(enc, quant)
    proj = _q(enc @ w_i2h.T, quant)                               # :81 (loop invariant)
    w_h2h = _q(sd[p + "h2h.weight"].astype(dt), quant); b_h2h = sd[p + "h2h.bias"].astype(dt)
    w_score = sd[p + "score.weight"].astype(dt)[0]
    w_ih = sd[p + "rnn.weight_ih"].astype(dt); w_hh = _q(sd[p + "rnn.weight_hh"].astype(dt), quant)
    w_ctx = _q(w_ih[:, :Hd], quant)
    bias = sd[p + "rnn.bias_ih"].astype(dt) + sd[p + "rnn.bias_hh"].astype(dt)
    w_gen = _q(sd["decoder.generator.weight"].astype(dt), quant); b_gen = sd["decoder.generator.bias"].astype(dt)
    h = np.zeros((B, Hd), dt); c = np.zeros((B, Hd), dt)
    tok = np.zeros(B, np.int64)                                   # [GO]
    logits = np.zeros((B, steps, num_classes), dt)
    alphas = np.zeros((B, steps, T), dt)
    for i in range(steps):
        if text is not None:
            tok = text[:, i].astype(np.int64)
        hq = _q(h, quant)
        hp = hq @ w_h2h.T + b_h2h                                 # :82
        e = np.tanh(proj + hp[:, None, :]) @ w_score              # :85  [B,T]
        alpha = softmax(e, axis=1)                                # :87
        ctx = np.einsum("bt,btc->bc", alpha, enc)                 # :88
        g = _q(ctx, quant) @ w_ctx.T + w_ih[:, Hd + tok].T + hq @ w_hh.T + bias   # :90,:99 (one-hot = column gather)
        ii = sigmoid(g[:, :Hd]); f = sigmoid(g[:, Hd:2 * Hd])
        gg = np.tanh(g[:, 2 * Hd:3 * Hd]); o = sigmoid(g[:, 3 * Hd:])
        c = f * c + ii * gg
        h = o * np.tanh(c)
        step_logits = _q(h, quant) @ w_gen.T + b_gen              # :59 / :50
        logits[:, i] = step_logits
        alphas[:, i] = alpha
        if text is None:
            tok = step_logits.argmax(axis=1)                      # :61-62
    return (logits, alphas) if return_alpha else logits


# --------------------------------------------------------------------------------------
# whole forward (model.py:85-107)
# --------------------------------------------------------------------------------------
def forward(sd: Dict[str, np.ndarray], image: np.ndarray, overlap=None, scene=None, ious=None,
            text=None, is_train: bool = False, max_len: int = 25, semantic_mode: str = "overlap",
            dtype=np.float32, quant: Optional[str] = None, stages: bool = False):
    """Model.forward for ENCODER='LSTM', DECODER='LSTM'.  Returns logits [B, max_len+1, classes]
    or, with ``stages=True``, a dict of every intermediate the parity tests compare."""
    dt = np.dtype(dtype)
    B, C, H, W = image.shape
    c_prime, grid, rect = tps_forward(image, sd, H, W, dt, quant)         # model.py:88
    feat = resnet_forward(rect, sd, dt, quant)                            # model.py:91
    vis = column_features(feat)                                           # model.py:92-93
    sem = None
    if overlap is not None:
        sem = linear_embedding(overlap, scene, ious, sd["get_semantic_vectors.embed.weight"],
                               semantic_mode, dt)                         # model.py:96
    enc0 = bidirectional_lstm(vis, sd, "encoder.0", dt, quant)            # model.py:100
    enc0 = _q(enc0, quant)
    enc1 = bidirectional_lstm(enc0, sd, "encoder.1", dt, quant)
    num_classes = sd["decoder.generator.weight"].shape[0]
    logits = attention_decode(enc1, sd, max_len + 1, num_classes,
                              text=(text if is_train else None), dt=dt, quant=quant)  # model.py:105
    if not stages:
        return logits
    return {"c_prime": c_prime, "grid": grid, "rectified": rect, "features": feat, "visual": vis,
            "semantics": sem, "enc0": enc0, "enc1": enc1, "logits": logits,
            "ids": logits.argmax(axis=2)}


def decode_ids(ids: np.ndarray, chars: str):
    """AttnLabelConverter.decode (utils.py:143-149) + the caller's cut at the first '[s]'
    (training_functions.py:188-189: ``pred[:pred.find('[s]')]``; when no '[s]' was emitted
    find() returns -1 and the slice drops the final character - restated as is)."""
    table = ["[GO]", "[s]", "[PAD]"] + list(chars)
    out = []
    for row in ids:
        s = "".join(table[int(i)] for i in row)
        k = s.find("[s]")
        out.append(s[:k])   # k == -1 drops the last character, exactly like the reference
    return out
