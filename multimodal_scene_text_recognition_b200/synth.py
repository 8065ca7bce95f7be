"""Seeded synthetic weights and crops for the recognition forward pass.

The reference ships no checkpoint and no fixtures (README.md:7 points at a Dropbox file that
is not in the tree), and its default random init is degenerate for parity work
(`localization_fc2.weight` is zero-filled at transformation.py:64, eval-mode
BatchNorm with running stats 0/1 collapses the ResNet output to ~1e-4, every crop decodes to
the same string).  This module is the *recipe* both sides of every parity test use:

* ``state_dict_spec(geom)``   – the 253 tensor names/shapes of the LSTM/LSTM model
                                (model.py:34-82; names listed in SURVEY.md 8b).
* ``make_state_dict(seed)``   – numpy-only, He-scaled conv weights, non-trivial BN statistics,
                                non-zero ``localization_fc2.weight`` so the TPS grid depends on
                                the crop and the border clamp of grid_sample is exercised.
* ``make_crops(seed, B)``     – low-frequency field + noise in [0,1] (coco_dataset.py:40-41
                                produces ToTensor() images in [0,1], no mean/std normalisation).
* ``make_semantics(seed, B)`` – COCO-Text-style object id vectors (coco_dataset.py:245-257).

Everything is generated with ``numpy.random.Generator(PCG64(seed))`` so it does not depend on
torch's RNG stream and can be regenerated bit-identically on the GPU box.
"""
from __future__ import annotations

import dataclasses
import string
from typing import Dict, List, Tuple

import numpy as np


@dataclasses.dataclass(frozen=True)
class Geometry:
    """Shape constants of the model (model.py:21-32, config.py:19-31)."""

    in_ch: int = 1
    H: int = 32
    W: int = 100
    F: int = 20
    max_len: int = 25          # config.MAX_TEXT_LENGTH
    hidden: int = 256          # model.py:30
    out_ch: int = 512          # model.py:27
    embed_dim: int = 256       # config.EMBED_DIM
    num_obj_classes: int = 2000  # semantic_vectors.py:19
    chars: str = string.printable[:-6]  # config.py:20

    @property
    def num_classes(self) -> int:  # utils.py:108-111: [GO], [s], [PAD] + chars
        return len(self.chars) + 3

    @property
    def steps(self) -> int:  # decoders.py:34
        return self.max_len + 1

    @property
    def feat_hw(self) -> Tuple[int, int]:
        """Spatial size of the ResNet output (feature_extraction.py:109-150)."""
        h, w = self.H // 2, self.W // 2          # maxpool1
        h, w = h // 2, w // 2                    # maxpool2
        h, w = h // 2, w + 1                     # maxpool3 k2 s(2,1) p(0,1)
        h, w = (h - 2) // 2 + 1, w + 1           # conv4_1 k2 s(2,1) p(0,1)
        h, w = h - 1, w - 1                      # conv4_2 k2 s1 p0
        return h, w

    @property
    def T(self) -> int:
        return self.feat_hw[1]


CFG_A = Geometry()
CFG_S = Geometry(in_ch=3, H=64, W=256, max_len=50)


def _bn(prefix: str, c: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (prefix + ".weight", (c,), "bn_weight"),
        (prefix + ".bias", (c,), "bn_bias"),
        (prefix + ".running_mean", (c,), "bn_mean"),
        (prefix + ".running_var", (c,), "bn_var"),
        (prefix + ".num_batches_tracked", (), "bn_count"),
    ]


def state_dict_spec(geom: Geometry = CFG_A) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor of ``Model().state_dict()`` in LSTM/LSTM config."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    loc = "Transformation.LocalizationNetwork."
    chans = [geom.in_ch, 64, 128, 256, 512]
    for i in range(4):  # transformation.py:48-58
        spec.append((f"{loc}conv.{4 * i}.weight", (chans[i + 1], chans[i], 3, 3), "conv"))
        spec += _bn(f"{loc}conv.{4 * i + 1}", chans[i + 1])
    spec += [
        (loc + "localization_fc1.0.weight", (256, 512), "linear"),
        (loc + "localization_fc1.0.bias", (256,), "linear_bias"),
        (loc + "localization_fc2.weight", (geom.F * 2, 256), "loc_fc2_weight"),
        (loc + "localization_fc2.bias", (geom.F * 2,), "loc_fc2_bias"),
        ("Transformation.GridGenerator.inv_delta_C", (geom.F + 3, geom.F + 3), "tps_inv_delta_C"),
        ("Transformation.GridGenerator.P_hat", (geom.H * geom.W, geom.F + 3), "tps_P_hat"),
    ]
    net = "FeatureExtraction.ConvNet."
    oc = geom.out_ch
    blocks = [oc // 4, oc // 2, oc, oc]
    spec.append((net + "conv0_1.weight", (oc // 16, geom.in_ch, 3, 3), "conv"))
    spec += _bn(net + "bn0_1", oc // 16)
    spec.append((net + "conv0_2.weight", (oc // 8, oc // 16, 3, 3), "conv"))
    spec += _bn(net + "bn0_2", oc // 8)
    inpl = oc // 8
    nblocks = [1, 2, 5, 3]  # feature_extraction.py:9

    def layer(name: str, planes: int, n: int) -> None:
        nonlocal inpl
        for b in range(n):
            cin = inpl if b == 0 else planes
            spec.append((f"{net}{name}.{b}.conv1.weight", (planes, cin, 3, 3), "conv"))
            spec.extend(_bn(f"{net}{name}.{b}.bn1", planes))
            spec.append((f"{net}{name}.{b}.conv2.weight", (planes, planes, 3, 3), "conv"))
            spec.extend(_bn(f"{net}{name}.{b}.bn2", planes))
            if b == 0 and inpl != planes:  # feature_extraction.py:94
                spec.append((f"{net}{name}.0.downsample.0.weight", (planes, inpl, 1, 1), "conv"))
                spec.extend(_bn(f"{net}{name}.0.downsample.1", planes))
        inpl = planes

    layer("layer1", blocks[0], nblocks[0])
    spec.append((net + "conv1.weight", (blocks[0], blocks[0], 3, 3), "conv"))
    spec += _bn(net + "bn1", blocks[0])
    layer("layer2", blocks[1], nblocks[1])
    spec.append((net + "conv2.weight", (blocks[1], blocks[1], 3, 3), "conv"))
    spec += _bn(net + "bn2", blocks[1])
    layer("layer3", blocks[2], nblocks[2])
    spec.append((net + "conv3.weight", (blocks[2], blocks[2], 3, 3), "conv"))
    spec += _bn(net + "bn3", blocks[2])
    layer("layer4", blocks[3], nblocks[3])
    spec.append((net + "conv4_1.weight", (blocks[3], blocks[3], 2, 2), "conv"))
    spec += _bn(net + "bn4_1", blocks[3])
    spec.append((net + "conv4_2.weight", (blocks[3], blocks[3], 2, 2), "conv"))
    spec += _bn(net + "bn4_2", blocks[3])

    spec.append(("get_semantic_vectors.embed.weight", (geom.num_obj_classes, geom.embed_dim), "embedding"))
    hd = geom.hidden
    for li, insz in enumerate([oc, hd]):  # model.py:64-66
        for sfx in ["", "_reverse"]:
            spec += [
                (f"encoder.{li}.rnn.weight_ih_l0{sfx}", (4 * hd, insz), "lstm"),
                (f"encoder.{li}.rnn.weight_hh_l0{sfx}", (4 * hd, hd), "lstm"),
                (f"encoder.{li}.rnn.bias_ih_l0{sfx}", (4 * hd,), "lstm"),
                (f"encoder.{li}.rnn.bias_hh_l0{sfx}", (4 * hd,), "lstm"),
            ]
        spec += [
            (f"encoder.{li}.linear.weight", (hd, 2 * hd), "linear"),
            (f"encoder.{li}.linear.bias", (hd,), "linear_bias"),
        ]
    nc = geom.num_classes
    dec = "decoder.attention_cell."
    spec += [
        (dec + "i2h.weight", (hd, hd), "linear"),
        (dec + "h2h.weight", (hd, hd), "linear"),
        (dec + "h2h.bias", (hd,), "linear_bias"),
        (dec + "score.weight", (1, hd), "score"),
        (dec + "rnn.weight_ih", (4 * hd, hd + nc), "lstm"),
        (dec + "rnn.weight_hh", (4 * hd, hd), "lstm"),
        (dec + "rnn.bias_ih", (4 * hd,), "lstm"),
        (dec + "rnn.bias_hh", (4 * hd,), "lstm"),
        ("decoder.generator.weight", (nc, hd), "generator"),
        ("decoder.generator.bias", (nc,), "generator_bias"),
    ]
    return spec


def tps_constants(F: int, H: int, W: int) -> Tuple[np.ndarray, np.ndarray]:
    """fp64 construction of ``inv_delta_C`` [F+3,F+3] and ``P_hat`` [H*W,F+3].

    Restates GridGenerator._build_C/_build_inv_delta_C/_build_P/_build_P_hat
    (transformation.py:103-152): fiducials on the top and bottom edges, U(r) = r^2 log r,
    pixel-centre target grid with x fastest, eps 1e-6 inside the log.  The reference casts the
    fp64 result to fp32 buffers (transformation.py:97-98); callers do the same.
    """
    half = F // 2
    cx = np.linspace(-1.0, 1.0, half)
    C = np.concatenate([np.stack([cx, -np.ones(half)], 1), np.stack([cx, np.ones(half)], 1)], 0)
    d = np.linalg.norm(C[:, None, :] - C[None, :, :], axis=2)
    np.fill_diagonal(d, 1.0)
    hat_C = d ** 2 * np.log(d)
    delta_C = np.concatenate([
        np.concatenate([np.ones((F, 1)), C, hat_C], 1),
        np.concatenate([np.zeros((2, 3)), C.T], 1),
        np.concatenate([np.zeros((1, 3)), np.ones((1, F))], 1)], 0)
    inv_delta_C = np.linalg.inv(delta_C)
    gx = (np.arange(-W, W, 2) + 1.0) / W
    gy = (np.arange(-H, H, 2) + 1.0) / H
    P = np.stack(np.meshgrid(gx, gy), axis=2).reshape(-1, 2)
    r = np.linalg.norm(P[:, None, :] - C[None, :, :], ord=2, axis=2)
    rbf = np.square(r) * np.log(r + 1e-6)
    P_hat = np.concatenate([np.ones((P.shape[0], 1)), P, rbf], 1)
    return inv_delta_C, P_hat


def rare_fiducial_bias(F: int) -> np.ndarray:
    """Initial ``localization_fc2.bias`` (RARE Fig. 6a), transformation.py:66-72."""
    half = F // 2
    cx = np.linspace(-1.0, 1.0, half)
    top = np.stack([cx, np.linspace(0.0, -1.0, half)], 1)
    bot = np.stack([cx, np.linspace(1.0, 0.0, half)], 1)
    return np.concatenate([top, bot], 0).reshape(-1)


# Gains over PyTorch's default U(-1/sqrt(fan), 1/sqrt(fan)) for the sequence model (see _sequence_gain).
SEQUENCE_GAINS = {"enc_w_ih": 3.0, "enc_linear": 8.0, "dec_w_ih": 3.0, "dec_i2h": 8.0, "dec_score": 8.0}


def _sequence_gain(name: str) -> float:
    """Gains over PyTorch's default U(-1/sqrt(fan), 1/sqrt(fan)) for the sequence model.

    With the default scale the two BiLSTM layers shrink the crop-dependent signal to ~4e-3 and
    every crop decodes to the same string.  These gains were tuned once against the reference
    modules (B=128, seed 999/123): 59/128 distinct strings, 29 distinct characters, [s] emitted
    by 23 % of the crops, median top-1/top-2 margin 0.36 x logit std.
    """
    g = SEQUENCE_GAINS
    if name.startswith("encoder.") and "weight_ih" in name:
        return g["enc_w_ih"]
    if name.startswith("encoder.") and name.endswith("linear.weight"):
        return g["enc_linear"]
    if name == "decoder.attention_cell.rnn.weight_ih":
        return g["dec_w_ih"]
    if name == "decoder.attention_cell.i2h.weight":
        return g["dec_i2h"]
    if name == "decoder.attention_cell.score.weight":
        return g["dec_score"]
    return 1.0


def make_state_dict(seed: int = 999, geom: Geometry = CFG_A) -> Dict[str, np.ndarray]:
    """Seeded, well-conditioned synthetic weights (fp32 numpy; ``num_batches_tracked`` int64).

    Seed default 999 = config.RANDOM_SEED (config.py:16).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    inv_delta_C, P_hat = tps_constants(geom.F, geom.H, geom.W)
    out: Dict[str, np.ndarray] = {}
    hd = geom.hidden
    for name, shape, kind in state_dict_spec(geom):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)
            if fan_in > 9:
                # zero-mean filters: the large positive mean of post-ReLU activations would
                # otherwise turn into an input-independent per-channel offset that swamps the
                # crop-dependent signal after a few layers (probe: across-batch std / std
                # 0.14 -> 0.45 at the ResNet output).
                v = v - v.mean(axis=(1, 2, 3), keepdims=True)
        elif kind == "bn_weight":
            v = rng.uniform(0.8, 1.2, shape)
            # residual branches add to the skip path: damp them so the variance does not
            # double per BasicBlock (feature_extraction.py:39-45)
            if name.endswith(".bn2.weight") and ".layer" in name:
                v = v * 0.5
        elif kind == "bn_bias":
            v = rng.standard_normal(shape) * 0.1
        elif kind == "bn_mean":
            v = rng.standard_normal(shape) * 0.1
        elif kind == "bn_var":
            v = rng.uniform(0.8, 1.25, shape)
        elif kind == "bn_count":
            out[name] = np.array(1, dtype=np.int64)
            continue
        elif kind in ("linear", "lstm", "score"):
            fan_in = shape[-1] if len(shape) > 1 else hd
            k = 1.0 / np.sqrt(hd if kind == "lstm" else fan_in)
            v = rng.uniform(-k, k, shape) * _sequence_gain(name)
        elif kind == "linear_bias":
            v = rng.uniform(-0.05, 0.05, shape)
        elif kind == "loc_fc2_weight":
            v = rng.standard_normal(shape) * 0.02
        elif kind == "loc_fc2_bias":
            v = rare_fiducial_bias(geom.F)
        elif kind == "tps_inv_delta_C":
            v = inv_delta_C
        elif kind == "tps_P_hat":
            v = P_hat
        elif kind == "embedding":
            v = rng.standard_normal(shape)  # nn.Embedding default N(0,1)
        elif kind == "generator":
            v = rng.uniform(-0.25, 0.25, shape)
        elif kind == "generator_bias":
            v = rng.uniform(-0.1, 0.1, shape)
        else:  # pragma: no cover
            raise KeyError(kind)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def _upsample_bilinear(coarse: np.ndarray, H: int, W: int) -> np.ndarray:
    gh, gw = coarse.shape[2:]
    ys = np.linspace(0, gh - 1, H)
    xs = np.linspace(0, gw - 1, W)
    y0 = np.floor(ys).astype(int); y1 = np.minimum(y0 + 1, gh - 1); wy = (ys - y0)[None, None, :, None]
    x0 = np.floor(xs).astype(int); x1 = np.minimum(x0 + 1, gw - 1); wx = (xs - x0)[None, None, None, :]
    top = coarse[:, :, y0][:, :, :, x0] * (1 - wx) + coarse[:, :, y0][:, :, :, x1] * wx
    bot = coarse[:, :, y1][:, :, :, x0] * (1 - wx) + coarse[:, :, y1][:, :, :, x1] * wx
    return top * (1 - wy) + bot * wy


def make_crops(seed: int, batch: int, geom: Geometry = CFG_A, stroke: float = 0.3,
               noise: float = 0.02) -> np.ndarray:
    """[B, C, H, W] fp32 in [0,1]: a coarse random field (background / illumination), a
    stroke-scale field (~4 px features, text-like) and a little per-pixel noise.

    The per-pixel noise amplitude matters for parity: the reference computes the TPS grid in
    fp32 with cancellation (~1e-5 normalised), and the sampled image inherits
    that error times the local image gradient.  With 0.15 noise two correct fp32
    implementations already differ by 4e-4 in the logits; 0.02 keeps fp32 noise below the
    north star's 1e-4 logits tolerance.  ``noise=1.0, stroke=0`` gives the pure-U[0,1] worst case.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (batch, geom.in_ch)
    low = _upsample_bilinear(rng.uniform(0.0, 1.0, shape + (4, 13)), geom.H, geom.W)
    mid = _upsample_bilinear(rng.uniform(0.0, 1.0, shape + (geom.H // 4, geom.W // 4)), geom.H, geom.W)
    img = low * (1.0 - stroke - noise) + mid * stroke + noise * rng.uniform(0.0, 1.0, low.shape)
    return np.ascontiguousarray(np.clip(img, 0.0, 1.0), dtype=np.float32)   # C order (fancy indexing above is not)


def make_semantics(seed: int, batch: int, n_overlap: int = 15, n_scene: int = 52,
                   n_classes: int = 1489) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(overlap[B,15], scene[B,52], ious[B,52]) as float32, like coco_dataset.py:245-257.

    ids are class+1 with 0 = padding (coco_dataset.py:278); the loaders fill ``ious`` with the
    constant -1000 (coco_dataset.py:247).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    overlap = np.zeros((batch, n_overlap), np.float32)
    scene = np.zeros((batch, n_scene), np.float32)
    for b in range(batch):
        k = int(rng.integers(0, n_overlap + 1))
        overlap[b, :k] = rng.choice(n_classes, size=k, replace=False) + 1
        k = int(rng.integers(0, n_scene + 1))
        scene[b, :k] = rng.choice(n_classes, size=k, replace=False) + 1
    ious = np.full((batch, n_scene), -1000.0, np.float32)
    return overlap, scene, ious


# ---- transformer family (reference config.py:27-31 defaults) ---------------------------------------
def positional_table(d_model: int, max_len: int = 26) -> np.ndarray:
    """PositionalEncoding's buffer ``pe`` [max_len, 1, d_model] (encoders.py:151-159), computed in fp32 like torch."""
    position = np.arange(max_len, dtype=np.float32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2).astype(np.float32) * np.float32(-np.log(10000.0) / d_model)).astype(np.float32)
    pe = np.zeros((max_len, d_model), np.float32)
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe[:, None, :]


def transformer_spec(decoder: str = "Transformer", hidden_dim: int = 512, embed_dim: int = 256, ff: int = 2048,
                     num_classes: int = 97, layers: int = 6, pre_encoder_mlp: bool = False,
                     pre_decoder_mlp: bool = False, cls_decoder_init: bool = False,
                     post_decoder_mlp: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) of the encoder / decoder tensors of the Transformer configurations (TF_Encoder with
    TF_Decoder: 462 tensors with the rest of the model, with Linear_Decoder: 314)."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def mha(p: str, d: int) -> None:
        spec.extend([(p + "in_proj_weight", (3 * d, d), "xavier"), (p + "in_proj_bias", (3 * d,), "bias"),
                     (p + "out_proj.weight", (d, d), "lin"), (p + "out_proj.bias", (d,), "bias")])

    def norm(p: str, d: int) -> None:
        spec.extend([(p + "weight", (d,), "ln_w"), (p + "bias", (d,), "ln_b")])

    def enc_layer(p: str) -> None:
        mha(p + "self_attn.", hidden_dim)
        spec.extend([(p + "linear1.weight", (ff, hidden_dim), "lin"), (p + "linear1.bias", (ff,), "bias"),
                     (p + "linear2.weight", (hidden_dim, ff), "lin"), (p + "linear2.bias", (hidden_dim,), "bias")])
        norm(p + "norm1.", hidden_dim); norm(p + "norm2.", hidden_dim)

    def dec_layer(p: str) -> None:
        mha(p + "self_attn.", embed_dim); mha(p + "multihead_attn.", embed_dim)
        spec.extend([(p + "linear1.weight", (ff, embed_dim), "lin"), (p + "linear1.bias", (ff,), "bias"),
                     (p + "linear2.weight", (embed_dim, ff), "lin"), (p + "linear2.bias", (embed_dim,), "bias"),
                     (p + "semantic_to_emb.weight", (embed_dim, hidden_dim), "lin"), (p + "semantic_to_emb.bias", (embed_dim,), "bias")])
        norm(p + "norm1.", embed_dim); norm(p + "norm2.", embed_dim); norm(p + "norm3.", embed_dim)

    spec.append(("encoder.pos_encoder.pe", (26, 1, hidden_dim), "pe"))
    enc_layer("encoder.encoder_layer.")
    norm("encoder.layer_norm.", hidden_dim)
    for i in range(layers):
        enc_layer("encoder.encoder.layers.%d." % i)
    spec.extend([("encoder.encoder.norm.weight", (hidden_dim,), "alias:encoder.layer_norm.weight"),
                 ("encoder.encoder.norm.bias", (hidden_dim,), "alias:encoder.layer_norm.bias")])

    def mlp(p: str, dims) -> None:      # MLP (encoders.py:200-243): layers.fc{i}
        for i in range(len(dims) - 1):
            spec.extend([("%slayers.fc%d.weight" % (p, i), (dims[i + 1], dims[i]), "mlp_w"),
                         ("%slayers.fc%d.bias" % (p, i), (dims[i + 1],), "bias")])
    if pre_encoder_mlp:                 # encoders.py:83-93
        mlp("encoder.sem_relevence_mlp.", (hidden_dim + embed_dim, hidden_dim, hidden_dim, 1))
        mlp("encoder.combine_mlp.", (hidden_dim + embed_dim, hidden_dim, hidden_dim, hidden_dim))
        spec.append(("encoder.emb_to_hid.weight", (hidden_dim, embed_dim), "lin"))
    if decoder == "Linear":
        spec.extend([("decoder.linear_decoder.weight", (num_classes, hidden_dim), "classes"),
                     ("decoder.linear_decoder.bias", (num_classes,), "bias")])
        return spec
    dec_layer("decoder.decoder_layer.")
    norm("decoder.layer_norm.", embed_dim)
    for i in range(layers):
        dec_layer("decoder.decoder.layers.%d." % i)
    spec.extend([("decoder.decoder.norm.weight", (embed_dim,), "alias:decoder.layer_norm.weight"),
                 ("decoder.decoder.norm.bias", (embed_dim,), "alias:decoder.layer_norm.bias"),
                 ("decoder.pos_encoder.pe", (26, 1, embed_dim), "pe"),
                 ("decoder.hid_to_emb.weight", (embed_dim, hidden_dim), "lin"), ("decoder.hid_to_emb.bias", (embed_dim,), "bias"),
                 ("decoder.emb.weight", (num_classes, embed_dim), "embedding"),
                 ("decoder.emb_to_classes.weight", (num_classes, embed_dim), "classes"),
                 ("decoder.emb_to_classes.bias", (num_classes,), "bias")])
    if pre_decoder_mlp:                 # decoders.py:122-124
        mlp("decoder.relevant_mlp.", (2 * embed_dim, embed_dim, embed_dim, 1))
        mlp("decoder.combine_mlp.", (2 * embed_dim, embed_dim, embed_dim))
    if cls_decoder_init:                # decoders.py:126-128
        mlp("decoder.sem_cls_mlp.", (2 * embed_dim, embed_dim, embed_dim, 1))
    if post_decoder_mlp:                # decoders.py:132-135
        mlp("decoder.post_decoder_mlp.", (2 * num_classes, num_classes, num_classes, 1))
        mlp("decoder.post_deocer_combine_mlp.", (2 * num_classes, num_classes, num_classes, num_classes))
        spec.extend([("decoder.sem_to_classes.weight", (num_classes, embed_dim), "lin"),
                     ("decoder.sem_to_classes.bias", (num_classes,), "bias")])
    return spec


# Gains over the default scales, tuned once against the reference's modules (B=32, seeds 999/123): sharper attention
# (q and k rows of every in_proj), a stronger encoder feed-forward branch and memory attention, a weaker token
# embedding - the decode then depends on the crop (32/32 distinct strings, 11 distinct characters, median top-1/top-2
# margin 0.19 x logit std) instead of on the token feedback alone (4 distinct strings with every gain at 1).
TRANSFORMER_GAINS = {"classes": 4.0, "embedding": 0.5, "qk": 4.0, "enc_linear2": 4.0, "dec_memory_out": 2.0}


def make_state_dict_transformer(seed: int = 999, decoder: str = "Transformer", geom: Geometry = CFG_A,
                                pre_encoder_mlp: bool = False, pre_decoder_mlp: bool = False, cls_decoder_init: bool = False,
                                post_decoder_mlp: bool = False) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights of the Transformer / Transformer (or Transformer / Linear) configuration: the rectifier,
    ResNet and embedding table of ``make_state_dict`` plus independent random weights for every encoder / decoder layer
    (the reference's default init makes the six layers equal copies, encoders.py:80, and decodes every crop to one
    repeated character).  ``*.norm`` aliases ``layer_norm`` like in the reference's modules."""
    base = make_state_dict(seed, geom)
    out = {k: v for k, v in base.items() if not (k.startswith("encoder.") or k.startswith("decoder."))}
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    for name, shape, kind in transformer_spec(decoder, num_classes=geom.num_classes, pre_encoder_mlp=pre_encoder_mlp,
                                              pre_decoder_mlp=pre_decoder_mlp, cls_decoder_init=cls_decoder_init,
                                              post_decoder_mlp=post_decoder_mlp):
        if kind == "pe":
            v = positional_table(shape[2], shape[0])
        elif kind == "xavier":
            a = np.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-a, a, shape)
            v[:2 * shape[1]] *= TRANSFORMER_GAINS["qk"]
        elif kind == "lin":
            k = 1.0 / np.sqrt(shape[1])
            v = rng.uniform(-k, k, shape)
            if name.startswith("encoder.") and name.endswith("linear2.weight"):
                v = v * TRANSFORMER_GAINS["enc_linear2"]
            if name.endswith("multihead_attn.out_proj.weight"):
                v = v * TRANSFORMER_GAINS["dec_memory_out"]
        elif kind == "mlp_w":
            v = rng.standard_normal(shape) * np.sqrt(2.0 / shape[1])     # the MLP's own init (encoders.py:226)
        elif kind == "classes":
            k = TRANSFORMER_GAINS["classes"] / np.sqrt(shape[1])
            v = rng.uniform(-k, k, shape)
        elif kind == "bias":
            v = rng.uniform(-0.05, 0.05, shape)
        elif kind == "ln_w":
            v = rng.uniform(0.8, 1.2, shape)
        elif kind == "ln_b":
            v = rng.standard_normal(shape) * 0.1
        elif kind == "embedding":
            v = rng.standard_normal(shape) * TRANSFORMER_GAINS["embedding"]
        elif kind.startswith("alias:"):
            v = out[kind[6:]]
        else:  # pragma: no cover
            raise KeyError(kind)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def perturb_parameters_(module, seed: int, scale: float = 0.05) -> None:
    """In place: add seeded noise (numpy PCG64, ``scale`` x standard normal) to every floating-point parameter of a torch
    module, in ``named_parameters()`` order.  Used on both sides of the Oscar_Bert parity fixture: BertModel's default
    init leaves every bias at 0 and every LayerNorm at (1, 0), which would leave those code paths untested."""
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    with torch.no_grad():
        for _name, p in module.named_parameters():
            if p.is_floating_point():
                p.add_(torch.from_numpy((rng.standard_normal(tuple(p.shape)) * scale).astype(np.float32)).to(p.device, p.dtype))
