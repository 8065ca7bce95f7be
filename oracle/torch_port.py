"""CPU ORACLE #2 (test / baseline infrastructure, NOT a product path): the recognition forward
restated with ``torch.nn.functional`` calls on CPU tensors.

Why a second restatement next to ``recognition_oracle.py`` (numpy): the reference's arithmetic is
PyTorch ATen, so timing ATen's own CPU kernels (oneDNN convolutions, native LSTM
cell math, multi-threaded) is the fairest stand-in for "the reference's CPU forward" on a box
where /root/reference does not exist.  ``bench.py`` uses it for ``cpu_baseline`` and for
``--impl reference`` (kind = "port"); ``tests/test_oracle_golden.py`` pins it against the golden
fixtures produced by the reference's own modules.

It follows the same reference lines as the numpy oracle:
  TPS      transformation.py:28-38 (forward), :74-82 (LocalizationNetwork), :154-163 (build_P_prime)
  ResNet   feature_extraction.py:32-47 (BasicBlock), :109-150 (ResNet.forward)
  pooling  model.py:92-93
  BiLSTM   encoders.py:17-30
  decoder  decoders.py:26-64, :79-101; the i2h projection is recomputed EVERY step like the
           reference does (:81) so the timing reflects the reference's cost
  semantic semantic_vectors.py:33-49

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def to_torch(sd: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def _cbr(x, sd, conv, bn, stride=1, pad=1, relu=True):
    y = F.conv2d(x, sd[conv + ".weight"], None, stride, pad)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.1, BN_EPS)
    return F.relu(y) if relu else y


def localization_network(x, sd):
    p = "Transformation.LocalizationNetwork."
    h = x
    for i in range(4):
        h = _cbr(h, sd, f"{p}conv.{4 * i}", f"{p}conv.{4 * i + 1}")
        h = F.max_pool2d(h, 2, 2) if i < 3 else F.adaptive_avg_pool2d(h, 1)
    h = h.view(x.size(0), -1)
    h = F.relu(F.linear(h, sd[p + "localization_fc1.0.weight"], sd[p + "localization_fc1.0.bias"]))
    c = F.linear(h, sd[p + "localization_fc2.weight"], sd[p + "localization_fc2.bias"])
    return c.view(x.size(0), -1, 2)


def tps_forward(x, sd):
    B, _, H, W = x.shape
    c_prime = localization_network(x, sd)
    inv_delta_C = sd["Transformation.GridGenerator.inv_delta_C"]
    P_hat = sd["Transformation.GridGenerator.P_hat"]
    cz = torch.cat([c_prime, torch.zeros(B, 3, 2, dtype=x.dtype)], dim=1)
    T = torch.bmm(inv_delta_C.repeat(B, 1, 1), cz)
    grid = torch.bmm(P_hat.repeat(B, 1, 1), T).reshape(B, H, W, 2)
    return F.grid_sample(x, grid, padding_mode="border", align_corners=True)


def _block(x, sd, pre):
    out = _cbr(x, sd, pre + ".conv1", pre + ".bn1")
    out = _cbr(out, sd, pre + ".conv2", pre + ".bn2", relu=False)
    res = x
    if (pre + ".downsample.0.weight") in sd:
        res = _cbr(x, sd, pre + ".downsample.0", pre + ".downsample.1", pad=0, relu=False)
    return F.relu(out + res)


def resnet_forward(x, sd):
    n = "FeatureExtraction.ConvNet."
    h = _cbr(x, sd, n + "conv0_1", n + "bn0_1")
    h = _cbr(h, sd, n + "conv0_2", n + "bn0_2")
    h = F.max_pool2d(h, 2, 2)
    h = _block(h, sd, n + "layer1.0")
    h = _cbr(h, sd, n + "conv1", n + "bn1")
    h = F.max_pool2d(h, 2, 2)
    for b in range(2):
        h = _block(h, sd, f"{n}layer2.{b}")
    h = _cbr(h, sd, n + "conv2", n + "bn2")
    h = F.max_pool2d(h, 2, (2, 1), (0, 1))
    for b in range(5):
        h = _block(h, sd, f"{n}layer3.{b}")
    h = _cbr(h, sd, n + "conv3", n + "bn3")
    for b in range(3):
        h = _block(h, sd, f"{n}layer4.{b}")
    h = _cbr(h, sd, n + "conv4_1", n + "bn4_1", stride=(2, 1), pad=(0, 1))
    h = _cbr(h, sd, n + "conv4_2", n + "bn4_2", pad=0)
    return h


def _lstm_cell(g, c):
    i, f, gg, o = g.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c), c


def bidirectional_lstm(x, sd, pre):
    """encoders.py:17-30 through ATen's native (fused, multi-threaded) CPU LSTM, the path ``nn.LSTM`` takes in the
    reference: ``torch.lstm`` with the flat weight list of a bidirectional single-layer LSTM, batch_first."""
    B, T, _ = x.shape
    Hd = sd[pre + ".rnn.weight_hh_l0"].shape[1]
    flat = [sd[f"{pre}.rnn.{n}_l0{sfx}"] for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    zeros = torch.zeros(2, B, Hd, dtype=x.dtype)
    rec = torch.lstm(x, (zeros, zeros), flat, True, 1, 0.0, False, True, True)[0]
    return F.linear(rec, sd[pre + ".linear.weight"], sd[pre + ".linear.bias"])


def attention_decode(enc, sd, steps, text=None):
    p = "decoder.attention_cell."
    B, T, Hd = enc.shape
    w_ih, w_hh = sd[p + "rnn.weight_ih"], sd[p + "rnn.weight_hh"]
    bias = sd[p + "rnn.bias_ih"] + sd[p + "rnn.bias_hh"]
    C = sd["decoder.generator.weight"].shape[0]
    h = torch.zeros(B, Hd, dtype=enc.dtype); c = torch.zeros(B, Hd, dtype=enc.dtype)
    tok = torch.zeros(B, dtype=torch.long)
    logits = torch.zeros(B, steps, C, dtype=enc.dtype)
    for i in range(steps):
        if text is not None:
            tok = text[:, i]
        onehot = F.one_hot(tok, C).to(enc.dtype)                                   # decoders.py:19-24
        proj = F.linear(enc, sd[p + "i2h.weight"])                                  # :81, every step
        hp = F.linear(h, sd[p + "h2h.weight"], sd[p + "h2h.bias"]).unsqueeze(1)     # :82
        e = F.linear(torch.tanh(proj + hp), sd[p + "score.weight"])                 # :85
        alpha = F.softmax(e, dim=1)                                                 # :87
        ctx = torch.bmm(alpha.permute(0, 2, 1), enc).squeeze(1)                     # :88
        g = F.linear(torch.cat([ctx, onehot], 1), w_ih) + F.linear(h, w_hh) + bias  # :90,:99
        h, c = _lstm_cell(g, c)
        step = F.linear(h, sd["decoder.generator.weight"], sd["decoder.generator.bias"])
        logits[:, i] = step
        if text is None:
            tok = step.argmax(dim=1)                                                # :61-62
    return logits


def linear_embedding(overlap, scene, ious, table, mode="overlap"):
    if mode == "overlap":
        return F.embedding(overlap.long(), table)
    return F.embedding(scene.long(), table) * F.softmax(ious, dim=1).unsqueeze(-1)


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], image: torch.Tensor, overlap=None, scene=None, ious=None,
            text=None, is_train: bool = False, max_len: int = 25, semantic_mode: str = "overlap",
            stages: bool = False):
    """Model.forward (model.py:85-107) for ENCODER = DECODER = 'LSTM' on CPU torch tensors."""
    rect = tps_forward(image, sd)
    feat = resnet_forward(rect, sd)
    vis = F.adaptive_avg_pool2d(feat.permute(0, 3, 1, 2), (None, 1)).squeeze(3)
    sem = None
    if overlap is not None:
        sem = linear_embedding(overlap, scene, ious, sd["get_semantic_vectors.embed.weight"], semantic_mode)
    enc = bidirectional_lstm(bidirectional_lstm(vis, sd, "encoder.0"), sd, "encoder.1")
    logits = attention_decode(enc.contiguous(), sd, max_len + 1, text if is_train else None)
    if stages:
        return {"rectified": rect, "features": feat, "visual": vis, "semantics": sem, "enc1": enc,
                "logits": logits, "ids": logits.argmax(2)}
    return logits
