"""Semantic-vector modules behind the reference's interface (modules/semantic_vectors.py).

``Linear_Embedding`` (semantic_vectors.py:11-49) is the runnable one: ``nn.Embedding(2000,
EMBED_DIM)`` as the parameter container (key ``embed.weight``), gather (+ softmax(ious) scaling in
'scene' mode) in ``mstr_embed_semantics``.  ``config.SEMANTIC_VECTOR`` is read at construction
and again at every forward, exactly where the reference reads it (:20,:35,:38).

``Zero`` and ``Random`` keep the reference's two-argument forward (:80,:95) - the reference's
``Model.forward`` passes three arguments (model.py:96), so there as here they are only usable
stand-alone.  ``Bert_Embedding`` cannot be constructed in the reference either (``self.config``
is undefined at :63).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from .._pack import require_inference


class Linear_Embedding(nn.Module):
    """[batch, num objs] object ids -> [batch, num objs, embed dim]."""

    def __init__(self):
        super(Linear_Embedding, self).__init__()
        self.num_obj_classes = 2000
        if config.SEMANTIC_VECTOR == 'overlap' or config.SEMANTIC_VECTOR == 'scene':
            self.embed = nn.Embedding(self.num_obj_classes, config.EMBED_DIM)
        elif config.SEMANTIC_VECTOR == 'combined':
            self.overlap_embed = nn.Embedding(self.num_obj_classes, config.EMBED_DIM)
            self.scene_embed = nn.Embedding(self.num_obj_classes, config.EMBED_DIM)
            self.combine = nn.Linear((config.EMBED_DIM * 2), config.EMBED_DIM)
        else:
            raise Exception("Config: SEMANTIC_VECTOR invalid value: " + config.SEMANTIC_VECTOR)

    def forward(self, overlap, scene, ious):
        if config.SEMANTIC_VECTOR == 'overlap':
            ids, weights, mode = overlap, None, 0
        elif config.SEMANTIC_VECTOR == 'scene':
            ids, weights, mode = scene, ious, 1
        else:
            # semantic_vectors.py:44-47 uses the undefined self.embed and concatenates [B,15,D]
            # with [B,52,D] on dim 2: the reference raises here as well
            raise RuntimeError("SEMANTIC_VECTOR='combined' is not runnable in the reference "
                               "(semantic_vectors.py:44-47)")
        table = self.embed.weight
        require_inference(table)
        ids = ids.contiguous().float()
        if config.CHECK_INPUTS and ids.numel():   # costs a device->host sync; off on the hot path
            lo, hi = int(ids.min()), int(ids.max())
            if lo < 0 or hi >= table.shape[0]:
                raise IndexError("object id out of range for the %d-row embedding table: [%d, %d]"
                                 % (table.shape[0], lo, hi))
        return ops.embed_semantics(ids, table.detach().float().contiguous(),
                                   None if weights is None else weights.contiguous().float(), mode)


class Bert_Embedding(nn.Module):
    def __init__(self):
        super(Bert_Embedding, self).__init__()
        raise NotImplementedError("Bert_Embedding is not constructible in the reference "
                                  "(semantic_vectors.py:63 reads the undefined self.config) and needs "
                                  "a downloaded DistilBERT checkpoint; out of scope")


class Random(nn.Module):
    """Random tensors of shape [batch, num objs, embed dim] (semantic_vectors.py:73-85)."""

    def forward(self, overlap, scene):
        overlap = torch.rand(overlap.shape[0], overlap.shape[1], config.EMBED_DIM, device=overlap.device)
        scene = torch.rand(scene.shape[0], overlap.shape[1], config.EMBED_DIM, device=scene.device)
        return overlap, scene


class Zero(nn.Module):
    """Zero tensors of shape [batch, num objs, embed dim] (semantic_vectors.py:88-100)."""

    def forward(self, overlap, scene):
        overlap = torch.zeros(overlap.shape[0], overlap.shape[1], config.EMBED_DIM, device=overlap.device)
        scene = torch.zeros(scene.shape[0], overlap.shape[1], config.EMBED_DIM, device=scene.device)
        return overlap, scene
