"""Run the UNMODIFIED reference (``baseline/_ref``, see install_ref.py) on the synthetic workload:
its own ``model.Model`` (model.py:34-107) loaded with the seeded synthetic state dict, ``Model.forward(...,
is_train=False)`` under ``torch.no_grad()`` / ``.eval()``, fp32 - on the host cores (bench.py's reference
arm / ``cpu_baseline``, kind "reference") or eagerly on the GPU with cuDNN disabled as model.py:6 sets it
(``gpu_eager_baseline``).  Checker / baseline infrastructure: nothing in the product package imports this."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import install_ref  # noqa: E402


def available() -> bool:
    return install_ref.available()


class ReferenceModel:
    """The reference's Model with a {name: numpy} state dict (or its own seeded default init)."""

    def __init__(self, state_dict=None, device: str = "cpu", seed=None):
        self.module, self.config = install_ref.load(device=device)
        if seed is not None:
            torch.manual_seed(seed)
        self.model = self.module.Model().eval()
        if state_dict is not None:
            self.model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state_dict.items()},
                                       strict=True)
        self.to(device)

    def to(self, device: str):
        self.device = device
        self.config.PRIMARY_DEVICE = device      # decoders.py:36-38,54 allocate on config.PRIMARY_DEVICE at call time
        self.model.to(device)
        return self

    def forward(self, image, overlap, scene, ious, text=None, is_train=False):
        """numpy or torch inputs -> logits tensor [B, MAX_TEXT_LENGTH+1, classes] on ``self.device``."""
        t = lambda a: (torch.from_numpy(a) if isinstance(a, np.ndarray) else a).to(self.device)
        image, overlap, scene, ious = t(image), t(overlap), t(scene), t(ious)
        if text is None:
            text = torch.zeros((image.shape[0], self.config.MAX_TEXT_LENGTH + 1), dtype=torch.long)
        with torch.no_grad():
            return self.model(image, t(text), overlap, scene, ious, is_train=is_train)

    def stages(self, image):
        """rectified image, ResNet features (NCHW), encoder outputs - the reference's own sub-modules."""
        m = self.model
        x = (torch.from_numpy(image) if isinstance(image, np.ndarray) else image).to(self.device)
        with torch.no_grad():
            rect = m.Transformation(x)
            feat = m.FeatureExtraction(rect)
            vis = m.AdaptiveAvgPool(feat.permute(0, 3, 1, 2)).squeeze(3)
            enc0 = m.encoder[0](vis)
            enc1 = m.encoder[1](enc0)
        return {"rectified": rect, "features": feat, "visual": vis, "enc0": enc0, "enc1": enc1}

    def forward_fp64(self, image):
        """The same network evaluated in float64 by the reference's own modules.  ``Attention.forward`` hard-codes
        ``torch.FloatTensor`` state (decoders.py:36-38,54), so its greedy loop (decoders.py:52-62) is driven from
        here on the dtype-agnostic ``AttentionCell`` + ``generator``.  CPU only."""
        import copy
        m = copy.deepcopy(self.model).double().cpu()
        x = (torch.from_numpy(image) if isinstance(image, np.ndarray) else image).double()
        steps = self.config.MAX_TEXT_LENGTH + 1
        with torch.no_grad():
            feat = m.FeatureExtraction(m.Transformation(x))
            vis = m.AdaptiveAvgPool(feat.permute(0, 3, 1, 2)).squeeze(3)
            enc = m.encoder(vis).contiguous()
            dec = m.decoder
            B = enc.shape[0]
            h = (torch.zeros(B, dec.hidden_size, dtype=torch.float64), torch.zeros(B, dec.hidden_size, dtype=torch.float64))
            tok = torch.zeros(B, dtype=torch.long)
            logits = torch.zeros(B, steps, dec.num_classes, dtype=torch.float64)
            for i in range(steps):
                onehot = torch.zeros(B, dec.num_classes, dtype=torch.float64).scatter_(1, tok[:, None], 1.0)
                h, _alpha = dec.attention_cell(h, enc, onehot, None)
                step = dec.generator(h[0])
                logits[:, i] = step
                tok = step.argmax(1)
        return {"features": feat, "enc1": enc, "logits": logits}


def time_forward(ref: ReferenceModel, image, overlap, scene, ious, reps: int = 1) -> float:
    """Best wall time (s) of ``reps`` forwards (synchronising the GPU when the model is on one)."""
    best = float("inf")
    for _ in range(reps):
        if ref.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref.forward(image, overlap, scene, ious)
        if ref.device != "cpu":
            torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
