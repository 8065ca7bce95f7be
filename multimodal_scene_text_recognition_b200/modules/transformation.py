"""TPS rectification behind the reference's module interface (modules/transformation.py).

Class names, constructor arguments, ``forward`` signatures and registered parameter / buffer
names are the reference's (transformation.py:7-38 TPS_SpatialTransformerNetwork, :41-82
LocalizationNetwork, :85-163 GridGenerator), so its state dicts load unchanged.  The arithmetic
runs in the sm_100a CUDA extension:

* LocalizationNetwork  - ``mstr_conv3x3_first`` (conv.0+BN+ReLU+MaxPool fused, SIMT),
                         ``mstr_conv_gemm_*`` (conv.4/.8/.12 + folded BN + ReLU),
                         ``mstr_maxpool_nhwc``, ``mstr_mean_hw``, fc1/fc2 as fp32 GEMMs;
* GridGenerator + F.grid_sample - ONE fused kernel, ``mstr_tps_rectify``.

There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import config, ops
from .._pack import cached, fold_bn, pack_conv, pack_first_conv, act_dtype, require_eval, require_inference
from ..synth import rare_fiducial_bias, tps_constants


class TPS_SpatialTransformerNetwork(nn.Module):
    """Rectification network of RARE (TPS based STN); interface of transformation.py:7-38."""

    def __init__(self, F, I_size, I_r_size, I_channel_num=1):
        super(TPS_SpatialTransformerNetwork, self).__init__()
        self.F = F
        self.I_size = I_size
        self.I_r_size = I_r_size
        self.I_channel_num = I_channel_num
        self.LocalizationNetwork = LocalizationNetwork(self.F, self.I_channel_num)
        self.GridGenerator = GridGenerator(self.F, self.I_r_size)

    def forward(self, batch_I):
        """batch_I [B, C, H, W] f32 -> rectified [B, C, H_r, W_r] f32 (transformation.py:28-38)."""
        require_inference(batch_I)
        batch_I = batch_I.contiguous()
        if tuple(batch_I.shape[2:]) != tuple(self.I_r_size):
            # the fused kernel samples an output grid of the input's own size, which is the only
            # configuration the reference model uses (model.py:40-41: I_size == I_r_size)
            raise NotImplementedError("TPS kernel requires I_size == I_r_size (model.py:40-41)")
        batch_C_prime = self.LocalizationNetwork(batch_I)
        return self.GridGenerator.rectify(batch_I, batch_C_prime)


class LocalizationNetwork(nn.Module):
    """Predicts the fiducial points C' [B, F, 2] from the image (transformation.py:41-82)."""

    def __init__(self, F, I_channel_num):
        super(LocalizationNetwork, self).__init__()
        self.F = F
        self.I_channel_num = I_channel_num
        # same Sequential layout as transformation.py:47-58 so parameter names match (conv.0, conv.1, ...)
        self.conv = nn.Sequential(
            nn.Conv2d(self.I_channel_num, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(True),
            nn.MaxPool2d(2, 2),
            nn.Conv2d(64, 128, 3, 1, 1, bias=False), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.MaxPool2d(2, 2),
            nn.Conv2d(128, 256, 3, 1, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(True),
            nn.MaxPool2d(2, 2),
            nn.Conv2d(256, 512, 3, 1, 1, bias=False), nn.BatchNorm2d(512), nn.ReLU(True),
            nn.AdaptiveAvgPool2d(1),
        )
        self.localization_fc1 = nn.Sequential(nn.Linear(512, 256), nn.ReLU(True))
        self.localization_fc2 = nn.Linear(256, self.F * 2)
        # transformation.py:64-72: zero weight, bias = RARE Fig. 6(a) fiducials
        self.localization_fc2.weight.data.fill_(0)
        self.localization_fc2.bias.data = torch.from_numpy(rare_fiducial_bias(F)).float().view(-1)

    def forward(self, batch_I):
        require_inference(batch_I)
        require_eval(self)
        precision = config.PRECISION
        batch_I = batch_I.contiguous()
        B = batch_I.size(0)
        w, scale, shift = pack_first_conv(self.conv[0], self.conv[1])
        h = ops.conv3x3_first(batch_I, w, scale, shift, act_dtype(precision), pool=True)
        for i in (4, 8, 12):
            w, scale, shift = pack_conv(self.conv[i], self.conv[i + 1], precision)
            h = ops.conv_gemm(h, w, scale, shift, pad=(1, 1), relu=True)
            if i != 12:
                h = ops.maxpool_nhwc(h, (2, 2), (2, 2), (0, 0))
        # AdaptiveAvgPool2d(1) + view, fc1 + ReLU, fc2 as one fp32 launch: the fiducial points stay fp32-accurate in every
        # precision mode (transformation.py:76-81)
        fc1, fc2 = self.localization_fc1[0], self.localization_fc2
        w1t, b1, w2t, b2 = cached(self, "head_f32", [fc1.weight, fc1.bias, fc2.weight, fc2.bias], lambda: (
            fc1.weight.detach().float().t().contiguous(), fc1.bias.detach().float().contiguous(),
            fc2.weight.detach().float().t().contiguous(), fc2.bias.detach().float().contiguous()))
        c = ops.locnet_head(h, w1t, b1, w2t, b2)
        return c.view(B, self.F, 2)


class GridGenerator(nn.Module):
    """Grid generator of RARE (transformation.py:85-163): P' = P_hat (inv_delta_C [C'; 0])."""

    def __init__(self, F, I_r_size):
        super(GridGenerator, self).__init__()
        self.eps = 1e-6
        self.I_r_height, self.I_r_width = I_r_size
        self.F = F
        inv_delta_C, P_hat = tps_constants(F, self.I_r_height, self.I_r_width)   # fp64, like :103-152
        self.register_buffer("inv_delta_C", torch.tensor(inv_delta_C).float())   # F+3 x F+3
        self.register_buffer("P_hat", torch.tensor(P_hat).float())               # n x F+3

    # -- packed constants ---------------------------------------------------------------------
    def _basis(self, mode: int) -> torch.Tensor:
        if mode == 0:   # reference order: basis = P_hat^T [F+3, n]
            return cached(self, "basis0", [self.P_hat], lambda: self.P_hat.t().contiguous())

        def folded():   # A = P_hat inv_delta_C[:, :F] in fp64 ([C';0] has three zero rows), transposed
            a = self.P_hat.double() @ self.inv_delta_C.double()[:, :self.F]
            return a.t().contiguous().float()
        return cached(self, "basis1", [self.P_hat, self.inv_delta_C], folded)

    @staticmethod
    def _mode() -> int:
        if config.TPS_GRID == 'reference':
            return 0
        if config.TPS_GRID == 'folded':
            return 1
        raise ValueError("config.TPS_GRID must be 'reference' or 'folded', got %r" % (config.TPS_GRID,))

    def build_P_prime(self, batch_C_prime):
        """batch_C_prime [B, F, 2] -> P' [B, n, 2]  (transformation.py:154-163)."""
        mode = self._mode()
        return ops.tps_rectify(None, batch_C_prime.contiguous().float(), self._basis(mode),
                               self.inv_delta_C, mode, want_grid=True)

    def rectify(self, batch_I, batch_C_prime, want_grid: bool = False):
        """Fused build_P_prime + grid_sample(bilinear, border, align_corners=True)."""
        mode = self._mode()
        return ops.tps_rectify(batch_I, batch_C_prime.contiguous(), self._basis(mode), self.inv_delta_C,
                               mode, want_grid=want_grid)
