"""ResNet feature extractor behind the reference's module interface (modules/feature_extraction.py).

``ResNet_FeatureExtractor`` / ``ResNet`` / ``BasicBlock`` keep the reference's constructor
arguments, attribute names and therefore state-dict keys (feature_extraction.py:4-107).  Every
Conv2d(bias=False) -> BatchNorm2d(eval) [-> += residual] -> ReLU group is ONE kernel launch:

* conv0_1 (C_in = image channels)  ``mstr_conv3x3_first`` - SIMT, reads the NCHW fp32 image,
                                    writes NHWC activations;
* every other convolution           ``mstr_conv_gemm_bf16`` - TMA-im2col fed tcgen05 implicit GEMM
                                    (config.PRECISION == 'bf16') or ``mstr_conv_gemm_f32``
                                    (exact fp32 parity mode), BN folded to scale/shift, the
                                    residual add and ReLU fused in the epilogue;
* max-pools                         ``mstr_maxpool_nhwc``.

Activations stay NHWC between layers; ``forward`` returns the reference's NCHW view.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from .._pack import (act_dtype, pack_block_tail, pack_conv, pack_conv_flat, pack_conv_pixel_pairs, pack_first_conv,
                     require_eval, require_inference)


def _conv_bn(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, precision: str, relu: bool, residual=None, **kw):
    B, H, W, C = x.shape
    if (precision == 'bf16' and config.CONV_PIXEL_PAIRS and C == 32 and W % 2 == 0 and residual is None
            and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1)):
        # 32-channel input (conv0_2): run it over horizontal pixel pairs (see _pack.pack_conv_pixel_pairs)
        w, scale, shift = pack_conv_pixel_pairs(conv, bn, precision)
        out = ops.conv_gemm(x.view(B, H, W // 2, 2 * C), w, scale, shift, None, stride=(1, 1), pad=(1, 1), relu=relu, **kw)
        return out.view(B, H, W, conv.out_channels)
    w, scale, shift = pack_conv(conv, bn, precision)
    return ops.conv_gemm(x, w, scale, shift, residual, stride=tuple(conv.stride), pad=tuple(conv.padding),
                         relu=relu, **kw)


def _to_nhwc(x: torch.Tensor, precision: str) -> torch.Tensor:
    """NCHW fp32 (module boundary) -> NHWC activation dtype (layout plumbing only)."""
    return ops.cast(x.permute(0, 2, 3, 1).contiguous().float(), act_dtype(precision))


def _to_nchw(x: torch.Tensor) -> torch.Tensor:
    return ops.cast(x, torch.float32).permute(0, 3, 1, 2)


class ResNet_FeatureExtractor(nn.Module):
    """FeatureExtractor of FAN; interface of feature_extraction.py:4-12."""

    def __init__(self, input_channel, output_channel=512):
        super(ResNet_FeatureExtractor, self).__init__()
        self.ConvNet = ResNet(input_channel, output_channel, BasicBlock, [1, 2, 5, 3])

    def forward(self, input):
        return self.ConvNet(input)

    def forward_nhwc(self, input):
        return self.ConvNet.forward_nhwc(input)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = self._conv3x3(inplanes, planes)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = self._conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def _conv3x3(self, in_planes, out_planes, stride=1):
        return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)

    def forward_nhwc(self, x, precision: str):
        """feature_extraction.py:32-47 on NHWC activations: three (two) fused launches."""
        out = _conv_bn(x, self.conv1, self.bn1, precision, relu=True)
        residual = x
        if (self.downsample is not None and precision == 'bf16' and config.FUSE_DOWNSAMPLE
                and x.shape[3] % 64 == 0 and out.shape[3] % 64 == 0 and tuple(self.downsample[0].stride) == (1, 1)):
            # conv2+bn2, the 1x1 downsample conv+bn, the add and the ReLU as ONE tensor-core GEMM
            w_cat, shift = pack_block_tail(self.conv2, self.bn2, self.downsample[0], self.downsample[1])
            return ops.conv_gemm_dual(out, x, w_cat, shift, kernel=tuple(self.conv2.kernel_size),
                                      pad=tuple(self.conv2.padding), relu=True)
        if self.downsample is not None:
            residual = _conv_bn(x, self.downsample[0], self.downsample[1], precision, relu=False)
        return _conv_bn(out, self.conv2, self.bn2, precision, relu=True, residual=residual)

    def forward(self, x):
        require_inference(x)
        require_eval(self)
        precision = config.PRECISION
        return _to_nchw(self.forward_nhwc(_to_nhwc(x, precision), precision))


class ResNet(nn.Module):

    def __init__(self, input_channel, output_channel, block, layers):
        super(ResNet, self).__init__()
        self.output_channel_block = [int(output_channel / 4), int(output_channel / 2), output_channel, output_channel]
        ocb = self.output_channel_block

        self.inplanes = int(output_channel / 8)
        self.conv0_1 = nn.Conv2d(input_channel, int(output_channel / 16), kernel_size=3, stride=1, padding=1, bias=False)
        self.bn0_1 = nn.BatchNorm2d(int(output_channel / 16))
        self.conv0_2 = nn.Conv2d(int(output_channel / 16), self.inplanes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn0_2 = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)

        self.maxpool1 = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)
        self.layer1 = self._make_layer(block, ocb[0], layers[0])
        self.conv1 = nn.Conv2d(ocb[0], ocb[0], kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(ocb[0])

        self.maxpool2 = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)
        self.layer2 = self._make_layer(block, ocb[1], layers[1], stride=1)
        self.conv2 = nn.Conv2d(ocb[1], ocb[1], kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(ocb[1])

        self.maxpool3 = nn.MaxPool2d(kernel_size=2, stride=(2, 1), padding=(0, 1))
        self.layer3 = self._make_layer(block, ocb[2], layers[2], stride=1)
        self.conv3 = nn.Conv2d(ocb[2], ocb[2], kernel_size=3, stride=1, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(ocb[2])

        self.layer4 = self._make_layer(block, ocb[3], layers[3], stride=1)
        self.conv4_1 = nn.Conv2d(ocb[3], ocb[3], kernel_size=2, stride=(2, 1), padding=(0, 1), bias=False)
        self.bn4_1 = nn.BatchNorm2d(ocb[3])
        self.conv4_2 = nn.Conv2d(ocb[3], ocb[3], kernel_size=2, stride=1, padding=0, bias=False)
        self.bn4_2 = nn.BatchNorm2d(ocb[3])

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    @staticmethod
    def _pool(x, pool: nn.MaxPool2d):
        def two(v):
            return (v, v) if isinstance(v, int) else tuple(v)
        return ops.maxpool_nhwc(x, two(pool.kernel_size), two(pool.stride), two(pool.padding))

    def forward_nhwc(self, x):
        """x [B, C, H, W] f32 (NCHW image) -> NHWC [B, H', W', C_out] in the activation dtype
        (feature_extraction.py:109-150)."""
        require_inference(x)
        require_eval(self)
        precision = config.PRECISION
        x = x.contiguous()
        c2, mp = self.conv0_2, self.maxpool1
        flat = (precision == 'bf16' and config.CONV0_FLAT and x.shape[1] <= 4 and x.shape[2] % 2 == 0 and x.shape[3] % 4 == 0
                and (c2.in_channels, c2.out_channels) == (32, 64) and tuple(c2.kernel_size) == (3, 3)
                and tuple(c2.stride) == (1, 1) and tuple(c2.padding) == (1, 1) and x.shape[3] + 2 <= 1000
                and mp.kernel_size in (2, (2, 2)) and mp.stride in (2, (2, 2)) and mp.padding in (0, (0, 0)))
        if flat:
            # conv0_1 -> padded-flat activation -> conv0_2 with one slab load per tile -> maxpool1 skipping the ring
            w, scale, shift = pack_first_conv(self.conv0_1, self.bn0_1)
            h = ops.conv3x3_first(x, w, scale, shift, torch.bfloat16, pool=False, ring=True)
            w, scale, shift = pack_conv_flat(c2, self.bn0_2)
            W = x.shape[3]
            if config.CONV0_POOL_FUSED and (W <= 128 or W % 128 == 0):
                h = ops.conv3x3_flat_pool(h, w, scale, shift, relu=True)      # maxpool1 in the epilogue
            else:
                h = ops.conv3x3_flat(h, w, scale, shift, relu=True)
                h = ops.maxpool2x2_ring(h)
        else:
            if x.shape[1] <= 4:
                w, scale, shift = pack_first_conv(self.conv0_1, self.bn0_1)
                h = ops.conv3x3_first(x, w, scale, shift, act_dtype(precision), pool=False)
            else:
                h = _conv_bn(_to_nhwc(x, precision), self.conv0_1, self.bn0_1, precision, relu=True)
            h = _conv_bn(h, self.conv0_2, self.bn0_2, precision, relu=True)
            h = self._pool(h, self.maxpool1)
        for blk in self.layer1:
            h = blk.forward_nhwc(h, precision)
        h = _conv_bn(h, self.conv1, self.bn1, precision, relu=True)
        h = self._pool(h, self.maxpool2)
        for blk in self.layer2:
            h = blk.forward_nhwc(h, precision)
        h = _conv_bn(h, self.conv2, self.bn2, precision, relu=True)
        h = self._pool(h, self.maxpool3)
        for blk in self.layer3:
            h = blk.forward_nhwc(h, precision)
        h = _conv_bn(h, self.conv3, self.bn3, precision, relu=True)
        for blk in self.layer4:
            h = blk.forward_nhwc(h, precision)
        h = _conv_bn(h, self.conv4_1, self.bn4_1, precision, relu=True)
        h = _conv_bn(h, self.conv4_2, self.bn4_2, precision, relu=True)
        return h

    def forward(self, x):
        """Reference layout at the module boundary: [B, C_out, H', W'] f32 (a view of the NHWC result)."""
        return _to_nchw(self.forward_nhwc(x))
