"""Weight pre-packing helpers: fold eval-mode BatchNorm, re-lay conv/LSTM weights for the kernels.

Parameters stay ordinary nn.Parameters with the reference's names and shapes (so .pt files load
unchanged, SURVEY.md 5 "checkpoint").  Packed copies are derived lazily and cached on the module;
the cache key is (data_ptr, _version, device, dtype) of every source tensor, so
``load_state_dict`` / ``.to()`` / in-place updates invalidate it.
"""
from __future__ import annotations

from typing import Callable, Iterable

import torch

BN_EPS_DEFAULT = 1e-5


def cached(module: torch.nn.Module, key, sources: Iterable[torch.Tensor], build: Callable[[], object]):
    store = module.__dict__.get('_mstr_pack')
    if store is None:
        store = {}
        object.__setattr__(module, '_mstr_pack', store)
    stamp = tuple((t.data_ptr(), t._version, str(t.device), t.dtype) for t in sources)
    hit = store.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    with torch.no_grad():
        value = build()
    store[key] = (stamp, value)
    return value


def act_dtype(precision: str) -> torch.dtype:
    if precision == 'bf16':
        return torch.bfloat16
    if precision in ('fp32', 'tf32'):
        return torch.float32
    raise ValueError("config.PRECISION must be 'bf16', 'tf32' or 'fp32', got %r" % (precision,))


def wrap_weight(w: torch.Tensor, precision: str):
    """GEMM weight in the activation dtype; for the 'tf32' mode the fp32 weight with its hi/lo split."""
    if precision == 'tf32':
        from .ops import SplitWeight
        return SplitWeight(w.float().contiguous())
    return w.to(act_dtype(precision)).contiguous()


def bn_sources(bn: torch.nn.BatchNorm2d):
    return [bn.weight, bn.bias, bn.running_mean, bn.running_var]


def fold_bn(bn: torch.nn.BatchNorm2d):
    """scale = gamma / sqrt(var + eps), shift = beta - mean * scale  (fp32, eval-mode BN)."""
    require_eval(bn)
    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
    shift = bn.bias.float() - bn.running_mean.float() * scale
    return scale.contiguous(), shift.contiguous()


def pack_conv(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, precision: str):
    """-> (w [Cout,R,S,Cin] in the activation dtype, scale, shift)."""
    def build():
        w = wrap_weight(conv.weight.detach().permute(0, 2, 3, 1).contiguous(), precision)
        scale, shift = fold_bn(bn)
        return w, scale, shift
    return cached(conv, ('conv', precision), [conv.weight] + bn_sources(bn), build)


def pack_conv_flat(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d):
    """Padded-flat kernel (ops.conv3x3_flat): (w [Cout,R,S,Cin] bf16 on the device, scale, shift as HOST f32 tensors -
    the kernel takes the folded BatchNorm by value)."""
    def build():
        scale, shift = fold_bn(bn)
        w = conv.weight.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        return w, scale.cpu().contiguous(), shift.cpu().contiguous()
    return cached(conv, ('flat',), [conv.weight] + bn_sources(bn), build)


def pack_first_conv(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d):
    """Image-reading conv: fp32 [Cout,Cin,3,3] weights (SIMT kernel) + folded BN."""
    def build():
        scale, shift = fold_bn(bn)
        return conv.weight.detach().float().contiguous(), scale, shift
    return cached(conv, ('first',), [conv.weight] + bn_sources(bn), build)


def require_inference(*tensors) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(
            "the B200 kernels implement the inference forward only (SURVEY.md 8b): call under "
            "torch.no_grad() / model.eval(); autograd through them is not supported")


def require_eval(module: torch.nn.Module) -> None:
    """BatchNorm is folded from the running statistics (eval mode).  In train mode the reference normalises with batch
    statistics and updates the running ones (feature_extraction.py:20, transformation.py:50): refuse instead of
    silently returning eval-mode results."""
    if module.training:
        raise RuntimeError(
            "%s is in train mode: the B200 kernels implement eval-mode BatchNorm only - call model.eval() "
            "(get_model() does)" % type(module).__name__)


def pack_conv_pixel_pairs(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, precision: str):
    """3x3 / stride 1 / pad 1 convolution with few channels, re-expressed over horizontal PIXEL PAIRS.

    NHWC [B,H,W,C] with even W is bit-for-bit the same memory as [B,H,W/2,2C]; a conv C_in -> C_out
    becomes a conv 2*C_in -> 2*C_out over super-pixels whose 3x3 weights are the original taps placed
    at (input sub-pixel, output sub-pixel) offsets, the other half structurally zero.  It doubles the
    MMA work but turns a K-block of 32 channels / N = 64 tile (64-byte TMA rows, 64-column MMAs) into
    64 channels / N = 128, which is what the tcgen05 pipeline is efficient at: conv0_2 (32 -> 64 at
    32x100) went from 2.5 ms to the figure in profiles/.  Results are identical: every output is the
    same sum of the same products (zeros add exactly).
    -> (w [2*Cout, 3, 3, 2*Cin], scale [2*Cout], shift [2*Cout])
    """
    def build():
        w = conv.weight.detach().float()                      # [Cout, Cin, 3, 3]
        cout, cin = w.shape[:2]
        wp = torch.zeros((2, cout, 3, 3, 2, cin), device=w.device, dtype=torch.float32)   # [po, co, r, s', pi, ci]
        for po in range(2):
            for sp in range(3):
                for pi in range(2):
                    s_tap = 2 * (sp - 1) + pi - po + 1          # original horizontal tap
                    if 0 <= s_tap <= 2:
                        wp[po, :, :, sp, pi, :] = w[:, :, :, s_tap].permute(0, 2, 1)
        wp = wp.reshape(2 * cout, 3, 3, 2 * cin).to(act_dtype(precision)).contiguous()
        scale, shift = fold_bn(bn)
        return wp, scale.repeat(2).contiguous(), shift.repeat(2).contiguous()
    return cached(conv, ('pairs', precision), [conv.weight] + bn_sources(bn), build)


def pack_block_tail(conv2: torch.nn.Conv2d, bn2: torch.nn.BatchNorm2d, ds_conv: torch.nn.Conv2d,
                    ds_bn: torch.nn.BatchNorm2d):
    """conv2/bn2 and the 1x1 downsample conv/bn of a BasicBlock as ONE GEMM: both BatchNorm scales folded into
    the bf16 weights (the two branches share an accumulator), weights concatenated along K, shifts added.
    -> (w_cat [Cout, R*S*planes + inplanes] bf16, shift [Cout] f32)"""
    def build():
        s2, b2 = fold_bn(bn2)
        sd, bd = fold_bn(ds_bn)
        w2 = conv2.weight.detach().float().permute(0, 2, 3, 1) * s2[:, None, None, None]     # [Cout, R, S, Cin]
        wd = ds_conv.weight.detach().float().permute(0, 2, 3, 1) * sd[:, None, None, None]   # [Cout, 1, 1, Cin2]
        cout = w2.shape[0]
        w_cat = torch.cat([w2.reshape(cout, -1), wd.reshape(cout, -1)], dim=1).to(torch.bfloat16).contiguous()
        return w_cat, (b2 + bd).contiguous()
    return cached(conv2, ('block_tail',), [conv2.weight, ds_conv.weight] + bn_sources(bn2) + bn_sources(ds_bn), build)
