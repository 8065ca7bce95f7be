"""Torch-tensor front ends of the C-ABI kernels.

Each function validates its tensors (CUDA, contiguous, dtype), allocates the output with torch
(memory stays torch-owned; the C side only borrows pointers) and enqueues the kernel on the
current CUDA stream.  No function here has a PyTorch/CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _cabi, config
from ._cabi import MSTR_BF16, MSTR_F32, check, lib, ptr


class KernelTimer:
    """Optional per-launch CUDA-event timing (bench.py's roofline / stage breakdown).  While
    ``ops.TIMER`` is set, every front end below brackets its launch with two events recorded on
    the current stream; ``summary()`` (after a synchronize) aggregates time and work per kernel."""

    def __init__(self):
        self.records = []   # (name, work, start event, end event, unit, desc)

    def summary(self):
        out = {}
        for name, work, s, e, _unit, _desc in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["work"] += work
        return out

    def table(self, steps: int = 1):
        """Per call site (in launch order of one step): mean duration and achieved rate."""
        per_step = len(self.records) // max(steps, 1)
        rows = []
        for i in range(per_step):
            recs = self.records[i::per_step] if per_step else []
            name, work, _, _, unit, desc = recs[0]
            ms = sum(s.elapsed_time(e) for _, _, s, e, _, _ in recs) / len(recs)
            rate = work / (ms * 1e-3) if ms > 0 else 0.0
            rows.append({"kernel": name, "what": desc, "ms": round(ms, 4), "work": work, "unit": unit,
                         "achieved": round(rate / (1e12 if unit == "flop" else 1e9), 1),
                         "achieved_unit": "TFLOP/s" if unit == "flop" else "GB/s"})
        return rows


TIMER: Optional[KernelTimer] = None


class _timed:
    def __init__(self, name: str, work: float = 0.0, unit: str = "flop", desc: str = ""):
        self.name, self.work, self.unit, self.desc = name, work, unit, desc

    def __enter__(self):
        if TIMER is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if TIMER is not None and exc[0] is None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            TIMER.records.append((self.name, self.work, self.s, e, self.unit, self.desc))
        return False


class SplitWeight:
    """fp32 weight with its TF32 head / tail split (w = hi + lo, hi = w with the low 13 mantissa bits
    cleared) for the 3xTF32 tensor-core parity mode.  Passing one as ``w`` to conv_gemm / linear selects
    mstr_conv_gemm_tf32 where the layer qualifies (C_in % 32 == 0, C_out % 128 == 0); other layers use
    ``full`` with the fp32 SIMT kernel."""

    def __init__(self, w: torch.Tensor, hi: Optional[torch.Tensor] = None, lo: Optional[torch.Tensor] = None):
        self.full = w.contiguous()
        if hi is None:
            hi = (self.full.view(torch.int32) & -8192).view(torch.float32)
            lo = self.full - hi
        self.hi, self.lo = hi.contiguous(), lo.contiguous()

    shape = property(lambda self: self.full.shape)
    dtype = property(lambda self: self.full.dtype)

    def reshape(self, *shape):
        return SplitWeight(self.full.reshape(*shape), self.hi.reshape(*shape), self.lo.reshape(*shape))


class _on:
    """Device guard of one launch: every tensor handed to the kernel must live on ONE CUDA device; that device is made
    current for the duration of the call (the C side launches, encodes TMA descriptors and sets function attributes on
    the current device) and the stream is that device's current stream.  A model on cuda:1 while the process's current
    device is cuda:0 - the reference's default is DEVICE_IDS = [1] (config.py:4) - therefore runs on cuda:1."""

    def __init__(self, *tensors):
        dev = None
        for t in tensors:
            if t is None or isinstance(t, SplitWeight):
                continue
            if not t.is_cuda:
                raise _cabi.MstrError("this path has no CPU implementation: got a %s tensor" % t.device)
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise _cabi.MstrError("tensors of one launch live on different devices: %s and %s" % (dev, t.device))
        self.index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self) -> int:
        cur = torch.cuda.current_device()
        if cur != self.index:
            self.prev = cur
            torch.cuda.set_device(self.index)
        return torch.cuda.current_stream(self.index).cuda_stream

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


_SM_COUNT = {}


def _sm_count(device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return MSTR_F32
    if t.dtype == torch.bfloat16:
        return MSTR_BF16
    raise TypeError("unsupported dtype %s" % t.dtype)


def _req(t: Optional[torch.Tensor], name: str, dtype=None) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise _cabi.MstrError("%s must be a CUDA tensor: this path has no CPU implementation" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))


def tps_rectify(img: Optional[torch.Tensor], c_prime: torch.Tensor, basis: torch.Tensor,
                inv_delta_C: Optional[torch.Tensor], mode: int, want_grid: bool = False):
    """img [B,C,H,W] f32, c_prime [B,F,2] f32 -> rectified [B,C,H,W] f32 (and the grid [B,H*W,2]).

    ``img=None`` with ``want_grid=True`` produces only the grid (GridGenerator.build_P_prime)."""
    _req(img, "img", torch.float32); _req(c_prime, "c_prime", torch.float32)
    _req(basis, "basis", torch.float32); _req(inv_delta_C, "inv_delta_C", torch.float32)
    B, F = c_prime.shape[0], c_prime.shape[1]
    n = basis.shape[1]
    assert c_prime.shape == (B, F, 2)
    assert basis.shape[0] == ((F + 3) if mode == 0 else F)
    if img is None:
        if not want_grid:
            raise ValueError("tps_rectify: nothing to compute (no image and no grid requested)")
        C, H, W, out = 1, 1, n, None
    else:
        _, C, H, W = img.shape
        assert img.shape[0] == B and H * W == n, (img.shape, basis.shape)
        out = torch.empty_like(img)
    grid = torch.empty((B, n, 2), device=c_prime.device, dtype=torch.float32) if want_grid else None
    with _timed("tps_rectify", float(B) * C * n * 8, "byte", "B%d %dx%dx%d image in + out" % (B, C, H, W)), _on(img, c_prime, basis, inv_delta_C) as st:
        check(lib().mstr_tps_rectify(ptr(img), ptr(c_prime), ptr(basis), ptr(inv_delta_C), ptr(out), ptr(grid),
                                     B, C, H, W, F, mode, st), "mstr_tps_rectify")
    if img is None:
        return grid
    return (out, grid) if want_grid else out


def conv3x3_first(x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                  out_dtype: torch.dtype, pool: bool, ring: bool = False) -> torch.Tensor:
    """x [B,Cin,H,W] f32 NCHW -> NHWC [B,H',W',Cout] (conv3x3 pad1 + folded BN + ReLU [+ maxpool 2x2]).
    ``ring=True`` (bf16, not pooled): the padded-flat activation [B,H+2,W+2,Cout] with a zero one-pixel ring that
    conv3x3_flat consumes."""
    _req(x, "x", torch.float32); _req(w, "w", torch.float32)
    _req(scale, "scale", torch.float32); _req(shift, "shift", torch.float32)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    assert w.shape == (Cout, Cin, 3, 3) and not (pool and ring)
    P, Q = (H // 2, W // 2) if pool else ((H + 2, W + 2) if ring else (H, W))
    out = torch.empty((B, P, Q, Cout), device=x.device, dtype=out_dtype)
    with _timed("conv3x3_first", float(out.numel() * out.element_size() + x.numel() * 4), "byte",
                "B%d %dx%d %d->%d%s: NHWC out + image in" % (B, H, W, Cin, Cout, " +pool" if pool else (" +ring" if ring else ""))), \
            _on(x, w, scale, shift) as st:
        check(lib().mstr_conv3x3_first(ptr(x), ptr(w), ptr(scale), ptr(shift), ptr(out), _dt(out), B, Cin, H, W,
                                       Cout, 2 if ring else int(pool), st), "mstr_conv3x3_first")
    return out


def conv3x3_flat(x_ring: torch.Tensor, w: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                 relu: bool = True) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 convolution + folded BN + ReLU over a padded-flat bf16 activation.
    x_ring [B,H+2,W+2,Cin] with a zero ring, w [Cout,3,3,Cin] bf16 -> [B,H+2,W+2,Cout] bf16, ring zero.
    scale / shift: HOST f32 tensors [Cout] (they are passed to the kernel by value; _pack.pack_conv_flat caches host
    copies so that no device->host copy happens per call) or None."""
    _req(x_ring, "x_ring", torch.bfloat16); _req(w, "w", torch.bfloat16)
    B, Hp, Wp, Cin = x_ring.shape
    Cout = w.shape[0]
    assert w.shape == (Cout, 3, 3, Cin) and Hp > 2 and Wp > 2
    for name, t in (("scale", scale), ("shift", shift)):
        if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != Cout):
            raise TypeError("conv3x3_flat: %s must be a contiguous host f32 tensor of %d values" % (name, Cout))
    H, W = Hp - 2, Wp - 2
    out = torch.empty((B, Hp, Wp, Cout), device=x_ring.device, dtype=torch.bfloat16)
    with _timed("conv3x3_flat", 2.0 * B * H * W * Cout * 9 * Cin, "flop",
                "M=%d (B%d %dx%d) %d->%d 3x3, padded-flat" % (B * H * W, B, H, W, Cin, Cout)), _on(x_ring, w) as st:
        check(lib().mstr_conv3x3_flat_bf16(ptr(x_ring), ptr(w), ptr(scale), ptr(shift), ptr(out), B, H, W, Cin, Cout,
                                           int(relu), st), "mstr_conv3x3_flat_bf16")
    return out


def conv3x3_flat_pool(x_ring: torch.Tensor, w: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                      relu: bool = True) -> torch.Tensor:
    """conv3x3_flat followed by the 2x2 / stride 2 max-pool, in one kernel: x_ring [B,H+2,W+2,Cin] (zero ring) ->
    POOLED plain NHWC [B,H/2,W/2,Cout] bf16.  Bit-identical to maxpool2x2_ring(conv3x3_flat(...))."""
    _req(x_ring, "x_ring", torch.bfloat16); _req(w, "w", torch.bfloat16)
    B, Hp, Wp, Cin = x_ring.shape
    Cout = w.shape[0]
    assert w.shape == (Cout, 3, 3, Cin) and Hp > 2 and Wp > 2
    for name, t in (("scale", scale), ("shift", shift)):
        if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != Cout):
            raise TypeError("conv3x3_flat_pool: %s must be a contiguous host f32 tensor of %d values" % (name, Cout))
    H, W = Hp - 2, Wp - 2
    out = torch.empty((B, H // 2, W // 2, Cout), device=x_ring.device, dtype=torch.bfloat16)
    with _timed("conv3x3_flat", 2.0 * B * H * W * Cout * 9 * Cin, "flop",
                "M=%d (B%d %dx%d) %d->%d 3x3, padded-flat, + 2x2 max-pool" % (B * H * W, B, H, W, Cin, Cout)), _on(x_ring, w) as st:
        check(lib().mstr_conv3x3_flat_pool_bf16(ptr(x_ring), ptr(w), ptr(scale), ptr(shift), ptr(out), B, H, W, Cin, Cout,
                                                int(relu), st), "mstr_conv3x3_flat_pool_bf16")
    return out


def flat_desc_mode(mode: int) -> None:
    """Test hook of conv3x3_flat (mstr_debug_flat_desc_mode): 0 = base-offset field left zero (default, correct on
    B200), 1 = set to the phase of the start address (wrong: kept as the experiment that pins the hardware rule)."""
    check(lib().mstr_debug_flat_desc_mode(int(mode)), "mstr_debug_flat_desc_mode")


def maxpool2x2_ring(x_ring: torch.Tensor) -> torch.Tensor:
    """2x2 / stride 2 max-pool of a padded-flat bf16 activation [B,H+2,W+2,C] -> [B,H/2,W/2,C] (ring skipped)."""
    _req(x_ring, "x_ring", torch.bfloat16)
    B, Hp, Wp, C = x_ring.shape
    H, W = Hp - 2, Wp - 2
    out = torch.empty((B, H // 2, W // 2, C), device=x_ring.device, dtype=torch.bfloat16)
    with _timed("maxpool_nhwc", float(B * H * W * C + out.numel()) * 2, "byte",
                "B%d %dx%dx%d (ring) -> %dx%d in + out" % (B, H, W, C, H // 2, W // 2)), _on(x_ring) as st:
        check(lib().mstr_maxpool2x2_ring_bf16(ptr(x_ring), ptr(out), B, H, W, C, st), "mstr_maxpool2x2_ring_bf16")
    return out


def conv_gemm(x: torch.Tensor, w: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
              residual: Optional[torch.Tensor] = None, *, stride: Tuple[int, int] = (1, 1),
              pad: Tuple[int, int] = (0, 0), relu: bool = False, want_f32: bool = False,
              want_bf16: bool = True):
    """Implicit-GEMM convolution + folded BN + residual + ReLU.

    x NHWC [B,H,W,Cin]; w [Cout,R,S,Cin] of the same dtype.  f32 tensors run the exact fp32-FMA
    kernel and return an f32 tensor; bf16 tensors run the tcgen05 kernel and return
    (out_bf16 or None, out_f32 or None) according to want_bf16 / want_f32.
    """
    split = w if isinstance(w, SplitWeight) else None
    if split is not None:
        w = split.full
    _req(x, "x"); _req(w, "w", x.dtype); _req(scale, "scale", torch.float32); _req(shift, "shift", torch.float32)
    _req(residual, "residual", x.dtype)
    B, H, W, Cin = x.shape
    Cout, R, S, Cin2 = w.shape
    assert Cin == Cin2, (x.shape, w.shape)
    P = (H + 2 * pad[0] - R) // stride[0] + 1
    Q = (W + 2 * pad[1] - S) // stride[1] + 1
    if residual is not None:
        assert residual.shape == (B, P, Q, Cout)
    flops = 2.0 * B * P * Q * Cout * R * S * Cin
    desc = "M=%d (B%d %dx%d) %d->%d %dx%d%s" % (B * P * Q, B, P, Q, Cin, Cout, R, S, " +res" if residual is not None else "")
    if x.dtype == torch.float32 and split is not None and Cin % 32 == 0 and Cout % 128 == 0:
        out = torch.empty((B, P, Q, Cout), device=x.device, dtype=torch.float32)
        with _timed("conv_gemm_tf32x3", flops, "flop", desc), _on(x, split.hi, split.lo, scale, shift, residual) as st:
            check(lib().mstr_conv_gemm_tf32(ptr(x), ptr(split.hi), ptr(split.lo), ptr(scale), ptr(shift), ptr(residual),
                                            ptr(out), B, H, W, Cin, Cout, R, S, stride[0], stride[1], pad[0], pad[1],
                                            int(relu), st), "mstr_conv_gemm_tf32")
        return out
    if x.dtype == torch.float32:
        out = torch.empty((B, P, Q, Cout), device=x.device, dtype=torch.float32)
        with _timed("conv_gemm_f32", flops, "flop", desc), _on(x, w, scale, shift, residual) as st:
            check(lib().mstr_conv_gemm_f32(ptr(x), ptr(w), ptr(scale), ptr(shift), ptr(residual), ptr(out), B, H, W,
                                           Cin, Cout, R, S, stride[0], stride[1], pad[0], pad[1], int(relu),
                                           st), "mstr_conv_gemm_f32")
        return out
    if x.dtype != torch.bfloat16:
        raise TypeError("conv_gemm: unsupported dtype %s" % x.dtype)
    out_b = torch.empty((B, P, Q, Cout), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    out_f = torch.empty((B, P, Q, Cout), device=x.device, dtype=torch.float32) if want_f32 else None
    # which kernel mstr_conv_gemm_bf16 dispatches to (conv_tc.cu): the CTA-pair kernel for C_out % 256 == 0
    one_wave = Cout % 128 == 0 and -(-(B * P * Q) // 128) * (Cout // 128) <= _sm_count(x.device)
    variant = "conv_gemm_bf16_2cta" if (Cin % 64 == 0 and Cout % 256 == 0 and not one_wave) else "conv_gemm_bf16_1cta"
    with _timed(variant, flops, "flop", desc), _on(x, w, scale, shift, residual) as st:
        check(lib().mstr_conv_gemm_bf16(ptr(x), ptr(w), ptr(scale), ptr(shift), ptr(residual), ptr(out_b), ptr(out_f),
                                        B, H, W, Cin, Cout, R, S, stride[0], stride[1], pad[0], pad[1], int(relu),
                                        st), "mstr_conv_gemm_bf16")
    if want_f32 and want_bf16:
        return out_b, out_f
    return out_f if want_f32 else out_b


def gemm_route(route: int) -> None:
    """Test hook (mstr_debug_gemm_route): 0 = default dispatch, 1 = single-CTA GEMM kernel always, 2 = CTA pairs always."""
    check(lib().mstr_debug_gemm_route(int(route)), "mstr_debug_gemm_route")


def conv_gemm_dual(x: torch.Tensor, x2: torch.Tensor, w_cat: torch.Tensor, shift: torch.Tensor, *, kernel: Tuple[int, int],
                   pad: Tuple[int, int] = (1, 1), relu: bool = True) -> torch.Tensor:
    """bf16 convolution (stride 1) with a fused 1x1 branch: x NHWC [B,H,W,Cin], x2 NHWC [B,P,Q,Cin2] on the output
    grid, w_cat [Cout, R*S*Cin + Cin2] bf16 with the BatchNorm scales folded in, shift [Cout] f32."""
    _req(x, "x", torch.bfloat16); _req(x2, "x2", torch.bfloat16); _req(w_cat, "w_cat", torch.bfloat16)
    _req(shift, "shift", torch.float32)
    B, H, W, Cin = x.shape
    R, S = kernel
    P, Q = H + 2 * pad[0] - R + 1, W + 2 * pad[1] - S + 1
    Cin2 = x2.shape[3]
    Cout = w_cat.shape[0]
    assert x2.shape == (B, P, Q, Cin2) and w_cat.shape == (Cout, R * S * Cin + Cin2)
    out = torch.empty((B, P, Q, Cout), device=x.device, dtype=torch.bfloat16)
    flops = 2.0 * B * P * Q * Cout * (R * S * Cin + Cin2)
    variant = "conv_gemm_bf16_2cta" if Cout % 256 == 0 else "conv_gemm_bf16_1cta"
    with _timed(variant, flops, "flop", "M=%d (B%d %dx%d) %d->%d %dx%d + fused 1x1 %d->%d" % (
            B * P * Q, B, P, Q, Cin, Cout, R, S, Cin2, Cout)), _on(x, x2, w_cat, shift) as st:
        check(lib().mstr_conv_gemm_bf16_dual(ptr(x), ptr(x2), ptr(w_cat), 0, ptr(shift), 0, ptr(out), 0, B, H, W, Cin,
                                             Cin2, Cout, R, S, 1, 1, pad[0], pad[1], int(relu), st),
              "mstr_conv_gemm_bf16_dual")
    return out


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           want_f32: bool = False, want_bf16: bool = True, residual: Optional[torch.Tensor] = None):
    """nn.Linear on the last dimension through the same GEMM kernels (x [..., K], w [N, K]); ``residual`` [..., N] of
    x's dtype is added in the epilogue (before the ReLU)."""
    lead = x.shape[:-1]
    K = x.shape[-1]
    M = 1
    for d in lead:
        M *= d
    x4 = x.reshape(M, 1, 1, K)
    w4 = w.reshape(w.shape[0], 1, 1, K)
    r4 = None if residual is None else residual.reshape(M, 1, 1, w.shape[0])
    res = conv_gemm(x4, w4, None, bias, r4, relu=relu, want_f32=want_f32, want_bf16=want_bf16)
    shape = tuple(lead) + (w.shape[0],)
    if isinstance(res, tuple):
        return tuple(r.reshape(shape) for r in res)
    return res.reshape(shape)


def maxpool_nhwc(x: torch.Tensor, kernel=(2, 2), stride=(2, 2), pad=(0, 0)) -> torch.Tensor:
    _req(x, "x")
    B, H, W, C = x.shape
    P = (H + 2 * pad[0] - kernel[0]) // stride[0] + 1
    Q = (W + 2 * pad[1] - kernel[1]) // stride[1] + 1
    out = torch.empty((B, P, Q, C), device=x.device, dtype=x.dtype)
    with _timed("maxpool_nhwc", float(x.numel() + out.numel()) * x.element_size(), "byte",
                "B%d %dx%dx%d -> %dx%d in + out" % (B, H, W, C, P, Q)), _on(x) as st:
        check(lib().mstr_maxpool_nhwc(ptr(x), ptr(out), _dt(x), B, H, W, C, kernel[0], kernel[1], stride[0], stride[1],
                                      pad[0], pad[1], st), "mstr_maxpool_nhwc")
    return out


def mean_hw(x: torch.Tensor) -> torch.Tensor:
    """NHWC [B,H,W,C] -> [B,C] f32 (AdaptiveAvgPool2d(1))."""
    _req(x, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, C), device=x.device, dtype=torch.float32)
    with _on(x) as st:
        check(lib().mstr_mean_hw(ptr(x), ptr(out), _dt(x), B, H, W, C, st), "mstr_mean_hw")
    return out


def locnet_head(feat: torch.Tensor, w1t: torch.Tensor, b1: torch.Tensor, w2t: torch.Tensor,
                b2: torch.Tensor) -> torch.Tensor:
    """Localisation head in one launch: NHWC feat [B,H,W,512] -> mean over HW -> fc1 + ReLU -> fc2, [B, O] f32.

    w1t [512, 256] and w2t [256, O] are the TRANSPOSED nn.Linear weights (transformation.py:60-61), f32."""
    _req(feat, "feat")
    for name, t in (("w1t", w1t), ("b1", b1), ("w2t", w2t), ("b2", b2)):
        _req(t, name)
        if t.dtype != torch.float32:
            raise TypeError(f"locnet_head: {name} must be float32")
    B, H, W, C = feat.shape
    H1, O = w2t.shape
    if tuple(w1t.shape) != (C, H1) or b1.numel() != H1 or b2.numel() != O:
        raise ValueError("locnet_head: weight shapes do not chain")
    out = torch.empty((B, O), device=feat.device, dtype=torch.float32)
    work = feat.numel() * feat.element_size() + 4 * (w1t.numel() + w2t.numel() + out.numel())
    with _on(feat, w1t, b1, w2t, b2) as st, _timed("locnet_head", work, "byte", "GAP + fc1 + fc2 of the localisation net"):
        check(lib().mstr_locnet_head(ptr(feat), _dt(feat), ptr(w1t), ptr(b1), ptr(w2t), ptr(b2), ptr(out), B, H * W, C, H1, O,
                                     st), "mstr_locnet_head")
    return out


def mean_h(x: torch.Tensor, want_f32: bool = False):
    """NHWC [B,H,W,C] -> [B,W,C] mean over H, in x.dtype (and optionally an f32 copy)."""
    _req(x, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, W, C), device=x.device, dtype=x.dtype)
    out_f = torch.empty((B, W, C), device=x.device, dtype=torch.float32) if want_f32 else None
    with _on(x) as st:
        check(lib().mstr_mean_h(ptr(x), ptr(out), ptr(out_f), _dt(x), B, H, W, C, st), "mstr_mean_h")
    return (out, out_f) if want_f32 else out


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _req(x, "x")
    if x.dtype == dtype:
        return x
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    with _on(x) as st:
        check(lib().mstr_cast(ptr(x), ptr(out), _dt(x), _dt(out), x.numel(), st), "mstr_cast")
    return out


# ---- transformer family: everything between the GEMMs --------------------------------------------
def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm over the last dimension, x [..., D] f32 or bf16 -> same dtype."""
    _req(x, "x"); _req(gamma, "gamma", torch.float32); _req(beta, "beta", torch.float32)
    D = x.shape[-1]
    out = torch.empty_like(x)
    with _timed("layer_norm", 2.0 * x.numel() * x.element_size(), "byte", "%d rows x %d" % (x.numel() // D, D)), \
            _on(x, gamma, beta) as st:
        check(lib().mstr_layer_norm(ptr(x), ptr(gamma), ptr(beta), ptr(out), x.numel() // D, D, float(eps), _dt(x), st),
              "mstr_layer_norm")
    return out


def add_pos_encoding(x: torch.Tensor, pe: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """x [B,T,D] + pe[:T] -> [B,T,D] of ``dtype`` (PositionalEncoding.forward, eval mode)."""
    _req(x, "x"); _req(pe, "pe", torch.float32)
    B, T, D = x.shape
    assert pe.shape[0] >= T and pe.shape[1] == D
    out = torch.empty((B, T, D), device=x.device, dtype=dtype)
    with _on(x, pe) as st:
        check(lib().mstr_add_pos_encoding(ptr(x), _dt(x), ptr(pe), ptr(out), _dt(out), B, T, D, st), "mstr_add_pos_encoding")
    return out


def embed_tokens(tokens: torch.Tensor, emb: torch.Tensor, pe: torch.Tensor, t0: int, dtype: torch.dtype) -> torch.Tensor:
    """tokens [B,T] int32 -> emb[tokens] + pe[t0:t0+T]  [B,T,D] of ``dtype``."""
    _req(tokens, "tokens", torch.int32); _req(emb, "emb", torch.float32); _req(pe, "pe", torch.float32)
    B, T = tokens.shape
    C, D = emb.shape
    assert pe.shape[0] >= t0 + T and pe.shape[1] == D
    out = torch.empty((B, T, D), device=tokens.device, dtype=dtype)
    with _on(tokens, emb, pe) as st:
        check(lib().mstr_embed_tokens(ptr(tokens), ptr(emb), ptr(pe), ptr(out), _dt(out), B, T, D, C, t0, st), "mstr_embed_tokens")
    return out


def mha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, causal: bool = False) -> torch.Tensor:
    """softmax(q k^T / sqrt(dh) [+ causal mask]) v per head, before out_proj.  q [B,Lq,E], k / v [B,Lk,E]: views with
    unit stride along E and dense batches (stride(0) == L * stride(1)), e.g. the thirds of an in_proj output."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    for t, L in ((q, Lq), (k, Lk), (v, Lk)):
        assert t.is_cuda and t.dtype == q.dtype and t.stride(2) == 1 and t.stride(0) == L * t.stride(1), "mha: unsupported view"
    assert k.stride() == v.stride() and k.shape == v.shape
    out = torch.empty((B, Lq, E), device=q.device, dtype=q.dtype)
    with _timed("mha", 4.0 * B * Lq * Lk * E, "flop", "B%d Lq%d Lk%d %d heads x %d" % (B, Lq, Lk, heads, E // heads)), \
            _on(q, k, v) as st:
        check(lib().mstr_mha(ptr(q), q.stride(1), ptr(k), ptr(v), k.stride(1), ptr(out), B, Lq, Lk, heads, E // heads,
                             int(causal), _dt(q), st), "mstr_mha")
    return out


def mha_step(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, Lk: int, heads: int,
             k_new: Optional[torch.Tensor] = None, v_new: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One query per crop: q [B,E] (a view with row stride q.stride(0)), k_cache / v_cache views [B,>=Lk,E] (any row /
    batch strides, unit stride along E); k_new / v_new [B,E] views are first stored as row Lk-1 of the caches."""
    B, E = q.shape
    for t in (q, k_cache, v_cache, k_new, v_new):
        if t is not None:
            assert t.is_cuda and t.stride(-1) == 1 and t.dtype == q.dtype
    assert k_cache.stride() == v_cache.stride() and k_cache.shape[1] >= Lk
    if k_new is not None:
        assert v_new is not None and k_new.stride(0) == v_new.stride(0)
    out = torch.empty((B, E), device=q.device, dtype=q.dtype)
    with _on(q, k_cache, v_cache, k_new, v_new) as st:
        check(lib().mstr_mha_step(ptr(q), q.stride(0), ptr(k_cache), ptr(v_cache), k_cache.stride(1), k_cache.stride(0),
                                  ptr(k_new), ptr(v_new), 0 if k_new is None else k_new.stride(0), ptr(out), B, Lk, heads,
                                  E // heads, _dt(q), st), "mstr_mha_step")
    return out


def pair_add_relu(a: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """a [B,T,C], s [B,O,C] -> relu(a[:, :, None] + s[:, None]) [B,T,O,C] (same dtype)."""
    _req(a, "a"); _req(s, "s", a.dtype)
    B, T, C = a.shape
    O = s.shape[1]
    assert s.shape == (B, O, C)
    h = torch.empty((B, T, O, C), device=a.device, dtype=a.dtype)
    with _on(a, s) as st:
        check(lib().mstr_pair_add_relu(ptr(a), ptr(s), ptr(h), B, T, O, C, _dt(a), st), "mstr_pair_add_relu")
    return h


def relevance_reduce(h2: torch.Tensor, w3: torch.Tensor, b3: float, sem: torch.Tensor) -> torch.Tensor:
    """h2 [B,T,O,C], w3 [C] f32, sem [B,O,E] f32 -> sum_o softmax_o(h2 . w3 + b3) sem[:, o]  [B,T,E] in h2's dtype."""
    _req(h2, "h2"); _req(w3, "w3", torch.float32); _req(sem, "sem", torch.float32)
    B, T, O, C = h2.shape
    E = sem.shape[2]
    assert sem.shape == (B, O, E) and w3.numel() == C
    out = torch.empty((B, T, E), device=h2.device, dtype=h2.dtype)
    with _on(h2, w3, sem) as st:
        check(lib().mstr_relevance_reduce(ptr(h2), ptr(w3), float(b3), ptr(sem), ptr(out), B, T, O, C, E, _dt(h2), st),
              "mstr_relevance_reduce")
    return out


def gelu_(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU in place."""
    _req(x, "x")
    with _on(x) as st:
        check(lib().mstr_gelu(ptr(x), ptr(x), x.numel(), _dt(x), st), "mstr_gelu")
    return x


def argmax_tokens(logits: torch.Tensor, C: int, logits_out: Optional[torch.Tensor], ids: Optional[torch.Tensor],
                  next_tok: Optional[torch.Tensor]) -> None:
    """logits [B,ld] f32; logits_out / ids / next_tok are (strided) views with one row / element per crop."""
    _req(logits, "logits", torch.float32)
    B, ld = logits.shape
    with _on(logits, logits_out, ids, next_tok) as st:
        check(lib().mstr_argmax_tokens(ptr(logits), ld, ptr(logits_out), 0 if logits_out is None else logits_out.stride(0),
                                       ptr(ids), 0 if ids is None else ids.stride(0), ptr(next_tok),
                                       0 if next_tok is None else next_tok.stride(0), B, C, st), "mstr_argmax_tokens")


def embed_semantics(ids: torch.Tensor, table: torch.Tensor, ious: Optional[torch.Tensor], mode: int) -> torch.Tensor:
    _req(ids, "ids", torch.float32); _req(table, "table", torch.float32); _req(ious, "ious", torch.float32)
    B, n = ids.shape
    rows, D = table.shape
    out = torch.empty((B, n, D), device=ids.device, dtype=torch.float32)
    with _on(ids, table, ious) as st:
        check(lib().mstr_embed_semantics(ptr(ids), ptr(table), ptr(ious), ptr(out), B, n, D, rows, mode, st),
              "mstr_embed_semantics")
    return out


def bilstm_recurrent(gates_x: torch.Tensor, w_hh: torch.Tensor, want_f32: bool, want_bf16: bool):
    """gates_x [B,T,2,4H] f32, w_hh [2,4H,H] f32 -> [B,T,2H] in f32 and/or bf16."""
    _req(gates_x, "gates_x", torch.float32); _req(w_hh, "w_hh", torch.float32)
    B, T = gates_x.shape[:2]
    Hd = w_hh.shape[2]
    assert gates_x.shape == (B, T, 2, 4 * Hd) and w_hh.shape == (2, 4 * Hd, Hd)
    out_f = torch.empty((B, T, 2 * Hd), device=gates_x.device, dtype=torch.float32) if want_f32 else None
    out_b = torch.empty((B, T, 2 * Hd), device=gates_x.device, dtype=torch.bfloat16) if want_bf16 else None
    with _timed("bilstm_recurrent", 2.0 * B * T * 2 * 4 * Hd * Hd, "flop", "B%d T%d both directions (fp32 SIMT)" % (B, T)), _on(gates_x, w_hh) as st:
        check(lib().mstr_bilstm_recurrent(ptr(gates_x), ptr(w_hh), ptr(out_f), ptr(out_b), B, T, Hd, st),
              "mstr_bilstm_recurrent")
    if want_f32 and want_bf16:
        return out_f, out_b
    return out_f if want_f32 else out_b


def pack_w_hh_bf16(w_hh: torch.Tensor) -> torch.Tensor:
    """[2, 4H, H] f32 (gate-major rows i,f,g,o) -> bf16 with rows ordered (chunk of 32 units, gate, unit),
    the layout mstr_bilstm_recurrent_bf16 streams (see include/mstr_b200.h)."""
    two, four_h, hd = w_hh.shape
    assert two == 2 and four_h == 4 * hd and hd % 32 == 0
    w = w_hh.reshape(2, 4, hd // 32, 32, hd).permute(0, 2, 1, 3, 4)      # [dir, chunk, gate, unit, k]
    return w.reshape(2, 4 * hd, hd).to(torch.bfloat16).contiguous()


def bilstm_recurrent_bf16(gates_x: torch.Tensor, w_hh_packed: torch.Tensor) -> torch.Tensor:
    """gates_x [B,T,2,4H] f32 or bf16, w_hh_packed [2,4H,H] bf16 (pack_w_hh_bf16) -> [B,T,2H] bf16."""
    _req(gates_x, "gates_x"); _req(w_hh_packed, "w_hh_packed", torch.bfloat16)
    B, T = gates_x.shape[:2]
    Hd = w_hh_packed.shape[2]
    assert gates_x.shape == (B, T, 2, 4 * Hd) and w_hh_packed.shape == (2, 4 * Hd, Hd)
    out = torch.empty((B, T, 2 * Hd), device=gates_x.device, dtype=torch.bfloat16)
    with _timed("bilstm_recurrent_bf16", 2.0 * B * T * 2 * 4 * Hd * Hd, "flop", "B%d T%d both directions (tcgen05)" % (B, T)), _on(gates_x, w_hh_packed) as st:
        check(lib().mstr_bilstm_recurrent_bf16(ptr(gates_x), _dt(gates_x), ptr(w_hh_packed), ptr(out), B, T, Hd, st),
              "mstr_bilstm_recurrent_bf16")
    return out


def set_bilstm_kernel(which: str = "auto") -> None:
    """Tensor-core BiLSTM recurrence: "auto" picks by occupancy (the cluster kernel while the crop tiles fit in three
    waves of clusters - ~2400 crops -, the one-CTA-per-tile kernel beyond); "single" / "cluster" force one (parity tests)."""
    check(lib().mstr_set_bilstm_kernel({"auto": 0, "single": 1, "cluster": 2}[which]), "mstr_set_bilstm_kernel")


def attn_decode(enc: torch.Tensor, proj: torch.Tensor, packed: dict, steps: int, text: Optional[torch.Tensor],
                early_exit: bool, h0: Optional[torch.Tensor] = None, c0: Optional[torch.Tensor] = None,
                want_alpha: bool = False, want_state: bool = False) -> dict:
    """Greedy / teacher-forced attention decode.

    Returns a dict with logits [B,steps,C] f32, ids [B,steps] i32, lengths [B] i32 and, on request,
    alpha [B,steps,T] f32 and the final (h, c) [B,H] f32.
    """
    _req(enc, "enc", torch.float32); _req(proj, "proj", torch.float32); _req(text, "text", torch.int64)
    _req(h0, "h0", torch.float32); _req(c0, "c0", torch.float32)
    B, T, Hd = enc.shape
    C = packed["w_gen"].shape[0]
    dev = enc.device
    logits = torch.zeros((B, steps, C), device=dev, dtype=torch.float32)
    ids = torch.empty((B, steps), device=dev, dtype=torch.int32)
    lengths = torch.empty((B,), device=dev, dtype=torch.int32)
    alpha = torch.zeros((B, steps, T), device=dev, dtype=torch.float32) if want_alpha else None
    h_out = torch.empty((B, Hd), device=dev, dtype=torch.float32) if want_state else None
    c_out = torch.empty((B, Hd), device=dev, dtype=torch.float32) if want_state else None
    if text is not None:
        assert text.shape[0] == B and text.shape[1] >= steps
        if text.shape[1] != steps:
            text = text[:, :steps].contiguous()
    macs = Hd * Hd + 2 * T * Hd + 4 * Hd * 2 * Hd + C * Hd      # h2h, score+context, LSTMCell, generator
    with _timed("attn_decode", 2.0 * B * steps * macs, "flop", "B%d T%d %d steps (fp32 SIMT, 1 launch)" % (B, T, steps)), _on(enc, proj, text, h0, c0, *[v for v in packed.values() if isinstance(v, torch.Tensor)]) as st:
        check(lib().mstr_attn_decode(ptr(enc), ptr(proj), ptr(packed["w_h2h"]), ptr(packed["b_h2h"]),
                                     ptr(packed["w_score"]), ptr(packed["w_ctx"]), ptr(packed["w_tok"]),
                                     ptr(packed["w_hh"]), ptr(packed["bias"]), ptr(packed["w_gen"]),
                                     ptr(packed["b_gen"]), ptr(text), ptr(h0), ptr(c0), ptr(logits), ptr(ids),
                                     ptr(lengths), ptr(alpha), ptr(h_out), ptr(c_out),
                                     B, T, Hd, C, steps, int(early_exit), st), "mstr_attn_decode")
    return {"logits": logits, "ids": ids, "lengths": lengths, "alpha": alpha, "h": h_out, "c": c_out}


def pack_gate_rows_bf16(w: torch.Tensor) -> torch.Tensor:
    """[4H, K] f32 (gate-major rows i,f,g,o) -> bf16 with rows ordered (chunk of 16 units, gate, unit): every 64 rows
    hold all four gates of 16 hidden units, so the decoder's cell GEMM can use 64- or 128-column tiles (attn_tc.cu)."""
    four_h, k = w.shape
    hd = four_h // 4
    assert four_h == 4 * hd and hd % 32 == 0
    return w.reshape(4, hd // 16, 16, k).permute(1, 0, 2, 3).reshape(four_h, k).to(torch.bfloat16).contiguous()


def resize_to_tensor_u8(pixels: torch.Tensor, offsets: torch.Tensor, sizes_hw: torch.Tensor, max_in_hw,
                        out_hw=(32, 100)) -> torch.Tensor:
    """Ragged batch of grayscale uint8 crops -> [B, 1, out_h, out_w] f32 in [0, 1]: transforms.Resize + ToTensor of
    the reference's loaders (coco_dataset.py:27,40-41), bit-exact with Pillow's bilinear resample.
    pixels: packed uint8 crops (device), offsets [B] int64, sizes_hw [B, 2] int32 (height, width)."""
    _req(pixels, "pixels", torch.uint8); _req(offsets, "offsets", torch.int64); _req(sizes_hw, "sizes_hw", torch.int32)
    B = offsets.shape[0]
    assert sizes_hw.shape == (B, 2)
    out = torch.empty((B, 1, out_hw[0], out_hw[1]), device=pixels.device, dtype=torch.float32)
    status = torch.zeros((1,), device=pixels.device, dtype=torch.int32)
    with _timed("resize_to_tensor_u8", float(pixels.numel() + out.numel() * 4), "byte",
                "B%d ragged u8 crops -> %dx%d f32" % (B, out_hw[0], out_hw[1])), _on(pixels, offsets, sizes_hw) as st:
        check(lib().mstr_resize_to_tensor_u8(ptr(pixels), ptr(offsets), ptr(sizes_hw), ptr(out), ptr(status), B,
                                             int(out_hw[0]), int(out_hw[1]), int(max_in_hw[0]), int(max_in_hw[1]),
                                             st), "mstr_resize_to_tensor_u8")
    if config.CHECK_INPUTS:      # costs a device->host sync; pipeline.pack_crops validates the sizes on the host anyway
        bad = int(status.item())
        if bad:
            raise ValueError("crop %d is empty or larger than max_in_hw=%s: its output was filled with NaN" % (bad - 1, tuple(max_in_hw)))
    return out


def set_decode_persistent(on=None) -> None:
    """Driver of the tensor-core decoder: None / True = ONE persistent cooperative launch for all steps (default),
    False = the chain of three launches per step (the cross-check).  Bit-identical results."""
    check(lib().mstr_set_decode_persistent(-1 if on is None else int(bool(on))), "mstr_set_decode_persistent")


def attn_decode_bf16(enc: torch.Tensor, proj: torch.Tensor, packed: dict, steps: int, text: Optional[torch.Tensor],
                     early_exit: bool, want_alpha: bool = False) -> dict:
    """Tensor-core greedy / teacher-forced attention decode (bf16 mode).  enc, proj [B,T,H] bf16;
    ``packed`` holds w_head, w_cat (bf16) and b_h2h, b_gen, w_score, w_tok, bias (f32)."""
    _req(enc, "enc", torch.bfloat16); _req(proj, "proj", torch.bfloat16); _req(text, "text", torch.int64)
    B, T, Hd = enc.shape
    C = packed["w_tok"].shape[0]
    dev = enc.device
    logits = torch.zeros((B, steps, C), device=dev, dtype=torch.float32)
    ids = torch.empty((B, steps), device=dev, dtype=torch.int32)
    lengths = torch.empty((B,), device=dev, dtype=torch.int32)
    alpha = torch.zeros((B, steps, T), device=dev, dtype=torch.float32) if want_alpha else None
    ws = torch.empty((int(lib().mstr_attn_decode_bf16_workspace_bytes(B, Hd)),), device=dev, dtype=torch.uint8)
    if text is not None:
        assert text.shape[0] == B and text.shape[1] >= steps
        if text.shape[1] != steps:
            text = text[:, :steps].contiguous()
    # the token rows of W_ih with the LSTMCell bias folded in (fp32 add, the same one the kernel would do per element)
    if "w_tokb" not in packed:
        packed["w_tokb"] = (packed["w_tok"] + packed["bias"][None, :]).contiguous()
    macs = Hd * Hd + 2 * T * Hd + 4 * Hd * 2 * Hd + C * Hd
    with _timed("attn_decode_bf16", 2.0 * B * steps * macs, "flop",
                "B%d T%d %d steps (tcgen05, 1 persistent launch)" % (B, T, steps)), _on(enc, proj, text, *[v for v in packed.values() if isinstance(v, torch.Tensor)]) as st:
        check(lib().mstr_attn_decode_bf16(ptr(enc), ptr(proj), ptr(packed["w_head"]), ptr(packed["b_h2h"]),
                                          ptr(packed["b_gen"]), ptr(packed["w_score"]), ptr(packed["w_cat"]),
                                          ptr(packed["w_tokb"]), 0, ptr(text), ptr(ws), ptr(logits),
                                          ptr(ids), ptr(lengths), ptr(alpha), B, T, Hd, C, steps, int(early_exit),
                                          st), "mstr_attn_decode_bf16")
    return {"logits": logits, "ids": ids, "lengths": lengths, "alpha": alpha, "h": None, "c": None}
