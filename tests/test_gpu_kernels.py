"""Per-kernel parity: every C-ABI entry point (called through ops.py -> ctypes -> libmstr_b200.so)
against the numpy oracle on the same seeded inputs.  GPU only."""
import os

import numpy as np
import pytest
import torch

from multimodal_scene_text_recognition_b200 import ops, synth
from oracle import recognition_oracle as orc
from gpu_util import DEV, assert_close, dev, err, host, nchw, nhwc

pytestmark = pytest.mark.gpu


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


# ------------------------------------------------------------------------------------------------
# TPS
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("geom,B", [(synth.CFG_A, 19), (synth.CFG_S, 3)])
def test_tps_rectify(geom, B):
    r = rng(1)
    inv_delta_C, P_hat = synth.tps_constants(geom.F, geom.H, geom.W)
    sd = {"Transformation.GridGenerator.inv_delta_C": inv_delta_C.astype(np.float32),
          "Transformation.GridGenerator.P_hat": P_hat.astype(np.float32)}
    img = synth.make_crops(7, B, geom)
    # fiducials around the RARE initialisation, perturbed enough that ~1-5 % of the grid leaves [-1,1]
    c_prime = (synth.rare_fiducial_bias(geom.F).reshape(1, geom.F, 2) +
               r.normal(0, 0.08, (B, geom.F, 2))).astype(np.float32)
    grid64 = orc.build_P_prime(c_prime.astype(np.float64), {k: v.astype(np.float64) for k, v in sd.items()})
    outside = float((np.abs(grid64) > 1).mean())
    assert outside > 0.002, "test recipe must exercise the border clamp (got %.4f)" % outside
    ref_img = orc.grid_sample_bilinear_border(img.astype(np.float64), grid64.reshape(B, geom.H, geom.W, 2))

    basis0 = dev(np.ascontiguousarray(sd["Transformation.GridGenerator.P_hat"].T))
    folded = (P_hat @ inv_delta_C[:, :geom.F]).T.astype(np.float32)
    inv_d = dev(sd["Transformation.GridGenerator.inv_delta_C"])
    # mode 1 carries the fp32 rounding of the folded [F, n] constant (|A| <= 0.97, 20 terms): ~1e-5
    for mode, basis, gtol in ((0, basis0, 3e-5), (1, dev(np.ascontiguousarray(folded)), 2e-5)):
        out, grid = ops.tps_rectify(dev(img), dev(c_prime), basis, inv_d, mode, want_grid=True)
        torch.cuda.synchronize()
        assert_close(host(grid), grid64, gtol, "tps grid mode %d" % mode)
        d, _ = err(host(out), ref_img)
        assert d < 4e-4, "tps image mode %d: max abs err %.3e" % (mode, d)
        only = ops.tps_rectify(None, dev(c_prime), basis, inv_d, mode, want_grid=True)
        assert torch.equal(only, grid), "grid-only call differs from the fused call"
    # against the fp32 oracle in the reference's own order the fused kernel is at fp32 noise level
    g32 = orc.build_P_prime(c_prime, sd)
    out0, grid0 = ops.tps_rectify(dev(img), dev(c_prime), basis0, inv_d, 0, want_grid=True)
    assert_close(host(grid0), g32, 3e-5, "tps grid vs fp32 oracle")


# ------------------------------------------------------------------------------------------------
# first-layer conv
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [5, 19])      # the kernel walks 8 images per block: partial and several groups
@pytest.mark.parametrize("cin,cout,pool,dtype", [(1, 32, False, torch.bfloat16), (1, 64, True, torch.bfloat16),
                                                 (3, 32, False, torch.float32), (3, 64, True, torch.float32)])
def test_conv3x3_first(cin, cout, pool, dtype, B):
    r = rng(2)
    H, W = 32, 100
    x = r.uniform(0, 1, (B, cin, H, W)).astype(np.float32)
    w = (r.standard_normal((cout, cin, 3, 3)) * 0.3).astype(np.float32)
    scale = r.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = r.normal(0, 0.2, cout).astype(np.float32)
    ref = orc.relu(orc.conv2d(x.astype(np.float64), w.astype(np.float64), (1, 1), (1, 1)) *
                   scale[None, :, None, None] + shift[None, :, None, None])
    if pool:
        ref = orc.maxpool2d(ref)
    out = ops.conv3x3_first(dev(x), dev(w), dev(scale), dev(shift), dtype, pool)
    torch.cuda.synchronize()
    assert_close(nchw(host(out)), ref, 1e-5 if dtype == torch.float32 else 5e-3, "conv3x3_first")


# ------------------------------------------------------------------------------------------------
# implicit-GEMM conv, fp32 and tcgen05 bf16
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, H, W, Cin, Cout, R, S, stride, pad, residual, relu
    (3, 8, 25, 128, 256, 3, 3, (1, 1), (1, 1), False, True),     # layer2 conv, M=600 (ragged tile)
    (2, 4, 26, 64, 64, 2, 2, (2, 1), (0, 1), False, True),       # conv4_1 geometry (k2 s(2,1) p(0,1))
    (2, 2, 27, 64, 128, 2, 2, (1, 1), (0, 0), False, True),      # conv4_2 geometry (k2 s1 p0)
    (5, 16, 50, 64, 128, 1, 1, (1, 1), (0, 0), False, False),    # 1x1 downsample
    (4, 4, 26, 512, 512, 3, 3, (1, 1), (1, 1), True, True),      # dominant layer + residual, K=4608
    (7, 32, 100, 32, 64, 3, 3, (1, 1), (1, 1), False, True),     # conv0_2: Cin=32 (64-byte swizzle), M=22400
    (1, 1, 1, 256, 1024, 1, 1, (1, 1), (0, 0), False, False),    # a single-row linear
]


def _conv_ref(x, w, scale, shift, res, stride, pad, relu):
    y = orc.conv2d(nchw(x).astype(np.float64), w.transpose(0, 3, 1, 2).astype(np.float64), stride, pad)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if res is not None:
        y = y + nchw(res).astype(np.float64)
    return nhwc(orc.relu(y) if relu else y)


def _conv_inputs(case, seed, bf16):
    B, H, W, Cin, Cout, R, S, stride, pad, residual, relu = case
    r = rng(seed)
    x = r.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (r.standard_normal((Cout, R, S, Cin)) / np.sqrt(R * S * Cin)).astype(np.float32)
    scale = r.uniform(0.5, 1.5, Cout).astype(np.float32)
    shift = r.normal(0, 0.2, Cout).astype(np.float32)
    P = (H + 2 * pad[0] - R) // stride[0] + 1
    Q = (W + 2 * pad[1] - S) // stride[1] + 1
    res = r.standard_normal((B, P, Q, Cout)).astype(np.float32) if residual else None
    if bf16:
        x, w = orc.round_bf16(x), orc.round_bf16(w)
        res = None if res is None else orc.round_bf16(res)
    return x, w, scale, shift, res


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_gemm_f32(case):
    stride, pad, relu = case[7], case[8], case[10]
    x, w, scale, shift, res = _conv_inputs(case, 3, False)
    ref = _conv_ref(x, w, scale, shift, res, stride, pad, relu)
    out = ops.conv_gemm(dev(x), dev(w), dev(scale), dev(shift), None if res is None else dev(res),
                        stride=stride, pad=pad, relu=relu)
    torch.cuda.synchronize()
    assert_close(host(out), ref, 5e-6, "conv_gemm_f32 %s" % (case,))


TF32_CASES = [c for c in CONV_CASES if c[3] % 32 == 0 and c[4] % 128 == 0]


@pytest.mark.parametrize("case", TF32_CASES)
def test_conv_gemm_tf32x3_tensor_cores(case):
    """3xTF32 (hi/lo split operands on tcgen05.mma.kind::tf32) must be as accurate as the fp32 FMA kernel."""
    stride, pad, relu = case[7], case[8], case[10]
    x, w, scale, shift, res = _conv_inputs(case, 3, False)
    ref = _conv_ref(x, w, scale, shift, res, stride, pad, relu)
    out = ops.conv_gemm(dev(x), ops.SplitWeight(dev(w)), dev(scale), dev(shift), None if res is None else dev(res),
                        stride=stride, pad=pad, relu=relu)
    torch.cuda.synchronize()
    assert_close(host(out), ref, 5e-6, "conv_gemm_tf32x3 %s" % (case,))
    # and it really is a different code path from the SIMT kernel: many tiles, ragged M
    if case[4] == 512:
        sw = ops.SplitWeight(dev(w))
        assert torch.equal(sw.hi + sw.lo, sw.full) and not torch.equal(sw.hi, sw.full)


def test_linear_f32_small_n():
    """nn.Linear with N=40 (localization_fc2) and a bias but no scale."""
    r = rng(4)
    x = r.standard_normal((9, 256)).astype(np.float32)
    w = (r.standard_normal((40, 256)) / 16).astype(np.float32)
    b = r.standard_normal(40).astype(np.float32)
    out = ops.linear(dev(x), dev(w), dev(b), relu=True)
    torch.cuda.synchronize()
    assert_close(host(out), np.maximum(x.astype(np.float64) @ w.T + b, 0), 2e-6, "linear f32 N=40")


@pytest.fixture(params=[0, 2], ids=["default-dispatch", "cta-pairs-forced"])
def gemm_route(request):
    """Small test shapes are one-wave problems, which the default dispatch gives to the single-CTA kernel: run the GEMM
    tests a second time with CTA pairs forced wherever they apply, so that both kernels keep their coverage."""
    ops.gemm_route(request.param)
    yield request.param
    ops.gemm_route(0)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_gemm_bf16_tcgen05(case, gemm_route):
    stride, pad, relu = case[7], case[8], case[10]
    x, w, scale, shift, res = _conv_inputs(case, 5, True)
    ref = _conv_ref(x, w, scale, shift, res, stride, pad, relu)      # fp64 on the bf16-rounded operands
    bf = torch.bfloat16
    out_b, out_f = ops.conv_gemm(dev(x, bf), dev(w, bf), dev(scale), dev(shift),
                                 None if res is None else dev(res, bf), stride=stride, pad=pad, relu=relu,
                                 want_f32=True, want_bf16=True)
    torch.cuda.synchronize()
    # fp32 accumulation of exact bf16 products: only summation-order noise remains
    assert_close(host(out_f), ref, 2e-5, "conv_gemm_bf16 f32 output %s" % (case,))
    assert np.array_equal(host(out_b), orc.round_bf16(host(out_f))), "bf16 output is not the rounded f32 output"


def test_conv_gemm_bf16_two_accumulators_per_cta():
    """C_out = 128 layers with >= 8 tiles per SM run with two 128-row accumulators per CTA sharing each
    weight tile (conv_gemm_bf16_kernel<128, 64, false, 2>): M = 193*16*50 = 154400 (ragged: not a multiple of
    256).  The oracle is evaluated on the first and last crops (convolution is per-crop independent)."""
    bf = torch.bfloat16
    for case in ((193, 16, 50, 64, 128, 3, 3, (1, 1), (1, 1), True, True),
                 (193, 16, 50, 128, 128, 3, 3, (1, 1), (1, 1), False, True)):
        x, w, scale, shift, res = _conv_inputs(case, 29, True)
        sel = [0, 1, 2, 95, 96, 190, 191, 192]
        ref = _conv_ref(x[sel], w, scale, shift, None if res is None else res[sel], case[7], case[8], case[10])
        out_b = ops.conv_gemm(dev(x, bf), dev(w, bf), dev(scale), dev(shift), None if res is None else dev(res, bf),
                              stride=case[7], pad=case[8], relu=case[10], want_f32=False, want_bf16=True)
        torch.cuda.synchronize()
        got = host(out_b)[sel].astype(np.float64)
        # one bf16 ulp of the result + the fp32 accumulation noise of sums whose terms are O(max |ref|)
        tol = np.abs(ref) * 2.0 ** -7 + 3e-5 * np.abs(ref).max()
        bad = np.abs(got - ref) > tol
        assert not bad.any(), "%d of %d outputs off by more than a bf16 ulp (max err %.3e) %s" % (
            bad.sum(), bad.size, np.abs(got - ref).max(), case)


@pytest.mark.parametrize("B,H,W,planes,inplanes", [(3, 8, 25, 256, 128), (5, 4, 26, 512, 256), (3, 16, 50, 128, 64),
                                                   (193, 16, 50, 128, 64)])
def test_conv_gemm_bf16_fused_downsample(B, H, W, planes, inplanes, gemm_route):
    """mstr_conv_gemm_bf16_dual: 3x3 conv + fused 1x1 branch + ReLU in one accumulator (CTA-pair kernel for
    C_out % 256 == 0, single-CTA kernel with one or two accumulators per CTA for C_out = 128)."""
    r = rng(31)
    x = orc.round_bf16(r.standard_normal((B, H, W, planes)).astype(np.float32))
    x2 = orc.round_bf16(r.standard_normal((B, H, W, inplanes)).astype(np.float32))
    w2 = orc.round_bf16((r.standard_normal((planes, 3, 3, planes)) / np.sqrt(9 * planes)).astype(np.float32))
    wd = orc.round_bf16((r.standard_normal((planes, 1, 1, inplanes)) / np.sqrt(inplanes)).astype(np.float32))
    shift = r.normal(0, 0.2, planes).astype(np.float32)
    sel = list(range(B)) if B <= 8 else [0, 1, 2, 95, 96, 190, 191, 192]
    ref = (orc.conv2d(nchw(x[sel]).astype(np.float64), w2.transpose(0, 3, 1, 2).astype(np.float64), (1, 1), (1, 1)) +
           orc.conv2d(nchw(x2[sel]).astype(np.float64), wd.transpose(0, 3, 1, 2).astype(np.float64), (1, 1), (0, 0)) +
           shift[None, :, None, None])
    ref = nhwc(orc.relu(ref))
    bf = torch.bfloat16
    w_cat = np.concatenate([w2.reshape(planes, -1), wd.reshape(planes, -1)], axis=1)
    out = ops.conv_gemm_dual(dev(x, bf), dev(x2, bf), dev(w_cat, bf), dev(shift), kernel=(3, 3), pad=(1, 1), relu=True)
    torch.cuda.synchronize()
    got = host(out)[sel].astype(np.float64)
    # one bf16 ulp of the result + the fp32 accumulation noise of sums whose terms are O(max |ref|)
    tol = np.abs(ref) * 2.0 ** -7 + 3e-5 * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), "%d of %d outputs off by more than a bf16 ulp (max err %.3e)" % (
        bad.sum(), bad.size, np.abs(got - ref).max())


def test_conv_gemm_bf16_many_tiles():
    """More M tiles than SMs: exercises the persistent tile loop, the smem ring wrap-around and both
    TMEM accumulator buffers.  M = 64*16*50 = 51200 -> 400 M tiles x 1 N tile."""
    case = (64, 16, 50, 64, 128, 3, 3, (1, 1), (1, 1), True, True)
    x, w, scale, shift, res = _conv_inputs(case, 6, True)
    ref = _conv_ref(x, w, scale, shift, res, (1, 1), (1, 1), True)
    bf = torch.bfloat16
    out_f = ops.conv_gemm(dev(x, bf), dev(w, bf), dev(scale), dev(shift), dev(res, bf), stride=(1, 1),
                          pad=(1, 1), relu=True, want_f32=True, want_bf16=False)
    torch.cuda.synchronize()
    assert_close(host(out_f), ref, 2e-5, "conv_gemm_bf16 400 tiles")


def test_small_im2col_tensor_map(gemm_route):
    """im2col tensor maps over tensors smaller than 128 KiB need the descriptor fix of tc_common.cuh
    (fix_small_im2col_map; drivers <= CUDA 13.1).  3x3 convolutions over 20 KiB .. 100 KiB activations on every
    tensor-core kernel that builds such a map - single-CTA bf16, CTA-pair bf16, 3xTF32 - against the oracle: if a
    driver changes the descriptor layout this fails loudly instead of the model drifting."""
    bf = torch.bfloat16
    for case in ((1, 4, 10, 64, 128, 3, 3, (1, 1), (1, 1), False, True),      # 5 KiB, single-CTA kernel
                 (2, 4, 13, 128, 256, 3, 3, (1, 1), (1, 1), True, True),      # 26 KiB, CTA-pair kernel
                 (3, 4, 26, 64, 256, 3, 3, (1, 1), (1, 1), False, False)):    # 39 KiB, CTA-pair kernel
        assert case[0] * case[1] * case[2] * case[3] * 2 < 131072
        x, w, scale, shift, res = _conv_inputs(case, 41, True)
        ref = _conv_ref(x, w, scale, shift, res, case[7], case[8], case[10])
        out_f = ops.conv_gemm(dev(x, bf), dev(w, bf), dev(scale), dev(shift), None if res is None else dev(res, bf),
                              stride=case[7], pad=case[8], relu=case[10], want_f32=True, want_bf16=False)
        torch.cuda.synchronize()
        assert_close(host(out_f), ref, 2e-5, "small im2col tensor (bf16) %s" % (case,))
    case = (2, 4, 13, 64, 128, 3, 3, (1, 1), (1, 1), False, True)             # 26 KiB of fp32, 3xTF32 kernel
    x, w, scale, shift, res = _conv_inputs(case, 43, False)
    ref = _conv_ref(x, w, scale, shift, res, case[7], case[8], case[10])
    out = ops.conv_gemm(dev(x), ops.SplitWeight(dev(w)), dev(scale), dev(shift), None, stride=case[7], pad=case[8],
                        relu=case[10])
    torch.cuda.synchronize()
    assert_close(host(out), ref, 5e-6, "small im2col tensor (3xTF32)")


# ------------------------------------------------------------------------------------------------
# padded-flat path: conv0_1 (ring output) -> conv3x3_flat -> maxpool over the ring
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,cin,H,W", [(5, 1, 32, 100), (9, 3, 8, 12), (2, 3, 64, 256)])
def test_padded_flat_conv_path(B, cin, H, W):
    """mstr_conv3x3_first(pool=2) writes [B,H+2,W+2,32] with a zero ring; mstr_conv3x3_flat_bf16 (one slab load per
    256-position tile, nine taps as row-shifted shared-memory descriptors) must equal the
    convolution of the interior evaluated exactly, rounded to bf16, and write a zero ring; mstr_maxpool2x2_ring_bf16
    pools the interior.  Tiles straddle image rows and images; the first / last tiles read rows outside the tensor."""
    r = rng(51)
    x = r.uniform(0, 1, (B, cin, H, W)).astype(np.float32)
    w1 = (r.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    s1, b1 = r.uniform(0.5, 1.5, 32).astype(np.float32), r.normal(0, 0.2, 32).astype(np.float32)
    w2 = orc.round_bf16((r.standard_normal((64, 3, 3, 32)) / np.sqrt(9 * 32)).astype(np.float32))
    s2, b2 = r.uniform(0.5, 1.5, 64).astype(np.float32), r.normal(0, 0.2, 64).astype(np.float32)
    bf = torch.bfloat16
    plain = ops.conv3x3_first(dev(x), dev(w1), dev(s1), dev(b1), bf, pool=False)
    ring = ops.conv3x3_first(dev(x), dev(w1), dev(s1), dev(b1), bf, pool=False, ring=True)
    torch.cuda.synchronize()
    assert tuple(ring.shape) == (B, H + 2, W + 2, 32)
    assert torch.equal(ring[:, 1:-1, 1:-1], plain), "interior of the ring output differs from the plain output"
    edge = ring.clone(); edge[:, 1:-1, 1:-1] = 0
    assert not edge.any(), "the ring must be zero"
    # flat convolution vs float64 on the same bf16 activations
    act = host(plain).astype(np.float64)                                     # [B,H,W,32]
    ref = orc.conv2d(nchw(act), w2.transpose(0, 3, 1, 2).astype(np.float64), (1, 1), (1, 1))
    ref = nhwc(orc.relu(ref * s2.astype(np.float64)[None, :, None, None] + b2.astype(np.float64)[None, :, None, None]))
    out = ops.conv3x3_flat(ring, dev(w2, bf), torch.from_numpy(s2), torch.from_numpy(b2), relu=True)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, H + 2, W + 2, 64)
    got = host(out[:, 1:-1, 1:-1]).astype(np.float64)
    tol = np.abs(ref) * 2.0 ** -7 + 2e-5 * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), "%d of %d outputs off by more than a bf16 ulp (max err %.3e of max %.3e)" % (
        bad.sum(), bad.size, np.abs(got - ref).max(), np.abs(ref).max())
    exact = (got == orc.round_bf16(ref.astype(np.float32)).astype(np.float64)).mean()
    assert exact > 0.99, "only %.4f of the outputs are the bf16 rounding of the exact result" % exact
    edge = out.clone(); edge[:, 1:-1, 1:-1] = 0
    assert not edge.any(), "the output ring must be zero"
    # equals the im2col kernel on the same operands up to accumulation order
    same = ops.conv_gemm(plain, dev(w2, bf), dev(s2), dev(b2), None, stride=(1, 1), pad=(1, 1), relu=True)
    torch.cuda.synchronize()
    assert (host(same) == host(out[:, 1:-1, 1:-1])).mean() > 0.995
    # pool over the ring == pool of the interior
    pooled = ops.maxpool2x2_ring(out)
    ref_pool = ops.maxpool_nhwc(out[:, 1:-1, 1:-1].contiguous(), (2, 2), (2, 2), (0, 0))
    torch.cuda.synchronize()
    assert torch.equal(pooled, ref_pool)
    # conv0_2 with maxpool1 in its epilogue (tiles of two image rows): bit-identical to the two launches
    if W <= 128 or W % 128 == 0:
        s2h, b2h = torch.from_numpy(s2), torch.from_numpy(b2)
        two = ops.maxpool2x2_ring(ops.conv3x3_flat(ring, dev(w2, bf), s2h, b2h, relu=True))
        fused = ops.conv3x3_flat_pool(ring, dev(w2, bf), s2h, b2h, relu=True)
        torch.cuda.synchronize()
        assert tuple(fused.shape) == (B, H // 2, W // 2, 64)
        assert torch.equal(fused, two), "fused conv0_2 + maxpool differs from conv3x3_flat + maxpool2x2_ring"


def test_padded_flat_conv_descriptor_base_offset():
    """Pins the hardware rule conv_flat.cu relies on: the tensor core applies the operand swizzle to the bits of the
    final shared-memory address, so an A descriptor that starts inside a swizzle atom (a row-shifted view of the slab)
    is correct with the "matrix base offset" field left at ZERO; setting the field to the phase of the start address
    (debug mode 1) gives a wrong result (measured: 0.58 of the output range).  If mode 1 ever becomes the correct one,
    the descriptor semantics have changed and the kernel needs a second look."""
    r = rng(52)
    B, H, W = 3, 8, 12
    bf = torch.bfloat16
    act = orc.round_bf16(r.uniform(0, 1, (B, H, W, 32)).astype(np.float32))
    ring = np.zeros((B, H + 2, W + 2, 32), np.float32); ring[:, 1:-1, 1:-1] = act
    w2 = orc.round_bf16((r.standard_normal((64, 3, 3, 32)) / np.sqrt(288)).astype(np.float32))
    outs = {}
    try:
        for mode in (0, 1):
            ops.flat_desc_mode(mode)
            outs[mode] = host(ops.conv3x3_flat(dev(ring, bf), dev(w2, bf), None, None, relu=False)[:, 1:-1, 1:-1])
            torch.cuda.synchronize()
    finally:
        ops.flat_desc_mode(0)
    ref = nhwc(orc.conv2d(nchw(act.astype(np.float64)), w2.transpose(0, 3, 1, 2).astype(np.float64), (1, 1), (1, 1)))
    e0 = np.abs(outs[0] - ref).max() / np.abs(ref).max()
    e1 = np.abs(outs[1] - ref).max() / np.abs(ref).max()
    print("\nflat conv: max error without the base offset %.2e, with it %.2e" % (e0, e1))
    assert e0 < 1e-2, "plain row-shifted descriptors must be right (%.3e)" % e0
    assert e1 > 1e-1, "with the base-offset field the result was expected to be wrong (%.3e)" % e1


# ------------------------------------------------------------------------------------------------
# pooling / reductions / cast / embedding
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,W,k,s,p", [(32, 100, (2, 2), (2, 2), (0, 0)), (8, 25, (2, 2), (2, 1), (0, 1)),
                                        (16, 50, (2, 2), (2, 2), (0, 0))])
def test_maxpool_nhwc(dtype, H, W, k, s, p):
    x = orc.round_bf16(rng(7).standard_normal((3, H, W, 64)).astype(np.float32))
    ref = nhwc(orc.maxpool2d(nchw(x), k, s, p))
    out = ops.maxpool_nhwc(dev(x, dtype), k, s, p)
    torch.cuda.synchronize()
    assert np.array_equal(host(out), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B", [1, 13, 2371])
def test_locnet_head_is_fp32_accurate_and_batch_invariant(dtype, B):
    """GAP + fc1 + ReLU + fc2 of the localisation net in one launch (transformation.py:57,60-61,76-81) against float64;
    a crop's fiducials must not depend on which batch (CTA tile, 8- or 16-crop kernel) it is in."""
    r = rng(81)
    x = orc.round_bf16(np.maximum(r.standard_normal((B, 4, 12, 512)), 0).astype(np.float32))
    w1 = (r.standard_normal((256, 512)) * 0.06).astype(np.float32)
    b1 = (r.standard_normal(256) * 0.1).astype(np.float32)
    w2 = (r.standard_normal((40, 256)) * 0.08).astype(np.float32)
    b2 = r.standard_normal(40).astype(np.float32)
    ref = np.maximum(x.astype(np.float64).mean(axis=(1, 2)) @ w1.T.astype(np.float64) + b1, 0) @ w2.T.astype(np.float64) + b2
    args = [dev(np.ascontiguousarray(w1.T)), dev(b1), dev(np.ascontiguousarray(w2.T)), dev(b2)]
    xd = dev(x, dtype)
    out = ops.locnet_head(xd, *args)
    torch.cuda.synchronize()
    assert_close(host(out), ref, 3e-6, "locnet_head")
    # same crops as a 5-crop batch starting at crop 3 (other tile position, 8-crop kernel): identical bits
    if B > 8:
        sub = ops.locnet_head(xd[3:8].contiguous(), *args)
        torch.cuda.synchronize()
        assert np.array_equal(host(sub), host(out)[3:8])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mean_reductions_and_cast(dtype):
    x = orc.round_bf16(rng(8).standard_normal((5, 3, 65, 512)).astype(np.float32))
    xd = dev(x, dtype)
    assert_close(host(ops.mean_hw(xd)), x.astype(np.float64).mean(axis=(1, 2)), 2e-6, "mean_hw")
    out, out_f = ops.mean_h(xd, want_f32=True)
    assert_close(host(out_f), x.astype(np.float64).mean(axis=1), 2e-6, "mean_h f32")
    if dtype == torch.bfloat16:
        assert np.array_equal(host(out), orc.round_bf16(host(out_f)))
    else:
        assert np.array_equal(host(out), host(out_f))
    y = rng(9).standard_normal(100003).astype(np.float32)
    assert np.array_equal(host(ops.cast(dev(y), torch.bfloat16)), orc.round_bf16(y))
    assert np.array_equal(host(ops.cast(dev(y, torch.bfloat16), torch.float32)), orc.round_bf16(y))


def test_embed_semantics():
    B = 33
    table = rng(10).standard_normal((2000, 256)).astype(np.float32)
    ov, sc, io = synth.make_semantics(11, B)
    out = ops.embed_semantics(dev(ov), dev(table), None, 0)
    assert np.array_equal(host(out), orc.linear_embedding(ov, sc, io, table, "overlap"))
    out = ops.embed_semantics(dev(sc), dev(table), dev(io), 1)
    assert_close(host(out), orc.linear_embedding(ov, sc, io.astype(np.float64), table, "scene", np.float64), 1e-6,
                 "embed scene (-1000 ious)")
    io2 = rng(12).uniform(0, 1, io.shape).astype(np.float32)
    out = ops.embed_semantics(dev(sc), dev(table), dev(io2), 1)
    assert_close(host(out), orc.linear_embedding(ov, sc, io2.astype(np.float64), table, "scene", np.float64), 1e-6,
                 "embed scene (random ious)")


# ------------------------------------------------------------------------------------------------
# BiLSTM recurrence
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T", [(1, 26), (60, 26), (57, 7), (3, 65)])
def test_bilstm_recurrent(B, T):
    r = rng(13)
    Hd = 256
    gx = r.standard_normal((B, T, 2, 4 * Hd)).astype(np.float32)
    w_hh = (r.uniform(-1, 1, (2, 4 * Hd, Hd)) / 16).astype(np.float32)
    ref = np.concatenate([orc._lstm_direction(gx[:, :, d].astype(np.float64), w_hh[d].astype(np.float64), bool(d))
                          for d in range(2)], axis=2)
    out_f, out_b = ops.bilstm_recurrent(dev(gx), dev(w_hh), want_f32=True, want_bf16=True)
    torch.cuda.synchronize()
    assert_close(host(out_f), ref, 2e-5, "bilstm_recurrent")
    assert np.array_equal(host(out_b), orc.round_bf16(host(out_f)))


@pytest.mark.parametrize("kernel", ["single", "cluster"])
@pytest.mark.parametrize("B,T", [(1, 26), (130, 26), (128, 5), (3, 65), (2, 1), (5, 2), (700, 3)])
def test_bilstm_recurrent_bf16_tensor_cores(B, T, kernel):
    """tcgen05 recurrence (h_t in bf16, tanh.approx gates) vs the oracle with bf16-rounded W_hh / h_t, for both
    kernels: one CTA per crop tile streaming W_hh (bilstm_tc.cu) and the 8-CTA cluster with stationary W_hh
    slices and the DSMEM exchange of h (bilstm_cluster.cu); the dispatcher picks by occupancy."""
    r = rng(17)
    Hd = 256
    gx = r.standard_normal((B, T, 2, 4 * Hd)).astype(np.float32)
    w_hh = (r.uniform(-1, 1, (2, 4 * Hd, Hd)) / 16).astype(np.float32)
    ref = np.concatenate([orc._lstm_direction(gx[:, :, d].astype(np.float64), w_hh[d].astype(np.float64), bool(d),
                                              quant="bf16") for d in range(2)], axis=2)
    try:
        ops.set_bilstm_kernel(kernel)
        out = ops.bilstm_recurrent_bf16(dev(gx), ops.pack_w_hh_bf16(dev(w_hh)))
        torch.cuda.synchronize()
    finally:
        ops.set_bilstm_kernel("auto")
    # tanh.approx (2^-11) and bf16 storage of h (2^-9): a rounding flip of h_t moves later steps by ~1 ulp
    assert_close(host(out), ref, 2e-2, "bilstm_recurrent_bf16")
    d = np.abs(host(out).astype(np.float64) - ref)
    assert d.mean() < 2e-3, "mean abs error %.3e" % d.mean()


@pytest.mark.parametrize("kernel", ["single", "cluster"])
def test_bilstm_gates_x_in_bf16_equals_f32_of_the_same_values(kernel):
    """config.GATES_X_BF16: the recurrence reads the input projection as bf16 (half the traffic) and widens it on the
    way into the accumulators - so fed with bf16 values it must give, bit for bit, what the f32 path gives on the same
    values (ragged tiles, both directions, T not a multiple of anything)."""
    g = torch.Generator(device=DEV).manual_seed(12)
    for B, T in [(130, 26), (5, 3), (700, 7), (1, 65)]:
        gx_b = (torch.randn(B, T, 2, 1024, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
        w = ops.pack_w_hh_bf16(torch.randn(2, 1024, 256, device=DEV, generator=g) / 16)
        try:
            ops.set_bilstm_kernel(kernel)
            a = ops.bilstm_recurrent_bf16(gx_b.float(), w)
            b = ops.bilstm_recurrent_bf16(gx_b, w)
            torch.cuda.synchronize()
        finally:
            ops.set_bilstm_kernel("auto")
        assert torch.equal(a, b), "B=%d T=%d %s" % (B, T, kernel)


def test_bilstm_cluster_kernel_bit_identical_to_single_cta_kernel():
    """Both tensor-core recurrences feed the same bf16 operands to the MMAs in the same k order, so the cluster
    kernel (DSMEM exchange of h) must reproduce the single-CTA kernel bit for bit: a race detector for the
    cross-CTA protocol (tools/soak_bilstm.py runs it longer)."""
    g = torch.Generator(device=DEV).manual_seed(11)
    for i, (B, T) in enumerate([(768, 26), (129, 65), (1, 3), (640, 26), (257, 10), (300, 26)]):
        gx = torch.randn(B, T, 2, 1024, device=DEV, generator=g) * (1.0 + (i % 3))
        w = ops.pack_w_hh_bf16(torch.randn(2, 1024, 256, device=DEV, generator=g) / 16)
        try:
            ops.set_bilstm_kernel("single")
            a = ops.bilstm_recurrent_bf16(gx, w)
            ops.set_bilstm_kernel("cluster")
            b = ops.bilstm_recurrent_bf16(gx, w)
            torch.cuda.synchronize()
        finally:
            ops.set_bilstm_kernel("auto")
        assert torch.equal(a, b), "B=%d T=%d" % (B, T)


# ------------------------------------------------------------------------------------------------
# attention decoder
# ------------------------------------------------------------------------------------------------
def _decoder_pack(sd, Hd=256):
    p = "decoder.attention_cell."
    w_ih = sd[p + "rnn.weight_ih"]
    pk = {"w_h2h": sd[p + "h2h.weight"], "b_h2h": sd[p + "h2h.bias"], "w_score": sd[p + "score.weight"][0],
          "w_ctx": w_ih[:, :Hd], "w_tok": w_ih[:, Hd:].T, "w_hh": sd[p + "rnn.weight_hh"],
          "bias": sd[p + "rnn.bias_ih"] + sd[p + "rnn.bias_hh"],
          "w_gen": sd["decoder.generator.weight"], "b_gen": sd["decoder.generator.bias"]}
    return {k: dev(np.ascontiguousarray(v, dtype=np.float32)) for k, v in pk.items()}


@pytest.fixture(scope="module")
def decoder_case(sd_cfg_a):
    B, T, Hd = 37, 26, 256
    enc = (rng(14).standard_normal((B, T, Hd)) * 0.5).astype(np.float32)
    proj = (enc.astype(np.float64) @ sd_cfg_a["decoder.attention_cell.i2h.weight"].astype(np.float64).T).astype(np.float32)
    return enc, proj, _decoder_pack(sd_cfg_a)


def test_attn_decode_greedy(sd_cfg_a, decoder_case):
    enc, proj, pk = decoder_case
    B, steps = enc.shape[0], 26
    ref, ref_alpha = orc.attention_decode(enc, sd_cfg_a, steps, 97, dt=np.float64, return_alpha=True)
    res = ops.attn_decode(dev(enc), dev(proj), pk, steps, None, early_exit=False, want_alpha=True)
    torch.cuda.synchronize()
    ids = host(res["ids"])
    ref_ids = ref.argmax(2)
    assert np.array_equal(ids, ref_ids), "greedy ids differ on %d of %d crops" % ((ids != ref_ids).any(1).sum(), B)
    assert_close(host(res["logits"]), ref, 1e-4, "greedy logits")
    assert_close(host(res["alpha"]), ref_alpha, 1e-4, "attention weights")
    first = np.where((ref_ids == 1).any(1), (ref_ids == 1).argmax(1), steps)
    assert np.array_equal(host(res["lengths"]), first)


def test_attn_decode_early_exit(sd_cfg_a, decoder_case):
    enc, proj, pk = decoder_case
    steps = 26
    full = ops.attn_decode(dev(enc), dev(proj), pk, steps, None, early_exit=False)
    early = ops.attn_decode(dev(enc), dev(proj), pk, steps, None, early_exit=True)
    torch.cuda.synchronize()
    L = host(full["lengths"])
    assert np.array_equal(host(early["lengths"]), L)
    fl, el, fi, ei = host(full["logits"]), host(early["logits"]), host(full["ids"]), host(early["ids"])
    for b in range(enc.shape[0]):
        k = min(int(L[b]) + 1, steps)           # steps up to and including the one that emitted [s]
        assert np.array_equal(el[b, :k], fl[b, :k]) and np.array_equal(ei[b, :k], fi[b, :k])
        assert not el[b, k:].any() and not ei[b, k:].any()


def test_attn_decode_teacher_forced_and_single_step(sd_cfg_a, decoder_case):
    enc, proj, pk = decoder_case
    B, steps = enc.shape[0], 26
    text = rng(15).integers(0, 97, (B, steps)).astype(np.int64)
    text[:, 0] = 0
    ref = orc.attention_decode(enc, sd_cfg_a, steps, 97, text=text, dt=np.float64)
    res = ops.attn_decode(dev(enc), dev(proj), pk, steps, dev(text), early_exit=False)
    assert_close(host(res["logits"]), ref, 1e-4, "teacher-forced logits")
    # two single steps chained through (h, c) equal the first two steps of the fused loop
    s1 = ops.attn_decode(dev(enc), dev(proj), pk, 1, dev(text[:, :1]), early_exit=False, want_state=True)
    s2 = ops.attn_decode(dev(enc), dev(proj), pk, 1, dev(np.ascontiguousarray(text[:, 1:2])), early_exit=False,
                         h0=s1["h"], c0=s1["c"], want_state=True)
    torch.cuda.synchronize()
    assert_close(host(s1["logits"])[:, 0], ref[:, 0], 1e-4, "single step 0")
    assert_close(host(s2["logits"])[:, 0], ref[:, 1], 1e-4, "single step 1")


def test_attn_decode_long_sequence(sd_cfg_a):
    """T=65 encoder columns, 51 steps (stress configuration)."""
    B, T, Hd, steps = 5, 65, 256, 51
    enc = (rng(16).standard_normal((B, T, Hd)) * 0.5).astype(np.float32)
    proj = (enc.astype(np.float64) @ sd_cfg_a["decoder.attention_cell.i2h.weight"].astype(np.float64).T).astype(np.float32)
    ref = orc.attention_decode(enc, sd_cfg_a, steps, 97, dt=np.float64)
    res = ops.attn_decode(dev(enc), dev(proj), _decoder_pack(sd_cfg_a), steps, None, early_exit=False)
    torch.cuda.synchronize()
    assert np.array_equal(host(res["ids"]), ref.argmax(2))
    assert_close(host(res["logits"]), ref, 2e-4, "long-sequence logits")


def _decoder_pack_bf16(sd, Hd=256):
    p = "decoder.attention_cell."
    t = lambda k: torch.from_numpy(np.ascontiguousarray(sd[k])).to(DEV)
    w_ih = t(p + "rnn.weight_ih")
    C = sd["decoder.generator.weight"].shape[0]
    head = torch.zeros((384, Hd), device=DEV)
    head[:Hd] = t(p + "h2h.weight")
    head[Hd:Hd + C] = t("decoder.generator.weight")
    return {"w_head": head.to(torch.bfloat16).contiguous(),
            "w_cat": ops.pack_gate_rows_bf16(torch.cat([w_ih[:, :Hd], t(p + "rnn.weight_hh")], 1)),
            "w_tok": w_ih[:, Hd:].t().contiguous(), "bias": (t(p + "rnn.bias_ih") + t(p + "rnn.bias_hh")).contiguous(),
            "b_h2h": t(p + "h2h.bias"), "b_gen": t("decoder.generator.bias"),
            "w_score": t(p + "score.weight").view(-1).contiguous()}


@pytest.mark.parametrize("B,T,steps", [(300, 26, 26), (1100, 26, 26), (9, 65, 51)])
def test_attn_decode_bf16_persistent_kernel_matches_launch_chain(sd_cfg_a, B, T, steps):
    """The persistent decoder (one cooperative launch, grid barriers between the phases of a step) must give exactly
    the result of the chain of three launches per step: greedy, early exit, teacher forcing; ragged crop tiles; more
    tiles than SMs (B = 1100: 72 cell tiles + several crops per warp); work queued behind the call."""
    Hd = 256
    enc = orc.round_bf16((rng(28).standard_normal((B, T, Hd)) * 0.5).astype(np.float32))
    w_i2h = orc.round_bf16(sd_cfg_a["decoder.attention_cell.i2h.weight"])
    proj = orc.round_bf16((enc.astype(np.float64) @ w_i2h.astype(np.float64).T).astype(np.float32))
    pk = _decoder_pack_bf16(sd_cfg_a)
    bf = torch.bfloat16
    text = rng(29).integers(0, 97, (B, steps)).astype(np.int64)
    text[:, 0] = 0
    e, p_ = dev(enc, bf), dev(proj, bf)

    def run():
        out = []
        for txt, ee in ((None, False), (None, True), (dev(text), False)):
            r = ops.attn_decode_bf16(e, p_, pk, steps, txt, early_exit=ee, want_alpha=True)
            # consumer on the caller's stream right behind the call
            out.append((r["logits"].clone(), r["ids"].clone(), r["lengths"].clone(), r["alpha"].clone()))
        torch.cuda.synchronize()
        return [[host(t) for t in r] for r in out]

    try:
        ops.set_decode_persistent(False)
        chain = run()
        ops.set_decode_persistent(True)
        persistent = run()
        again = run()
    finally:
        ops.set_decode_persistent(None)
    for other in (persistent, again):
        for a, b in zip(chain, other):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("B,T,steps", [(37, 26, 26), (130, 26, 26), (5, 65, 51)])
def test_attn_decode_bf16_tensor_cores(sd_cfg_a, B, T, steps):
    """Tensor-core decode vs the oracle with bf16 rounding at the same operands.  Teacher-forced logits
    are compared tightly (no feedback); greedy ids must agree wherever the oracle's top-2 margin is
    larger than the bf16-level logit error."""
    Hd = 256
    enc = orc.round_bf16((rng(18).standard_normal((B, T, Hd)) * 0.5).astype(np.float32))
    w_i2h = orc.round_bf16(sd_cfg_a["decoder.attention_cell.i2h.weight"])
    proj = orc.round_bf16((enc.astype(np.float64) @ w_i2h.astype(np.float64).T).astype(np.float32))
    pk = _decoder_pack_bf16(sd_cfg_a)
    bf = torch.bfloat16
    text = rng(19).integers(0, 97, (B, steps)).astype(np.int64)
    text[:, 0] = 0
    ref_tf, ref_alpha = orc.attention_decode(enc, sd_cfg_a, steps, 97, text=text, dt=np.float64, quant="bf16",
                                             return_alpha=True)
    res = ops.attn_decode_bf16(dev(enc, bf), dev(proj, bf), pk, steps, dev(text), early_exit=False, want_alpha=True)
    torch.cuda.synchronize()
    # bf16 rounding of h / ctx flips (1 ulp = 0.4 %) decorrelate the two sides: ~1e-2 of the logit range
    assert_close(host(res["logits"]), ref_tf, 3e-2, "teacher-forced logits (tensor cores)")
    assert_close(host(res["alpha"]), ref_alpha, 3e-2, "attention weights (tensor cores)")
    d = np.abs(host(res["logits"]).astype(np.float64) - ref_tf)
    assert d.mean() < 3e-3 * np.abs(ref_tf).max(), "mean logit error %.3e" % d.mean()
    # greedy, with and without early exit
    ref = orc.attention_decode(enc, sd_cfg_a, steps, 97, dt=np.float64, quant="bf16")
    full = ops.attn_decode_bf16(dev(enc, bf), dev(proj, bf), pk, steps, None, early_exit=False)
    early = ops.attn_decode_bf16(dev(enc, bf), dev(proj, bf), pk, steps, None, early_exit=True)
    torch.cuda.synchronize()
    ids, ref_ids = host(full["ids"]), ref.argmax(2)
    srt = np.sort(ref, axis=2)
    margin = srt[:, :, -1] - srt[:, :, -2]
    # compare step by step until the first disagreement; a disagreement is only legitimate at a thin margin
    for b in range(B):
        neq = np.nonzero(ids[b] != ref_ids[b])[0]
        if len(neq):
            k = neq[0]
            assert margin[b, k] < 5e-2 * np.abs(ref).max(), "crop %d step %d differs at margin %.3e" % (b, k, margin[b, k])
    assert (ids == ref_ids).all(axis=1).mean() > (0.7 if B >= 30 else 0.35)
    L = host(full["lengths"])
    first = np.where((ids == 1).any(1), (ids == 1).argmax(1), steps)
    assert np.array_equal(L, first) and np.array_equal(host(early["lengths"]), L)
    fl, el, ei = host(full["logits"]), host(early["logits"]), host(early["ids"])
    for b in range(B):
        k = min(int(L[b]) + 1, steps)
        assert np.array_equal(el[b, :k], fl[b, :k]) and np.array_equal(ei[b, :k], ids[b, :k])
        assert not el[b, k:].any() and not ei[b, k:].any()


# ------------------------------------------------------------------------------------------------
# error behaviour of the boundary
# ------------------------------------------------------------------------------------------------
def test_errors_are_loud():
    from multimodal_scene_text_recognition_b200._cabi import MstrError
    x = torch.zeros((1, 4, 4, 24), device=DEV)
    w = torch.zeros((8, 3, 3, 24), device=DEV)
    with pytest.raises(MstrError, match="multiple of 16"):
        ops.conv_gemm(x, w, None, None, pad=(1, 1))
    with pytest.raises(MstrError, match="CUDA tensor"):
        ops.maxpool_nhwc(torch.zeros(1, 4, 4, 8))


# ------------------------------------------------------------------------------------------------
# input side: Resize((32,100)) + ToTensor on ragged uint8 crops
# ------------------------------------------------------------------------------------------------
def test_resize_to_tensor_bit_exact_with_pillow_restatement(golden_dir):
    """CUDA kernel vs oracle/resize_oracle.py (pinned to Pillow) and vs the committed Pillow outputs: ragged
    batch incl. 1x1, identity (32x100), single-axis, strongly minified crops; bit-exact."""
    import sys
    sys.path.insert(0, golden_dir)
    import make_resize_golden as mg
    from oracle import resize_oracle as ro
    from multimodal_scene_text_recognition_b200 import pipeline
    g = np.load(os.path.join(golden_dir, "resize_cases.npz"))
    crops = [mg.make_crop(1000 + i, int(h), int(w)) for i, (h, w) in enumerate(g["shapes"])]
    out = pipeline.crops_to_input(crops, DEV)
    torch.cuda.synchronize()
    got = host(out)
    assert got.shape == (len(crops), 1, 32, 100) and got.dtype == np.float32
    for i, c in enumerate(crops):
        assert np.array_equal(got[i], ro.resize_to_tensor(c)), "crop %d %s" % (i, c.shape)
        assert np.array_equal(np.rint(got[i, 0] * 255.0).astype(np.uint8), g["resized_%d" % i])
    assert np.array_equal(got[0], g["to_tensor_0"])
    # a larger ragged batch with random sizes
    rs = np.random.RandomState(9)
    many = [(rs.rand(int(rs.randint(1, 120)), int(rs.randint(1, 500))) * 256).astype(np.uint8) for _ in range(300)]
    got = host(pipeline.crops_to_input(many, DEV))
    for i in range(0, 300, 7):
        assert np.array_equal(got[i], ro.resize_to_tensor(many[i]))


def test_resize_to_tensor_flags_oversized_crop():
    px = torch.zeros(64, dtype=torch.uint8, device=DEV)
    off = torch.zeros(1, dtype=torch.int64, device=DEV)
    sz = torch.tensor([[8, 8]], dtype=torch.int32, device=DEV)
    out = ops.resize_to_tensor_u8(px, off, sz, (4, 4))      # crop larger than the declared bound: NaN, no overrun
    torch.cuda.synchronize()
    assert torch.isnan(out).all()


def test_gemm_kernels_bit_identical_and_one_wave_rule():
    """mstr_conv_gemm_bf16 picks the single-CTA kernel for problems of at most one wave of 128 x 128 tiles and the CTA-pair
    kernel beyond; the choice depends on the batch, so the two kernels must give bit-identical results (bias, ReLU,
    residual, 1x1 and 3x3, bf16 and f32 outputs) - otherwise a crop's result would depend on its batch."""
    r = rng(61)
    bf = torch.bfloat16
    cases = [((512, 1, 1, 256), (768, 1, 1, 256)), ((4096, 1, 1, 2048), (256, 1, 1, 2048)), ((300, 1, 1, 512), (2048, 1, 1, 512)),
             ((6, 4, 26, 256), (512, 3, 3, 256)), ((40, 8, 25, 128), (256, 3, 3, 128))]
    try:
        for xs, ws in cases:
            x = dev(orc.round_bf16(r.standard_normal(xs).astype(np.float32)), bf)
            w = dev(orc.round_bf16((r.standard_normal(ws) / np.sqrt(ws[1] * ws[2] * ws[3])).astype(np.float32)), bf)
            sc, sh = dev(r.uniform(0.5, 1.5, ws[0]).astype(np.float32)), dev(r.normal(0, 0.2, ws[0]).astype(np.float32))
            pad = (ws[1] // 2, ws[2] // 2)
            res = dev(orc.round_bf16(r.standard_normal((xs[0], xs[1], xs[2], ws[0])).astype(np.float32)), bf)
            outs = []
            for route in (1, 2, 0):
                ops.gemm_route(route)
                ob, of = ops.conv_gemm(x, w, sc, sh, res, pad=pad, relu=True, want_f32=True, want_bf16=True)
                outs.append((ob.clone(), of.clone()))
            torch.cuda.synchronize()
            for ob, of in outs[1:]:
                assert torch.equal(ob, outs[0][0]) and torch.equal(of, outs[0][1]), (xs, ws)
    finally:
        ops.gemm_route(0)
