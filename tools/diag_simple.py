"""Runs the memory-bound helper kernels once at the bench geometry (for ncu) and times them with CUDA events.
Usage (GPU box): python tools/diag_simple.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = float(np.median(ts))
    print("%-34s %.3f ms  %.0f GB/s" % (name, t, nbytes / t / 1e6))


img = torch.rand(B, 1, 32, 100, device=dev)
w32 = torch.randn(32, 1, 3, 3, device=dev); w64 = torch.randn(64, 1, 3, 3, device=dev)
sc32, sh32 = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev)
sc64, sh64 = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
timeit("conv3x3_first 1->32", lambda: ops.conv3x3_first(img, w32, sc32, sh32, bf, False), B * 3200 * (4 + 64))
timeit("conv3x3_first 1->64 +pool", lambda: ops.conv3x3_first(img, w64, sc64, sh64, bf, True), B * (3200 * 4 + 800 * 128))
for (H, W, C, k, s, p) in ((32, 100, 64, (2, 2), (2, 2), (0, 0)), (16, 50, 128, (2, 2), (2, 2), (0, 0)),
                           (8, 25, 256, (2, 2), (2, 2), (0, 0)), (8, 25, 256, (2, 2), (2, 1), (0, 1))):
    x = torch.randn(B, H, W, C, device=dev).to(bf)
    P = (H + 2 * p[0] - 2) // s[0] + 1; Q = (W + 2 * p[1] - 2) // s[1] + 1
    timeit("maxpool %dx%dx%d s%s p%s" % (H, W, C, s, p), lambda: ops.maxpool_nhwc(x, k, s, p), 2 * B * C * (H * W + P * Q))
    del x
# input side: 4096 ragged grayscale crops (heights 16..96, widths 40..400) -> [B,1,32,100] f32
from multimodal_scene_text_recognition_b200 import pipeline
rs = np.random.RandomState(0)
crops = [(rs.rand(int(rs.randint(16, 97)), int(rs.randint(40, 401))) * 256).astype(np.uint8) for _ in range(B)]
pixels, offsets, sizes, max_hw = pipeline.pack_crops(crops)
pd, od, sd_ = pixels.to(dev), offsets.to(dev), sizes.to(dev)
timeit("resize_to_tensor_u8 (%.1f MB of crops)" % (pixels.numel() / 1e6),
       lambda: ops.resize_to_tensor_u8(pd, od, sd_, max_hw), pixels.numel() + B * 3200 * 4)
F = 20
from multimodal_scene_text_recognition_b200.modules.transformation import GridGenerator
gg = GridGenerator(F, (32, 100)).to(dev)
cp = (torch.rand(B, F, 2, device=dev) * 2 - 1)
basis = gg.P_hat.t().contiguous() if hasattr(gg, "P_hat") else None
try:
    from multimodal_scene_text_recognition_b200.modules import transformation as tr
    tps = tr.TPS_SpatialTransformerNetwork(F, (32, 100), (32, 100), 1).to(dev).eval()
    with torch.no_grad():
        grid_in = tps.GridGenerator
        timeit("tps_rectify (via module, +LocNet)", lambda: tps(img), B * 3200 * 8)
except Exception as ex:  # noqa
    print("tps skipped:", ex)
