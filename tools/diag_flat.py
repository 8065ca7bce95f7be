"""conv0_1 (ring) -> conv3x3_flat -> maxpool (ring) at the bench geometry, timed in isolation.
Usage (GPU box): python tools/diag_flat.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(3)
x = torch.rand(B, 1, 32, 100, device=dev, generator=g)
w1 = torch.randn(32, 1, 3, 3, device=dev, generator=g) / 3
s1 = torch.rand(32, device=dev, generator=g) + 0.5
b1 = torch.randn(32, device=dev, generator=g) * 0.1
w2 = (torch.randn(64, 3, 3, 32, device=dev, generator=g) / 17).to(torch.bfloat16)
s2 = torch.rand(64) + 0.5
b2 = torch.randn(64) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


ring = ops.conv3x3_first(x, w1, s1, b1, torch.bfloat16, pool=False, ring=True)
out = ops.conv3x3_flat(ring, w2, s2, b2)
print("B=%d: conv0_1 (ring) %.3f ms | conv3x3_flat %.3f ms | maxpool (ring) %.3f ms" % (
    B, timeit(lambda: ops.conv3x3_first(x, w1, s1, b1, torch.bfloat16, pool=False, ring=True)),
    timeit(lambda: ops.conv3x3_flat(ring, w2, s2, b2)), timeit(lambda: ops.maxpool2x2_ring(out))))
