"""x-projection GEMM (M=106496, N=2048) timing with f32 / bf16 output and K = 512 / 256, in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_scene_text_recognition_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print("%-40s %.3f ms" % (name, float(np.median(ts))))
M = 106496
for K in (512, 256):
    for N in (2048, 256):
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / 16).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        timeit("K=%d N=%d f32 out" % (K, N), lambda: ops.linear(x, w, b, want_f32=True, want_bf16=False))
        timeit("K=%d N=%d bf16 out" % (K, N), lambda: ops.linear(x, w, b, want_f32=False, want_bf16=True))
