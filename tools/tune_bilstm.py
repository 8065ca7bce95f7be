"""BiLSTM recurrence: one CTA per (crop tile, direction) vs clusters of 8 CTAs, per batch size.  GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_scene_text_recognition_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))
w = torch.randn(2, 1024, 256, device=dev) / 16
wp = ops.pack_w_hh_bf16(w)
for B in [int(a) for a in sys.argv[1:]] or [512, 1024, 1536, 2048, 3072, 4096]:
    gx = torch.randn(B, 26, 2, 1024, device=dev)
    r = {}
    for k in ("single", "cluster"):
        ops.set_bilstm_kernel(k)
        r[k] = timeit(lambda: ops.bilstm_recurrent_bf16(gx, wp))
    ops.set_bilstm_kernel("auto")
    r["auto"] = timeit(lambda: ops.bilstm_recurrent_bf16(gx, wp))
    print("B=%5d: one CTA per tile %.3f ms | cluster of 8 %.3f ms | auto %.3f ms" % (B, r["single"], r["cluster"], r["auto"]))
