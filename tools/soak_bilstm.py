"""Race detector for the cluster BiLSTM kernel: its output must equal the single-CTA kernel's bit for bit (same
operands, same accumulation order), over many launches with fresh random data.  Usage: python tools/soak_bilstm.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_scene_text_recognition_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
g = torch.Generator(device="cuda").manual_seed(7)
bad = 0
for i in range(reps):
    B = [768, 129, 1, 640, 257][i % 5]
    T = [26, 65, 3, 26, 10][i % 5]
    gx = torch.randn(B, T, 2, 1024, device="cuda", generator=g) * (1.0 + (i % 3))
    w = ops.pack_w_hh_bf16(torch.randn(2, 1024, 256, device="cuda", generator=g) / 16)
    ops.set_bilstm_kernel("single")
    a = ops.bilstm_recurrent_bf16(gx, w)
    ops.set_bilstm_kernel("cluster")
    b = ops.bilstm_recurrent_bf16(gx, w)
    torch.cuda.synchronize()
    if not torch.equal(a, b):
        bad += 1
        d = (a.float() - b.float()).abs()
        print("rep %d B=%d T=%d: %d of %d elements differ, max %.3e" % (i, B, T, int((d > 0).sum()), d.numel(), float(d.max())))
print("soak: %d launches, %d mismatching" % (reps, bad))
