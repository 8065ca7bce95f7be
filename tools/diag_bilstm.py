"""One launch of the bf16 BiLSTM recurrence (for ncu).  Usage: python tools/diag_bilstm.py B T"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_scene_text_recognition_b200 import ops
B, T = int(sys.argv[1]), int(sys.argv[2])
gx = torch.randn(B, T, 2, 1024, device="cuda")
w = torch.randn(2, 1024, 256, device="cuda") / 16
out = ops.bilstm_recurrent_bf16(gx, ops.pack_w_hh_bf16(w))
torch.cuda.synchronize()
print("ok", B, T, float(out.float().abs().mean()))
