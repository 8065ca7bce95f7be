"""One warm forward of the Transformer / Transformer configuration (for ncu launch lists): python tools/tf_one_forward.py B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_scene_text_recognition_b200 import config, model as M, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
config.ENCODER, config.DECODER, config.PRECISION = "Transformer", "Transformer", "bf16"
m = M.Model().eval()
M.load_numpy_state_dict(m, synth.make_state_dict_transformer(999))
m = m.to("cuda")
x = torch.from_numpy(synth.make_crops(123, B)).cuda()
sem = [torch.from_numpy(a).cuda() for a in synth.make_semantics(124, B)]
with torch.no_grad():
    for _ in range(2):
        out = m.recognize(x, *sem)
torch.cuda.synchronize()
print(out["ids"][0].tolist())
