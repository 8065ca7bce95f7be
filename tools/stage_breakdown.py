"""Per-kernel-family time of one recognition pass for a geometry / batch / precision (CUDA events around every C-ABI
call, like bench.py's stage_ms_per_step).  Usage: python tools/stage_breakdown.py [a|s] [B] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_scene_text_recognition_b200 import config, model as M, ops, synth
geom = synth.CFG_S if (len(sys.argv) > 1 and sys.argv[1] == "s") else synth.CFG_A
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
config.PRECISION = sys.argv[3] if len(sys.argv) > 3 else "bf16"
config.MAX_TEXT_LENGTH = geom.max_len
config.DECODER_EARLY_EXIT = False
dev = torch.device("cuda", 0)
m = M.load_numpy_state_dict(M.Model(imgH=geom.H, imgW=geom.W, input_channel=geom.in_ch).eval(), synth.make_state_dict(999, geom)).to(dev)
x = torch.from_numpy(synth.make_crops(123, B, geom)).to(dev)
sem = [torch.from_numpy(a).to(dev) for a in synth.make_semantics(124, B)]
with torch.no_grad():
    for _ in range(3):
        m.recognize(x, *sem)
    torch.cuda.synchronize()
    ops.TIMER = ops.KernelTimer()
    reps = 3
    for _ in range(reps):
        m.recognize(x, *sem)
    torch.cuda.synchronize()
    t, ops.TIMER = ops.TIMER, None
tot = 0.0
for k, v in sorted(t.summary().items(), key=lambda kv: -kv[1]["ms"]):
    print("%-26s %8.3f ms  %4d launches" % (k, v["ms"] / reps, v["launches"] // reps)); tot += v["ms"] / reps
print("%-26s %8.3f ms  (%s, B=%d, %s) -> %.0f crops/s" % ("sum", tot, "cfg-S" if geom is synth.CFG_S else "cfg-A", B, config.PRECISION, B / tot * 1e3))
