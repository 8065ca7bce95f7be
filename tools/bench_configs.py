"""Device-timed crops/s of the other BASELINE.json configurations (they are parity-test cases, not bench lines):
configs[1] cfg-A at batch 256 (bf16 and the 3xTF32 parity mode) and configs[4] cfg-S (3x64x256 crops, 51 decode
steps).  Usage (GPU box): python tools/bench_configs.py > gpurun_out/other_configs.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import config, model as M, synth

dev = torch.device("cuda", 0)


def run(geom, B, precision, reps=5):
    config.PRECISION = precision
    config.MAX_TEXT_LENGTH = geom.max_len
    config.DECODER_EARLY_EXIT = False
    config.TPS_GRID = "reference"
    config.SEMANTIC_VECTOR = "overlap"
    sd = synth.make_state_dict(999, geom)
    m = M.load_numpy_state_dict(M.Model(imgH=geom.H, imgW=geom.W, input_channel=geom.in_ch).eval(), sd).to(dev)
    x = torch.from_numpy(synth.make_crops(123, B, geom)).to(dev)
    sem = [torch.from_numpy(a).to(dev) for a in synth.make_semantics(124, B)]
    with torch.no_grad():
        for _ in range(3):
            m.recognize(x, *sem)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); m.recognize(x, *sem); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    return {"geometry": "%dx%dx%d, %d decode steps" % (geom.in_ch, geom.H, geom.W, geom.steps), "batch": B,
            "precision": precision, "ms_per_batch": round(ms, 3), "crops_per_s": round(B / ms * 1e3, 1)}


out = [run(synth.CFG_A, 256, "bf16"), run(synth.CFG_A, 256, "tf32"), run(synth.CFG_A, 4096, "tf32", reps=2),
       run(synth.CFG_S, 256, "bf16"), run(synth.CFG_S, 1024, "bf16", reps=3)]
for r in out:
    print(json.dumps(r))
