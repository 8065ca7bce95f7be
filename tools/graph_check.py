"""CUDA-graph replay of the forward vs eager launches: same results, device time and host time per step.
Usage (GPU box): python tools/graph_check.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import config, model as M, synth
from multimodal_scene_text_recognition_b200.pipeline import GraphedRecognizer

dev = torch.device("cuda", 0)
config.PRECISION, config.DECODER_EARLY_EXIT = "bf16", False
m = M.load_numpy_state_dict(M.Model().eval(), synth.make_state_dict(999)).to(dev)
g = GraphedRecognizer(m)
for B in [int(a) for a in sys.argv[1:]] or [512]:
    x = torch.from_numpy(synth.make_crops(5, B)).to(dev)
    sem = [torch.from_numpy(a).to(dev) for a in synth.make_semantics(6, B)]
    with torch.no_grad():
        ref = m.recognize(x, *sem)
        ref = {k: ref[k].clone() for k in ("logits", "ids", "lengths")}
    out = g(x, *sem)
    torch.cuda.synchronize()
    same = all(torch.equal(out[k], ref[k]) for k in ref)
    x2 = torch.from_numpy(synth.make_crops(7, B)).to(dev)
    with torch.no_grad():
        ref2 = m.recognize(x2, *sem)["ids"].clone()
    same2 = torch.equal(g(x2, *sem)["ids"], ref2)

    def timed(fn, reps=20):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); s.record()
        for _ in range(reps):
            fn()
        e.record(); host = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps, host * 1e3

    with torch.no_grad():
        eager = timed(lambda: m.recognize(x, *sem))
    graph = timed(lambda: g(x, *sem))
    print("B=%d: graph == eager %s / %s | eager %.3f ms device (%.3f ms host enqueue) | graph %.3f ms device (%.3f ms host)"
          % (B, same, same2, eager[0], eager[1], graph[0], graph[1]))
