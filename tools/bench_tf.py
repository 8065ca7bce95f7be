"""Throughput of the transformer configurations at the bench geometry: eager launches and CUDA-graph
replay, plus the per-kernel-family breakdown of one eager step.  Usage (GPU box): python tools/bench_tf.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import config, model as M, ops, synth, _cabi
from multimodal_scene_text_recognition_b200.pipeline import GraphedRecognizer

BATCHES = [int(a) for a in sys.argv[1:]] or [4096]
dev = torch.device("cuda")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


for decoder in ("Transformer", "Linear"):
    config.ENCODER, config.DECODER, config.PRECISION = "Transformer", decoder, "bf16"
    m = M.Model().eval()
    M.load_numpy_state_dict(m, synth.make_state_dict_transformer(999, decoder=decoder))
    m = m.to(dev)
    for B in BATCHES:
        x = torch.from_numpy(synth.make_crops(123, B)).to(dev)
        sem = [torch.from_numpy(a).to(dev) for a in synth.make_semantics(124, B)]
        with torch.no_grad():
            n0 = _cabi.launch_count()
            m.recognize(x, *sem)
            launches = _cabi.launch_count() - n0
            eager = timeit(lambda: m.recognize(x, *sem))
            g = GraphedRecognizer(m)
            graph = timeit(lambda: g(x, *sem))
            ops.TIMER = ops.KernelTimer()
            m.recognize(x, *sem)
            torch.cuda.synchronize()
            kern = ops.TIMER.summary()
            ops.TIMER = None
        print("Transformer encoder + %s decoder, B=%d bf16: %d launches per forward | eager %.2f ms (%.0f crops/s) | graph replay "
              "%.2f ms (%.0f crops/s)" % (decoder, B, launches, eager, B / eager * 1e3, graph, B / graph * 1e3))
        print("   per kernel family (eager, timed launches only): " + ", ".join(
            "%s %.2f ms / %d" % (k, v["ms"], v["launches"]) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])[:9]))
