"""Per-launch time of small linear layers (decoder-step shapes) through ops.linear, measured as a CUDA graph of 50
back-to-back launches (no host launch overhead).  Usage (GPU box): python tools/bench_small_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_scene_text_recognition_b200 import ops

dev = torch.device("cuda")
REPS = 50


def graph_time(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / REPS * 1e3


for M in (512, 4096):
    for K, N in ((256, 768), (256, 256), (256, 2048), (2048, 256), (256, 128), (512, 1536), (512, 2048), (2048, 512)):
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        b = torch.zeros(N, device=dev)
        r = torch.randn(M, N, device=dev).to(torch.bfloat16)
        t = graph_time(lambda: ops.linear(x, w, b, residual=r))
        print("M=%5d K=%4d N=%4d: %.1f us per launch (%.1f TFLOP/s)" % (M, K, N, t, 2.0 * M * K * N / t / 1e6))
    x = torch.randn(M, 256, device=dev).to(torch.bfloat16)
    g_, b_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    print("M=%5d layer_norm 256: %.1f us" % (M, graph_time(lambda: ops.layer_norm(x, g_, b_))))
    kc = torch.randn(M, 25, 256, device=dev).to(torch.bfloat16); vc = torch.randn(M, 25, 256, device=dev).to(torch.bfloat16)
    print("M=%5d mha_step Lk=25: %.1f us" % (M, graph_time(lambda: ops.mha_step(x, kc, vc, 25, 8))))
