"""Timing of the sequence-stage kernels in isolation (BiLSTM recurrence; tensor-core decoder as one persistent
launch and as the launch chain) at the bench geometry.  Usage (GPU box): python tools/tune_seq.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import ops, synth

BATCHES = [int(a) for a in sys.argv[1:]] or [4096]
T, H, steps = 26, 256, 26
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(min(ts))


sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(999).items() if k.startswith("decoder.")}
p = "decoder.attention_cell."
w_ih = sd[p + "rnn.weight_ih"]
C = sd["decoder.generator.weight"].shape[0]
head = torch.zeros((384, H), device=dev)
head[:H] = sd[p + "h2h.weight"]; head[H:H + C] = sd["decoder.generator.weight"]
pk = {"w_head": head.to(torch.bfloat16).contiguous(),
      "w_cat": ops.pack_gate_rows_bf16(torch.cat([w_ih[:, :H], sd[p + "rnn.weight_hh"]], 1)),
      "w_tok": w_ih[:, H:].t().contiguous(), "bias": (sd[p + "rnn.bias_ih"] + sd[p + "rnn.bias_hh"]).contiguous(),
      "b_h2h": sd[p + "h2h.bias"], "b_gen": sd["decoder.generator.bias"],
      "w_score": sd[p + "score.weight"].view(-1).contiguous()}
w = torch.randn(2, 4 * H, H, device=dev) / 16
wp = ops.pack_w_hh_bf16(w)
for B in BATCHES:
    gx = torch.randn(B, T, 2, 4 * H, device=dev)
    print("bilstm_recurrent_bf16 B=%d: median %.3f ms, best %.3f ms" % ((B,) + timeit(lambda: ops.bilstm_recurrent_bf16(gx, wp))))
    enc = (torch.randn(B, T, H, device=dev) * 0.5).to(torch.bfloat16)
    proj = (enc.float() @ sd[p + "i2h.weight"].t()).to(torch.bfloat16)
    for persistent in (False, True):
        ops.set_decode_persistent(persistent)
        med, best = timeit(lambda: ops.attn_decode_bf16(enc, proj, pk, steps, None, early_exit=False))
        print("attn_decode_bf16 B=%d %s: median %.3f ms, best %.3f ms" % (
            B, "persistent (1 launch)" if persistent else "chain (80 launches)", med, best))
    ops.set_decode_persistent(None)

# phase trace of the persistent decoder (CTA 0): per step [head work, barrier, attention work, barrier, cell work, barrier]
import ctypes
from multimodal_scene_text_recognition_b200 import _cabi
for B in BATCHES:
    enc = (torch.randn(B, T, H, device=dev) * 0.5).to(torch.bfloat16)
    proj = (enc.float() @ sd[p + "i2h.weight"].t()).to(torch.bfloat16)
    ops.set_decode_persistent(True)
    ops.attn_decode_bf16(enc, proj, pk, steps, None, early_exit=False)
    ops.attn_decode_bf16(enc, proj, pk, steps, None, early_exit=False)
    torch.cuda.synchronize()
    ops.set_decode_persistent(None)
    n = 1 + 6 * steps
    buf = (ctypes.c_ulonglong * n)()
    NB, NC = 48, 160
    per = (ctypes.c_ulonglong * (NB * NC))()
    _cabi.check(_cabi.lib().mstr_debug_decode_trace(ctypes.cast(buf, ctypes.c_void_p), n, ctypes.cast(per, ctypes.c_void_p), NB * NC))
    t = np.array(list(buf), dtype=np.float64)
    d = np.diff(t).reshape(steps, 6) / 1e3
    names = ["head work", "barrier 1", "attention", "barrier 2", "cell work", "barrier 3"]
    print("persistent decoder phase trace B=%d (us, median over steps): %s | step total %.1f" % (
        B, ", ".join("%s %.2f" % (nm, v) for nm, v in zip(names, np.median(d, axis=0))), np.median(d.sum(axis=1))))
    # arrival of every CTA at the barriers of steps 3..14, relative to the barrier's first arrival
    arr = np.array(list(per), dtype=np.float64).reshape(NB, NC)[:, :148]
    rel = (arr - arr.min(axis=1, keepdims=True)) / 1e3
    for ph, nm in enumerate(["after head", "after attention", "after cell"]):
        r = rel[9 + ph:45:3]
        late = np.median(r, axis=0)
        order = np.argsort(-late)[:6]
        print("  arrival spread %s: median last-first %.2f us; latest CTAs %s" % (
            nm, np.median(r.max(axis=1)), ", ".join("%d(+%.1f)" % (c, late[c]) for c in order)))
    fine = (ctypes.c_ulonglong * 192)()
    _cabi.check(_cabi.lib().mstr_debug_decode_trace(ctypes.cast(buf, ctypes.c_void_p), n, ctypes.cast(fine, ctypes.c_void_p), 192))
    f = np.array(list(fine), dtype=np.float64).reshape(3, 4, 16)
    slots = ["entry", "tma issued", "stage arrived", "mma committed", "acc complete", "-", "acc released", "epilogue done"]
    for ph, nm in enumerate(["hp", "cell", "logits"]):
        for cta in range(4):
            row = f[ph, cta]
            if row[0] == 0:
                continue
            print("  fine %s tile %d (us since phase entry): %s" % (nm, cta, ", ".join(
                "%s %.2f" % (slots[k], (row[k] - row[0]) / 1e3) for k in (1, 2, 3, 4, 6, 7) if row[k] > 0)))
