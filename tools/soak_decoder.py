"""Race detector for the persistent tensor-core decoder (grid barriers, split arrive / wait barrier, dynamic crop hand-out):
its ids / logits / lengths must equal the launch chain's bit for bit, over many launches with fresh random data, batch
sizes on both sides of every dispatch boundary, greedy with and without early exit, teacher forcing, and with other work
queued on the stream right behind it.  Usage: python tools/soak_decoder.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_scene_text_recognition_b200 import ops, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(999).items() if k.startswith("decoder.")}
p = "decoder.attention_cell."
H = 256
w_ih = sd[p + "rnn.weight_ih"]
C = sd["decoder.generator.weight"].shape[0]
head = torch.zeros((384, H), device=dev)
head[:H] = sd[p + "h2h.weight"]; head[H:H + C] = sd["decoder.generator.weight"]
pk = {"w_head": head.to(torch.bfloat16).contiguous(),
      "w_cat": ops.pack_gate_rows_bf16(torch.cat([w_ih[:, :H], sd[p + "rnn.weight_hh"]], 1)),
      "w_tok": w_ih[:, H:].t().contiguous(), "bias": (sd[p + "rnn.bias_ih"] + sd[p + "rnn.bias_hh"]).contiguous(),
      "b_h2h": sd[p + "h2h.bias"], "b_gen": sd["decoder.generator.bias"],
      "w_score": sd[p + "score.weight"].view(-1).contiguous()}
g = torch.Generator(device=dev).manual_seed(3)
sizes = [1, 127, 128, 129, 300, 512, 592, 593, 1100, 1184, 1185, 1700, 2368, 2369, 4096]
bad = 0
for i in range(reps):
    B = sizes[i % len(sizes)]
    T, steps = (26, 26) if i % 4 else (65, 51)
    enc = (torch.randn(B, T, H, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    proj = (enc.float() @ sd[p + "i2h.weight"].t()).to(torch.bfloat16)
    mode = i % 3
    text = torch.randint(0, C, (B, steps), device=dev, generator=g) if mode == 2 else None
    ee = mode == 1
    outs = []
    for persistent in (False, True):
        ops.set_decode_persistent(persistent)
        r = ops.attn_decode_bf16(enc, proj, pk, steps, text, early_exit=ee)
        junk = torch.zeros(1 << 22, device=dev).add_(1.0)          # work queued right behind the call
        outs.append((r["logits"].clone(), r["ids"].clone(), r["lengths"].clone()))
    torch.cuda.synchronize()
    ok = all(torch.equal(a, b) for a, b in zip(*outs))
    if not ok:
        bad += 1
        print("rep %d B=%d T=%d mode=%d: persistent != chain" % (i, B, T, mode))
ops.set_decode_persistent(None)
print("soak: %d decodes (x2 drivers), %d mismatching" % (reps, bad))
