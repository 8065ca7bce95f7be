"""CPU sweep of the synthetic-weight gains: decode diversity vs sensitivity to bf16 rounding.
For each candidate: fp32 oracle vs bf16-quantised oracle on the same crops."""
import itertools, json, sys, time
import numpy as np
sys.path.insert(0, ".")
from multimodal_scene_text_recognition_b200 import synth
from oracle import recognition_oracle as orc

B = 24
x = synth.make_crops(777, B)
cands = []
for enc_ih, enc_lin, dec_ih, i2h, score in [
        (3, 8, 3, 8, 8), (3, 8, 3, 4, 4), (3, 8, 3, 2, 2), (3, 4, 3, 4, 4), (2, 4, 2, 4, 4), (3, 8, 3, 3, 8),
        (3, 8, 3, 8, 3), (3, 4, 3, 2, 4), (2, 8, 2, 2, 2), (3, 6, 2, 3, 3), (3, 8, 2, 4, 2), (4, 8, 3, 2, 4)]:
    cands.append(dict(enc_w_ih=enc_ih, enc_linear=enc_lin, dec_w_ih=dec_ih, dec_i2h=i2h, dec_score=score))
feat_cache = {}
for g in cands:
    t0 = time.time()
    synth.SEQUENCE_GAINS.update(g)
    sd = synth.make_state_dict(999)
    a = orc.forward(sd, x, stages=True)
    q = orc.forward(sd, x, quant="bf16", stages=True)
    ids = a["ids"]
    uniq = len({tuple(r) for r in ids})
    chars = len(set(ids.flatten().tolist()))
    stop = float((ids == 1).any(1).mean())
    lg = a["logits"]
    srt = np.sort(lg, axis=2)
    margin = float(np.median(srt[:, :, -1] - srt[:, :, -2]) / lg.std())
    rel = float(np.abs(q["logits"] - lg).max() / np.abs(lg).max())
    same = int((q["ids"] == ids).all(1).sum())
    steps_same = float((q["ids"] == ids).mean())
    frel = float(np.abs(q["features"] - a["features"]).max() / np.abs(a["features"]).max())
    erel = float(np.abs(q["enc1"] - a["enc1"]).max() / np.abs(a["enc1"]).max())
    print(json.dumps(dict(g=g, uniq=uniq, chars=chars, stop=stop, margin=round(margin, 3), feat=round(frel, 4),
                          enc1=round(erel, 4), logits=round(rel, 4), ids_same=same, steps_same=round(steps_same, 3),
                          sec=round(time.time() - t0, 1))), flush=True)
