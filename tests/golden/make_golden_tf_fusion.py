"""Golden fixture of the semantic-fusion switches PRE_ENCODER_MLP + PRE_DECODER_MLP (config.py:43,47) from the reference.

Run in the BUILD container only:  python tests/golden/make_golden_tf_fusion.py  ->  tests/golden/cfg_tf_fusion.npz

Reference model with ENCODER = DECODER = 'Transformer' and both switches on (they are read at construction), loaded with
``synth.make_state_dict_transformer(999, pre_encoder_mlp=True, pre_decoder_mlp=True)``, fp32 on CPU.  The encoder is
driven like in make_golden_tf.py (``TF_Encoder.forward`` raises under torch 2.11): its fusion prologue is the four
statements of encoders.py:128-134 on the reference's own ``get_relevant_semantic`` / ``combine_mlp`` objects, then
pos_encoder, the layers, the norm; ``TF_Decoder.forward`` runs unmodified (with the semantic vectors)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import config  # noqa: E402

config.ENCODER, config.DECODER, config.PRIMARY_DEVICE, config.DEVICE_IDS = "Transformer", "Transformer", "cpu", [0]
config.PRE_ENCODER_MLP = True
config.PRE_DECODER_MLP = True

import torch  # noqa: E402
import model as ref_model  # noqa: E402

from multimodal_scene_text_recognition_b200 import synth  # noqa: E402

B = 8


def main():
    torch.set_num_threads(8)
    sd = synth.make_state_dict_transformer(999, pre_encoder_mlp=True, pre_decoder_mlp=True)
    m = ref_model.Model().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = torch.from_numpy(synth.make_crops(123, B))
    ov, sc, io = [torch.from_numpy(a) for a in synth.make_semantics(124, B)]
    text = torch.zeros(B, 26, dtype=torch.long)
    with torch.no_grad():
        rect = m.Transformation(x)
        vis = m.AdaptiveAvgPool(m.FeatureExtraction(rect).permute(0, 3, 1, 2)).squeeze(3)
        sem = m.get_semantic_vectors(ov, sc, io)                                       # [B, 15, 256]
        enc = m.encoder
        rel = enc.get_relevant_semantic(vis, sem, is_train=False)                      # encoders.py:128
        fused = vis + enc.combine_mlp(torch.cat((vis, rel), dim=2))                    # encoders.py:130-134
        h = enc.pos_encoder(fused.permute(1, 0, 2))
        for mod in enc.encoder.layers:
            h = mod(h)
        out = enc.encoder.norm(h).permute(1, 0, 2).contiguous()
        logits = m.decoder(out, text=text, semantics=sem, is_train=False)
        logits_plain = m.decoder(out, text=text, semantics=torch.zeros_like(sem), is_train=False)
    ids = logits.argmax(2).numpy().astype(np.int32)
    print("distinct strings %d / %d; ids change with the semantic vectors on %.2f of the steps; relevance std %.3f" % (
        len({tuple(r) for r in ids}), B, float((ids != logits_plain.argmax(2).numpy()).mean()), rel.std().item()))
    np.savez_compressed(os.path.join(HERE, "cfg_tf_fusion.npz"), visual=vis.numpy(), semantics=sem.numpy(), relevant=rel.numpy(),
                        fused=fused.numpy(), enc=out.numpy(), logits=logits.numpy(), ids=ids)


if __name__ == "__main__":
    main()
