"""Golden fixtures of the transformer family from the reference's own module objects.

Run in the BUILD container only (it imports the unmodified reference from /root/reference):

    python tests/golden/make_golden_tf.py        ->  tests/golden/cfg_tf.npz

The reference model is built with its default ``ENCODER = DECODER = 'Transformer'`` (config.py:27-28), loaded with
``synth.make_state_dict_transformer(999)`` and run in fp32 on CPU on ``synth.make_crops(123, B)``.
``Model.forward`` cannot be called as a whole: under torch 2.11 ``nn.TransformerEncoder.forward`` passes
``is_causal=`` to the reference's ``TransformerEncoderLayer.forward`` (encoders.py:185), which raises TypeError (checked
below).  The encoder is therefore driven from the loop ``nn.TransformerEncoder.forward`` consisted of when the reference
was written - ``for mod in layers: x = mod(x)``, then ``norm`` - over the reference's OWN ``pos_encoder``, layer and norm
objects; ``TF_Decoder.forward`` (greedy and teacher-forced) and ``Linear_Decoder.forward`` run unmodified.
Weights and inputs are not stored: both are regenerated from seeds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import config  # noqa: E402  (reference config; patched before importing model)

config.ENCODER = "Transformer"
config.DECODER = "Transformer"
config.PRIMARY_DEVICE = "cpu"
config.DEVICE_IDS = [0]

import torch  # noqa: E402
import model as ref_model  # noqa: E402
from modules.decoders import Linear_Decoder  # noqa: E402

from multimodal_scene_text_recognition_b200 import synth  # noqa: E402

WEIGHT_SEED, DATA_SEED, B = 999, 123, 8


def main():
    torch.set_num_threads(8)
    sd = synth.make_state_dict_transformer(WEIGHT_SEED)
    m = ref_model.Model().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = torch.from_numpy(synth.make_crops(DATA_SEED, B))
    ov, sc, io = [torch.from_numpy(a) for a in synth.make_semantics(DATA_SEED + 1, B)]
    text = torch.from_numpy(np.random.Generator(np.random.PCG64(DATA_SEED + 2)).integers(3, 97, (B, 26)).astype(np.int64))
    text[:, 0] = 0
    with torch.no_grad():
        try:
            m(x, text, ov, sc, io, is_train=False)
            whole_model = "runs"
        except TypeError as e:
            whole_model = "TypeError: %s" % e
        rect = m.Transformation(x)
        feat = m.FeatureExtraction(rect)
        vis = m.AdaptiveAvgPool(feat.permute(0, 3, 1, 2)).squeeze(3)
        h = m.encoder.pos_encoder(vis.permute(1, 0, 2))
        layer0 = None
        for i, mod in enumerate(m.encoder.encoder.layers):
            h = mod(h)
            if i == 0:
                layer0 = h.permute(1, 0, 2).clone()
        enc = m.encoder.encoder.norm(h).permute(1, 0, 2).contiguous()
        logits = m.decoder(enc, text=text, semantics=None, is_train=False)            # greedy, [B, 25, 97]
        logits_tf = m.decoder(enc, text=text, semantics=None, is_train=True)          # teacher forced, [B, 26, 97]
        sd_lin = synth.make_state_dict_transformer(WEIGHT_SEED, decoder="Linear")
        lin = Linear_Decoder(97).eval()
        lin.load_state_dict({k[len("decoder."):]: torch.from_numpy(v) for k, v in sd_lin.items() if k.startswith("decoder.")})
        logits_lin = lin(enc, None, None, False)
    ids = logits.argmax(2).numpy().astype(np.int32)
    print("Model.forward as a whole under torch %s: %s" % (torch.__version__, whole_model))
    print("greedy ids:\n", ids[:4])
    print("distinct strings %d / %d" % (len({tuple(r) for r in ids}), B))
    np.savez_compressed(os.path.join(HERE, "cfg_tf.npz"), visual=vis.numpy(), enc_layer0=layer0.numpy(), enc=enc.numpy(),
                        logits=logits.numpy(), ids=ids, text=text.numpy(), logits_tf=logits_tf.numpy(),
                        logits_linear=logits_lin.numpy(), whole_model=np.array(whole_model),
                        torch_version=np.array(torch.__version__))


if __name__ == "__main__":
    main()
