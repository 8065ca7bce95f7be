"""Golden fixture of the decoder-side switches CLS_DECODER_INIT + POST_DECODER_MLP (config.py:48,53) from the reference's
unmodified TF_Decoder.forward.  Run in the BUILD container only:  python tests/golden/make_golden_tf_fusion2.py
-> tests/golden/cfg_tf_fusion2.npz.  Encoder output: the `enc` of cfg_tf.npz (same seeds); semantic vectors: the
Linear_Embedding of synth.make_semantics(124, B)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import config  # noqa: E402

config.ENCODER, config.DECODER, config.PRIMARY_DEVICE, config.DEVICE_IDS = "Transformer", "Transformer", "cpu", [0]
config.CLS_DECODER_INIT = True
config.POST_DECODER_MLP = True

import torch  # noqa: E402
from modules.decoders import TF_Decoder  # noqa: E402

from multimodal_scene_text_recognition_b200 import synth  # noqa: E402


def main():
    torch.set_num_threads(8)
    sd = synth.make_state_dict_transformer(999, cls_decoder_init=True, post_decoder_mlp=True)
    dec = TF_Decoder(97).eval()
    dec.load_state_dict({k[len("decoder."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    g = np.load(os.path.join(HERE, "cfg_tf.npz"))
    enc = torch.from_numpy(g["enc"])
    B = enc.shape[0]
    ov, _sc, _io = synth.make_semantics(124, B)
    sem = torch.from_numpy(sd["get_semantic_vectors.embed.weight"][ov.astype(np.int64)])        # Linear_Embedding 'overlap'
    with torch.no_grad():
        cls = dec.get_semantic_cls(dec.hid_to_emb(enc), sem, False)
        logits = dec(enc, text=torch.from_numpy(g["text"]), semantics=sem, is_train=False)
        logits_tf = dec(enc, text=torch.from_numpy(g["text"]), semantics=sem, is_train=True)
    print("semantic cls: min %.8f max %.8f (a vector of ones)" % (cls.min().item(), cls.max().item()))
    ids = logits.argmax(2).numpy().astype(np.int32)
    print("distinct strings %d / %d; differ from the plain decoder's ids on %.2f of the steps" % (
        len({tuple(r) for r in ids}), B, float((ids != g["ids"]).mean())))
    np.savez_compressed(os.path.join(HERE, "cfg_tf_fusion2.npz"), semantics=sem.numpy(), logits=logits.numpy(), ids=ids,
                        logits_tf=logits_tf.numpy(), cls_min=np.array(cls.min().item()), cls_max=np.array(cls.max().item()))


if __name__ == "__main__":
    main()
