"""Golden fixture of the Oscar / BERT encoder from the reference's own module.

Run in the BUILD container only (imports /root/reference):  python tests/golden/make_golden_oscar.py
-> tests/golden/cfg_oscar.npz

``torch.manual_seed(999); Oscar_Bert()`` (the reference's class: hid_to_bert, transformers.BertModel(BertConfig()),
bert_to_hid, encoders.py:41-49), then ``synth.perturb_parameters_(m, 1234)`` so that biases and LayerNorm parameters
are not the trivial 0 / (1, 0) of BertModel's default init, eval mode, fp32 on CPU.  The mirror consumes the RNG
identically, so the test rebuilds the same 203 tensors from the two seeds; only the input and the output are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import config  # noqa: E402

config.ENCODER, config.DECODER, config.PRIMARY_DEVICE, config.DEVICE_IDS = "Oscar", "Transformer", "cpu", [0]

import torch  # noqa: E402
from modules.encoders import Oscar_Bert  # noqa: E402

from multimodal_scene_text_recognition_b200 import synth  # noqa: E402


def main():
    torch.set_num_threads(8)
    torch.manual_seed(999)
    m = Oscar_Bert().eval()
    synth.perturb_parameters_(m, 1234)
    B = 6
    x = np.random.Generator(np.random.PCG64(77)).standard_normal((B, 26, 512)).astype(np.float32)
    with torch.no_grad():
        y = m(torch.from_numpy(x), None, False)
    sd = m.state_dict()
    checks = {k: float(v.double().sum()) for k, v in sd.items() if v.is_floating_point()}
    keys = sorted(checks)
    print("output", tuple(y.shape), "std %.4f" % y.std().item(), "| %d tensors" % len(sd))
    np.savez_compressed(os.path.join(HERE, "cfg_oscar.npz"), x=x, y=y.numpy(), keys=np.array(keys),
                        sums=np.array([checks[k] for k in keys]), n_tensors=np.array(len(sd)))


if __name__ == "__main__":
    main()
