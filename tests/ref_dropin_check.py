"""Drop-in check, run in its OWN process on the GPU box (tests/test_gpu_reference.py spawns it):

the reference's UNMODIFIED ``model.py`` and ``training_functions.py`` (baseline/_ref, sha256-verified) are
executed with ``config``, ``utils`` and ``modules.*`` resolving to this package's mirrors - the import change
INTEGRATION.md describes.  Then
  1. ``model.Model()`` (reference code: model.py:34-83) builds the mirrors; its ``forward`` (model.py:85-107: the
     NCHW feature tensor, ``permute`` + ``AdaptiveAvgPool2d`` + ``squeeze``, ``nn.Sequential`` encoder, decoder call
     with ``text=`` / ``semantics=`` keywords) must give the logits of this package's own ``Model`` bit for bit;
  2. ``model.get_model()`` (model.py:109-128) wraps it in ``DataParallel(device_ids=config.DEVICE_IDS)`` on
     ``config.PRIMARY_DEVICE``: state-dict keys carry the ``module.`` prefix and the forward runs through the wrapper;
  3. ``training_functions.validate`` (training_functions.py:150-207) runs unmodified over a stub dataloader whose
     ground truth equals the expected prediction for every second sample: the score must be exactly that share.
``coco_dataset`` (lmdb / shapely / COCO-Text API: the out-of-scope data side) is replaced by an empty stub module so
that ``training_functions`` imports.  Prints one JSON line.
"""
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

import numpy as np
import torch

import install_ref
import multimodal_scene_text_recognition_b200 as pkg  # noqa: F401
from multimodal_scene_text_recognition_b200 import _cabi, config as our_config, model as our_model, synth, utils as our_utils
from multimodal_scene_text_recognition_b200.modules import decoders, encoders, feature_extraction, semantic_vectors, transformation


def load_unmodified(name):
    install_ref.verify()
    path = os.path.join(install_ref.REF_DST, name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    assert torch.cuda.is_available(), "needs the GPU: the mirrors have no CPU path"
    # --- the import switch: the reference's top-level names resolve to the mirrors ---------------------------
    sys.modules["config"] = our_config
    sys.modules["utils"] = our_utils
    mods = types.ModuleType("modules")
    mods.__path__ = []
    sys.modules["modules"] = mods
    for name, m in (("transformation", transformation), ("feature_extraction", feature_extraction),
                    ("semantic_vectors", semantic_vectors), ("encoders", encoders), ("decoders", decoders)):
        sys.modules["modules." + name] = m
        setattr(mods, name, m)
    stub = types.ModuleType("coco_dataset")
    for fn in ("get_synth_datasets", "get_cocotext_datasets", "get_cocotext_single_image_data", "get_textocr_datasets"):
        setattr(stub, fn, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("data side is out of scope")))
    sys.modules["coco_dataset"] = stub

    our_config.DEVICE_IDS = [0]
    our_config.PRIMARY_DEVICE = "cuda:0"
    our_config.PRECISION = "bf16"
    our_config.DECODER_EARLY_EXIT = False
    our_config.BATCH_SIZE = 16
    ref_model = load_unmodified("model")
    tf = load_unmodified("training_functions")
    assert ref_model.Model is not our_model.Model and ref_model.__file__.startswith(install_ref.REF_DST)

    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(999).items()}
    B = 48
    x = torch.from_numpy(synth.make_crops(31, B)).cuda()
    ov, sc, io = [torch.from_numpy(a).cuda() for a in synth.make_semantics(32, B)]
    text = torch.zeros(B, 26, dtype=torch.long, device="cuda")
    launches0 = _cabi.launch_count()

    # 1. the reference's Model class over the mirrors vs this package's Model
    m_ref = ref_model.Model().eval()
    m_ref.load_state_dict(sd, strict=True)
    m_ref = m_ref.to("cuda:0")
    m_own = our_model.Model().eval()
    m_own.load_state_dict(sd, strict=True)
    m_own = m_own.to("cuda:0")
    with torch.no_grad():
        a = m_ref(x, text, ov, sc, io, is_train=False)
        b = m_own(x, text, ov, sc, io, is_train=False)
    same = bool(torch.equal(a, b))

    # 2. get_model(): DataParallel wrapper on config.PRIMARY_DEVICE
    gm = ref_model.get_model()
    keys = list(gm.state_dict().keys())
    prefixed = all(k.startswith("module.") for k in keys) and len(keys) == 253
    gm.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)
    gm.eval()
    with torch.no_grad():
        c = gm(x, text, ov, sc, io, is_train=False)
    same_dp = bool(torch.equal(c, b))

    # 3. validate() over a stub dataloader: ground truth = expected prediction for every second sample
    conv = our_utils.AttnLabelConverter(our_config.CHARS)
    pred = conv.decode_ids(b.argmax(2))
    truth = [p if i % 2 == 0 else p + "#" for i, p in enumerate(pred)]
    batches = []
    for lo in range(0, B, 16):
        hi = lo + 16
        batches.append((torch.arange(lo, hi), x[lo:hi].cpu(), truth[lo:hi], ov[lo:hi].cpu(), sc[lo:hi].cpu(), io[lo:hi].cpu()))
    score = tf.validate(gm, batches)
    expected = round(sum(1 for i in range(B) if i % 2 == 0) * 100 / B, 5)
    torch.cuda.synchronize()
    print(json.dumps({"forward_equals_mirror_model": same and same_dp, "get_model_keys_prefixed": prefixed,
                      "validate_score": score, "expected_score": expected, "distinct_predictions": len(set(pred)),
                      "native_kernel_launches": int(_cabi.launch_count() - launches0)}))


if __name__ == "__main__":
    main()
