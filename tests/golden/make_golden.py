"""Generate the golden fixtures that pin ``oracle/recognition_oracle.py``.

Run in the BUILD container only (it imports the unmodified reference from /root/reference,
which does not exist on the GPU box):

    python tests/golden/make_golden.py

For each case the reference's own modules (model.py, modules/*.py) are instantiated, loaded with
the seeded synthetic state dict of ``synth.make_state_dict`` and run on seeded synthetic crops in
fp32 on CPU, exactly as the oracle-import recipe of SURVEY.md App. C.  Stage tensors are stored
in small ``.npz`` files; weights and inputs are NOT stored - both are regenerated from seeds.

Cases
  cfg_a      Model.forward, 1x32x100, B=4, greedy (is_train=False), overlap semantics
  cfg_a_tf   same weights, teacher-forced branch (is_train=True) with a seeded text
  scene      Linear_Embedding in 'scene' mode
  cfg_s      3x64x256 crops, MAX_TEXT_LENGTH=50, B=1 (modules built directly, model.py hard-codes cfg-A)
  default_init  the reference's OWN random init: torch.manual_seed(config.RANDOM_SEED) then Model() - no
             synthetic recipe.  The mirror's constructor consumes the RNG identically, so the GPU test
             rebuilds the same 253 tensors from the seed alone (tests/test_gpu_model.py).  This
             network is degenerate (zero fc2.weight, 0/1 BatchNorm statistics: every crop yields the same
             logits), which is why the other cases use synth.make_state_dict - it is kept because it is the
             literal "random-init weights" configuration of the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import config  # noqa: E402  (reference config; must be patched before importing model)

config.ENCODER = "LSTM"
config.DECODER = "LSTM"
config.PRIMARY_DEVICE = "cpu"
config.DEVICE_IDS = [0]

import torch  # noqa: E402
import model as ref_model  # noqa: E402
from modules.transformation import TPS_SpatialTransformerNetwork  # noqa: E402
from modules.feature_extraction import ResNet_FeatureExtractor  # noqa: E402
from modules.encoders import BidirectionalLSTM  # noqa: E402
from modules.decoders import Attention  # noqa: E402
from modules.semantic_vectors import Linear_Embedding  # noqa: E402

from multimodal_scene_text_recognition_b200 import synth  # noqa: E402

WEIGHT_SEED = 999   # config.RANDOM_SEED
DATA_SEED = 123


def to_torch(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def case_cfg_a():
    torch.set_num_threads(8)
    geom = synth.CFG_A
    sd = synth.make_state_dict(WEIGHT_SEED, geom)
    m = ref_model.Model().eval()
    m.load_state_dict(to_torch(sd), strict=True)
    B = 4
    x = torch.from_numpy(synth.make_crops(DATA_SEED, B, geom))
    ov, sc, io = [torch.from_numpy(a) for a in synth.make_semantics(DATA_SEED + 1, B)]
    text = torch.zeros(B, geom.steps, dtype=torch.long)
    with torch.no_grad():
        c_prime = m.Transformation.LocalizationNetwork(x)
        grid = m.Transformation.GridGenerator.build_P_prime(c_prime)
        rect = m.Transformation(x)
        feat = m.FeatureExtraction(rect)
        vis = m.AdaptiveAvgPool(feat.permute(0, 3, 1, 2)).squeeze(3)
        sem = m.get_semantic_vectors(ov, sc, io)
        enc0 = m.encoder[0](vis)
        enc1 = m.encoder[1](enc0)
        logits = m(x, text, ov, sc, io, is_train=False)
        # teacher-forced branch with a seeded text
        rng = np.random.Generator(np.random.PCG64(DATA_SEED + 2))
        tf_text = torch.from_numpy(rng.integers(0, geom.num_classes, (B, geom.steps)).astype(np.int64))
        tf_text[:, 0] = 0
        logits_tf = m(x, tf_text, ov, sc, io, is_train=True)
        config.SEMANTIC_VECTOR = "scene"
        sem_scene = m.get_semantic_vectors(ov, sc, io)
        io2 = torch.from_numpy(np.random.Generator(np.random.PCG64(5)).uniform(0, 1, io.shape).astype(np.float32))
        sem_scene2 = m.get_semantic_vectors(ov, sc, io2)
        config.SEMANTIC_VECTOR = "overlap"
    np.savez_compressed(
        os.path.join(HERE, "cfg_a.npz"),
        c_prime=c_prime.numpy(), grid=grid.numpy()[:, ::7], rectified=rect.numpy(),
        features=feat.numpy(), visual=vis.numpy(), semantics=sem.numpy()[:, :, :16],
        enc0=enc0.numpy(), enc1=enc1.numpy(), logits=logits.numpy(),
        ids=logits.argmax(2).numpy(), tf_text=tf_text.numpy(), logits_tf=logits_tf.numpy(),
        sem_scene=sem_scene.numpy()[:, :, :16], sem_scene_rand_ious=sem_scene2.numpy()[:, :, :16],
        rand_ious=io2.numpy(),
        meta=np.array([WEIGHT_SEED, DATA_SEED, B]),
    )
    print("cfg_a ids:", logits.argmax(2)[:, :12].tolist())


def case_cfg_s():
    geom = synth.CFG_S
    sd = synth.make_state_dict(WEIGHT_SEED, geom)
    tsd = to_torch(sd)
    config.MAX_TEXT_LENGTH = geom.max_len   # decoders.py:34 reads it at call time
    tps = TPS_SpatialTransformerNetwork(F=geom.F, I_size=(geom.H, geom.W), I_r_size=(geom.H, geom.W),
                                        I_channel_num=geom.in_ch).eval()
    fe = ResNet_FeatureExtractor(geom.in_ch, geom.out_ch).eval()
    enc = torch.nn.Sequential(BidirectionalLSTM(geom.out_ch, geom.hidden, geom.hidden),
                              BidirectionalLSTM(geom.hidden, geom.hidden, geom.hidden)).eval()
    dec = Attention(geom.hidden, geom.hidden, geom.num_classes).eval()

    def load(mod, prefix):
        mod.load_state_dict({k[len(prefix):]: v for k, v in tsd.items() if k.startswith(prefix)}, strict=True)

    load(tps, "Transformation."); load(fe, "FeatureExtraction."); load(enc, "encoder."); load(dec, "decoder.")
    B = 1
    x = torch.from_numpy(synth.make_crops(DATA_SEED, B, geom))
    with torch.no_grad():
        rect = tps(x)
        feat = fe(rect)
        vis = torch.nn.AdaptiveAvgPool2d((None, 1))(feat.permute(0, 3, 1, 2)).squeeze(3)
        enc1 = enc(vis)
        logits = dec(enc1.contiguous(), text=None, semantics=None, is_train=False)
    config.MAX_TEXT_LENGTH = 25
    np.savez_compressed(
        os.path.join(HERE, "cfg_s.npz"),
        rectified=rect.numpy()[:, :, ::4, ::4], features=feat.numpy(), visual=vis.numpy(),
        enc1=enc1.numpy(), logits=logits.numpy(), ids=logits.argmax(2).numpy(),
        meta=np.array([WEIGHT_SEED, DATA_SEED, B]),
    )
    print("cfg_s feat", tuple(feat.shape), "ids:", logits.argmax(2)[:, :12].tolist())


def case_default_init():
    torch.manual_seed(WEIGHT_SEED)            # config.RANDOM_SEED (config.py:16)
    m = ref_model.Model().eval()
    B = 4
    x = torch.from_numpy(synth.make_crops(DATA_SEED, B))
    ov, sc, io = [torch.from_numpy(a) for a in synth.make_semantics(DATA_SEED + 1, B)]
    text = torch.zeros(B, 26, dtype=torch.long)
    with torch.no_grad():
        rect = m.Transformation(x)
        feat = m.FeatureExtraction(rect)
        enc1 = m.encoder(m.AdaptiveAvgPool(feat.permute(0, 3, 1, 2)).squeeze(3))
        logits = m(x, text, ov, sc, io, is_train=False)
    # a checksum of the freshly initialised parameters pins "same weights" without storing 200 MB
    sd = m.state_dict()
    names = sorted(k for k in sd if sd[k].dtype.is_floating_point)
    sums = np.array([float(sd[k].double().sum()) for k in names])
    abss = np.array([float(sd[k].double().abs().sum()) for k in names])
    np.savez_compressed(
        os.path.join(HERE, "default_init.npz"),
        rectified=rect.numpy()[:, :, ::4, ::4], enc1=enc1.numpy(), logits=logits.numpy(),
        ids=logits.argmax(2).numpy(), param_sums=sums, param_abs_sums=abss,
        meta=np.array([WEIGHT_SEED, DATA_SEED, B]),
    )
    print("default_init ids:", logits.argmax(2)[:, :12].tolist())


if __name__ == "__main__":
    case_cfg_a()
    case_cfg_s()
    case_default_init()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
