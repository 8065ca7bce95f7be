"""CPU: the transformer-family oracle (oracle/transformer_oracle.py) against the golden fixtures produced by the
reference's own module objects (tests/golden/make_golden_tf.py), and the host-side contract of the mirrors (state-dict
keys, construction on CPU)."""
import os

import numpy as np
import pytest
import torch

from multimodal_scene_text_recognition_b200 import synth
from oracle import transformer_oracle as tfo


@pytest.fixture(scope="module")
def g_tf(golden_dir):
    return np.load(os.path.join(golden_dir, "cfg_tf.npz"))


@pytest.fixture(scope="module")
def sd_tf():
    return synth.make_state_dict_transformer(999)


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def test_fixture_records_why_the_encoder_is_driven_from_a_loop(g_tf):
    # the reference's Model.forward raises under the installed torch; the fixture says so instead of hiding it
    assert str(g_tf["whole_model"]).startswith("TypeError") and "is_causal" in str(g_tf["whole_model"])


def test_encoder_oracle_matches_reference_layers(sd_tf, g_tf):
    x = g_tf["visual"]
    one = tfo.encoder_layer(x + sd_tf["encoder.pos_encoder.pe"][:, 0][None, :x.shape[1]], sd_tf, "encoder.encoder.layers.0.")
    assert _rel(one, g_tf["enc_layer0"]) < 2e-5
    enc = tfo.tf_encoder(x, sd_tf)
    assert _rel(enc, g_tf["enc"]) < 5e-5
    enc64 = tfo.tf_encoder(x, sd_tf, dt=np.float64)
    assert _rel(enc64, g_tf["enc"]) < 5e-5


def test_decoder_oracle_matches_reference_forward(sd_tf, g_tf):
    enc = g_tf["enc"]
    lg = tfo.tf_decoder(enc, sd_tf)                                   # greedy, recomputing the prefix like the reference
    assert lg.shape == g_tf["logits"].shape == (enc.shape[0], 25, 97)
    assert np.array_equal(lg.argmax(2), g_tf["ids"])
    assert _rel(lg, g_tf["logits"]) < 1e-4
    tf = tfo.tf_decoder(enc, sd_tf, text=g_tf["text"])               # teacher forced
    assert tf.shape == g_tf["logits_tf"].shape == (enc.shape[0], 26, 97)
    assert _rel(tf, g_tf["logits_tf"]) < 1e-4
    assert len({tuple(r) for r in g_tf["ids"]}) == enc.shape[0], "degenerate decode: the parity would be vacuous"


def test_linear_decoder_oracle(g_tf):
    sd = synth.make_state_dict_transformer(999, decoder="Linear")
    assert _rel(tfo.linear_decoder(g_tf["enc"], sd), g_tf["logits_linear"]) < 2e-5


@pytest.mark.parametrize("decoder,count", [("Transformer", 462), ("Linear", 314)])
def test_mirror_state_dict_is_the_spec(decoder, count):
    """Model() under the Transformer configurations registers exactly the tensors of synth.transformer_spec (whose
    names and shapes were checked against the reference's state dict when the fixtures were made: 462 / 314 tensors,
    same order, same values for the same seed) and loads the synthetic state dict strictly."""
    from multimodal_scene_text_recognition_b200 import config, model as M
    saved = (config.ENCODER, config.DECODER)
    try:
        config.ENCODER, config.DECODER = "Transformer", decoder
        torch.manual_seed(config.RANDOM_SEED)
        m = M.Model()
        sd = synth.make_state_dict_transformer(999, decoder=decoder)
        assert len(m.state_dict()) == count == len(sd)
        M.load_numpy_state_dict(m, sd)
        # the six layers are deep copies of one initialised layer and `norm` IS layer_norm, like in the reference
        assert m.encoder.encoder.norm is m.encoder.layer_norm
        if decoder == "Transformer":
            with pytest.raises(NotImplementedError):
                config.MULTIHEAD_PRE_MEMORY = True
                M.Model()
    finally:
        config.MULTIHEAD_PRE_MEMORY = False
        config.ENCODER, config.DECODER = saved


def _oscar_mirror():
    """The mirror built from the two seeds of tests/golden/make_golden_oscar.py."""
    from multimodal_scene_text_recognition_b200.modules.encoders import Oscar_Bert
    torch.manual_seed(999)
    m = Oscar_Bert().eval()
    synth.perturb_parameters_(m, 1234)
    return m


def test_oscar_bert_mirror_and_oracle_match_reference(golden_dir):
    """Oscar_Bert: the mirror consumes the RNG like the reference's class (203 tensors, equal checksums), and the numpy
    restatement of hid_to_bert -> BertModel(inputs_embeds) -> bert_to_hid reproduces the reference's output."""
    g = np.load(os.path.join(golden_dir, "cfg_oscar.npz"))
    m = _oscar_mirror()
    sd = m.state_dict()
    assert len(sd) == int(g["n_tensors"]) == 203
    sums = {k: float(v.double().sum()) for k, v in sd.items() if v.is_floating_point()}
    assert sorted(sums) == list(g["keys"])
    np.testing.assert_allclose([sums[k] for k in g["keys"]], g["sums"], rtol=1e-9, atol=1e-9)
    npsd = {"encoder." + k: v.numpy() for k, v in sd.items()}
    y = tfo.oscar_bert(g["x"], npsd)
    assert _rel(y, g["y"]) < 5e-5


def test_fusion_oracle_matches_reference(golden_dir):
    """PRE_ENCODER_MLP + PRE_DECODER_MLP: the oracle's get_relevant_semantic / combine restatement, the encoder from the
    fused features and the greedy decoder with the fused memory against the reference's own modules."""
    g = np.load(os.path.join(golden_dir, "cfg_tf_fusion.npz"))
    sd = synth.make_state_dict_transformer(999, pre_encoder_mlp=True, pre_decoder_mlp=True)
    rel = tfo.get_relevant_semantic(g["visual"], g["semantics"], sd, "encoder.sem_relevence_mlp.")
    assert _rel(rel, g["relevant"]) < 2e-5
    fused = tfo.pre_mlp_fusion(g["visual"], g["semantics"], sd, "encoder.sem_relevence_mlp.", "encoder.combine_mlp.")
    assert _rel(fused, g["fused"]) < 2e-5
    enc = tfo.tf_encoder(g["visual"], sd, semantics=g["semantics"])
    assert _rel(enc, g["enc"]) < 1e-4
    lg = tfo.tf_decoder(g["enc"], sd, semantics=g["semantics"])
    assert np.array_equal(lg.argmax(2), g["ids"]) and _rel(lg, g["logits"]) < 1e-4


def test_fusion_mirror_state_dict():
    from multimodal_scene_text_recognition_b200 import config, model as M
    saved = (config.ENCODER, config.DECODER)
    try:
        config.ENCODER, config.DECODER, config.PRE_ENCODER_MLP, config.PRE_DECODER_MLP = "Transformer", "Transformer", True, True
        m = M.Model()
        sd = synth.make_state_dict_transformer(999, pre_encoder_mlp=True, pre_decoder_mlp=True)
        assert len(m.state_dict()) == len(sd) == 485          # the reference's count with both switches on
        M.load_numpy_state_dict(m, sd)
        config.PRE_ENCODER_MLP = config.PRE_DECODER_MLP = False
        config.CLS_DECODER_INIT = config.POST_DECODER_MLP = True
        m2 = M.Model()
        sd2 = synth.make_state_dict_transformer(999, cls_decoder_init=True, post_decoder_mlp=True)
        assert len(m2.state_dict()) == len(sd2) == 482
        M.load_numpy_state_dict(m2, sd2)
    finally:
        config.PRE_ENCODER_MLP = config.PRE_DECODER_MLP = config.POST_DECODER_MLP = config.CLS_DECODER_INIT = False
        config.ENCODER, config.DECODER = saved


def test_cls_and_post_decoder_switches_oracle(golden_dir, g_tf):
    """CLS_DECODER_INIT + POST_DECODER_MLP: oracle against the reference's unmodified TF_Decoder.forward."""
    g = np.load(os.path.join(golden_dir, "cfg_tf_fusion2.npz"))
    assert abs(float(g["cls_min"]) - 1) < 1e-6 and abs(float(g["cls_max"]) - 1) < 1e-6   # the reference's "semantic cls" is ones
    sd = synth.make_state_dict_transformer(999, cls_decoder_init=True, post_decoder_mlp=True)
    lg = tfo.tf_decoder(g_tf["enc"], sd, cls_init=True, post_semantics=g["semantics"])
    assert np.array_equal(lg.argmax(2), g["ids"]) and _rel(lg, g["logits"]) < 1e-4
    tf = tfo.tf_decoder(g_tf["enc"], sd, text=g_tf["text"], cls_init=True, post_semantics=g["semantics"])
    assert _rel(tf, g["logits_tf"]) < 1e-4
