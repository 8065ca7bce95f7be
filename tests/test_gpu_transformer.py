"""GPU parity of the transformer family: the kernels of csrc/transformer.cu against the numpy oracle, and
TF_Encoder / TF_Decoder / Linear_Decoder / Model (Transformer configurations) against the golden fixtures produced by
the reference's own module objects (tests/golden/cfg_tf.npz).  Everything goes through the C ABI (ops -> ctypes)."""
import os

import numpy as np
import pytest
import torch

from multimodal_scene_text_recognition_b200 import config, model as M, ops, synth
from oracle import recognition_oracle as orc
from oracle import transformer_oracle as tfo
from gpu_util import DEV, assert_close, dev, err, host

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


@pytest.fixture(scope="module")
def g_tf(golden_dir):
    return np.load(os.path.join(golden_dir, "cfg_tf.npz"))


@pytest.fixture(scope="module")
def sd_tf():
    return synth.make_state_dict_transformer(999)


@pytest.fixture()
def tf_config():
    saved = (config.ENCODER, config.DECODER, config.PRECISION, config.MAX_TEXT_LENGTH)
    config.ENCODER, config.DECODER, config.MAX_TEXT_LENGTH = "Transformer", "Transformer", 25
    yield config
    config.ENCODER, config.DECODER, config.PRECISION, config.MAX_TEXT_LENGTH = saved


def _model(sd, precision, decoder="Transformer"):
    config.ENCODER, config.DECODER, config.PRECISION = "Transformer", decoder, precision
    m = M.Model().eval()
    M.load_numpy_state_dict(m, sd)
    return m.to(DEV)


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,D", [(7, 256), (1000, 512), (33, 1024)])
def test_layer_norm(rows, D):
    r = rng(1)
    x = (r.standard_normal((rows, D)) * 3 + 1).astype(np.float32)
    g, b = r.uniform(0.5, 1.5, D).astype(np.float32), r.normal(0, 0.2, D).astype(np.float32)
    ref = tfo.layer_norm(x.astype(np.float64), g.astype(np.float64), b.astype(np.float64))
    out = ops.layer_norm(dev(x), dev(g), dev(b))
    assert_close(host(out), ref, 2e-6, "layer_norm f32")
    xb = orc.round_bf16(x)
    refb = tfo.layer_norm(xb.astype(np.float64), g.astype(np.float64), b.astype(np.float64))
    outb = ops.layer_norm(dev(xb, BF), dev(g), dev(b))
    assert outb.dtype == BF
    assert_close(host(outb), refb, 5e-3, "layer_norm bf16 (one rounding of the result)")


@pytest.mark.parametrize("B,Lq,Lk,heads,dh,causal", [(3, 26, 26, 8, 64, False), (5, 26, 26, 8, 32, True),
                                                     (2, 26, 26, 8, 32, False), (4, 1, 26, 8, 32, False),
                                                     (2, 32, 32, 4, 64, True), (3, 7, 19, 8, 32, False)])
def test_mha_core_matches_oracle(B, Lq, Lk, heads, dh, causal):
    """softmax(q k^T / sqrt(dh)) v on strided views of packed projections, f32 and bf16."""
    r = rng(2)
    E = heads * dh
    if Lq == Lk:
        qkv = r.standard_normal((B, Lq, 3 * E)).astype(np.float32)
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
    else:
        q = r.standard_normal((B, Lq, E)).astype(np.float32)
        kv = r.standard_normal((B, Lk, 2 * E)).astype(np.float32)
        k, v = kv[..., :E], kv[..., E:]

    def ref_of(q, k, v):
        qh = q.astype(np.float64).reshape(B, Lq, heads, dh).transpose(0, 2, 1, 3) / np.sqrt(dh)
        kh = k.astype(np.float64).reshape(B, Lk, heads, dh).transpose(0, 2, 1, 3)
        vh = v.astype(np.float64).reshape(B, Lk, heads, dh).transpose(0, 2, 1, 3)
        s = qh @ kh.transpose(0, 1, 3, 2)
        if causal:
            s = s + tfo._causal_mask(Lq)
        return (tfo.softmax(s, -1) @ vh).transpose(0, 2, 1, 3).reshape(B, Lq, E)

    for dt, tol in ((torch.float32, 2e-6), (BF, 5e-3)):
        if Lq == Lk:
            t = dev(qkv, dt)
            tq, tk, tv = t[..., :E], t[..., E:2 * E], t[..., 2 * E:]
        else:
            tq, tkv = dev(q, dt), dev(kv, dt)
            tk, tv = tkv[..., :E], tkv[..., E:]
        out = ops.mha(tq, tk, tv, heads, causal=causal)
        ref = ref_of(host(tq), host(tk), host(tv))
        assert_close(host(out), ref, tol, "mha %s" % dt)


@pytest.mark.parametrize("dh", [32, 64])
def test_mha_step_with_cache_equals_full_causal_attention(dh):
    """Feeding the positions one at a time through mha_step (appending k / v to the cache) gives the rows of the
    causal whole-sequence attention - the K/V-cache equivalence the greedy loop relies on - bit for bit in f32."""
    r = rng(3)
    B, L, heads = 6, 25, 8
    E = heads * dh
    qkv = dev(r.standard_normal((B, L, 3 * E)).astype(np.float32))
    full = ops.mha(qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:], heads, causal=True)
    kc = torch.full((B, L, E), float("nan"), device=DEV)
    vc = torch.full((B, L, E), float("nan"), device=DEV)
    for t in range(L):
        row = qkv[:, t].contiguous()
        out = ops.mha_step(row[:, :E], kc, vc, t + 1, heads, k_new=row[:, E:2 * E], v_new=row[:, 2 * E:])
        assert_close(host(out), host(full[:, t]), 2e-6, "step %d" % t)
    assert torch.equal(kc, qkv[..., E:2 * E]) and torch.equal(vc, qkv[..., 2 * E:])


def test_embed_pos_argmax():
    r = rng(4)
    B, T, D, C = 5, 26, 256, 97
    emb = r.standard_normal((C, D)).astype(np.float32)
    pe = synth.positional_table(D)[:, 0]
    tok = r.integers(0, C, (B, T)).astype(np.int32)
    out = ops.embed_tokens(dev(tok), dev(emb), dev(pe), 0, torch.float32)
    assert np.array_equal(host(out), emb[tok] + pe[None, :T])
    one = ops.embed_tokens(dev(tok[:, 3:4]), dev(emb), dev(pe), 3, torch.float32)
    assert np.array_equal(host(one)[:, 0], emb[tok[:, 3]] + pe[3])
    x = r.standard_normal((B, T, D)).astype(np.float32)
    assert np.array_equal(host(ops.add_pos_encoding(dev(x), dev(pe), torch.float32)), x + pe[None, :T])
    assert np.array_equal(host(ops.add_pos_encoding(dev(x), dev(pe), BF)), orc.round_bf16(x + pe[None, :T]))
    # argmax: ties go to the lowest index, padded columns are ignored, strided destinations
    lg = r.standard_normal((B, 128)).astype(np.float32)
    lg[:, 100] = 50.0                      # beyond C: must be ignored
    lg[1, 7] = lg[1, 30] = 9.0             # tie
    logits = torch.zeros((B, 4, C), device=DEV)
    ids = torch.full((B, 4), -1, device=DEV, dtype=torch.int32)
    nxt = torch.zeros((B,), device=DEV, dtype=torch.int32)
    ops.argmax_tokens(dev(lg), C, logits[:, 2], ids[:, 2], nxt)
    assert np.array_equal(host(logits)[:, 2], lg[:, :C]) and not host(logits)[:, [0, 1, 3]].any()
    assert np.array_equal(host(ids)[:, 2], lg[:, :C].argmax(1)) and host(ids)[1, 2] == 7
    assert np.array_equal(host(nxt), lg[:, :C].argmax(1))


# ------------------------------------------------------------------------------------------------
# modules and model against the reference's outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol_enc,tol_logits", [("fp32", 1e-4, 3e-4), ("tf32", 1e-4, 3e-4)])
def test_encoder_and_decoder_match_reference_golden(tf_config, sd_tf, g_tf, precision, tol_enc, tol_logits):
    """TF_Encoder from the reference's visual features, TF_Decoder (greedy with the K/V cache, and teacher forced) from
    the reference's encoder output: ids identical, logits within the parity tolerance."""
    m = _model(sd_tf, precision)
    with torch.no_grad():
        enc = m.encoder.forward_features(dev(g_tf["visual"]))
        assert_close(host(enc), g_tf["enc"], tol_enc, "TF_Encoder output vs reference (%s)" % precision)
        out = m.decoder.decode(dev(g_tf["enc"]), None, False)
        assert np.array_equal(host(out["ids"]), g_tf["ids"]), "greedy ids differ from the reference"
        assert_close(host(out["logits"]), g_tf["logits"], tol_logits, "greedy logits vs reference")
        assert np.array_equal(host(out["lengths"]), np.full(g_tf["ids"].shape[0], 25, np.int32))
        tf = m.decoder(dev(g_tf["enc"]), dev(g_tf["text"]), None, True)
        assert_close(host(tf), g_tf["logits_tf"], tol_logits, "teacher-forced logits vs reference")
        # the reference-facing layer interface ([T, B, d] in and out) on one encoder layer
        x0 = dev(g_tf["visual"] + sd_tf["encoder.pos_encoder.pe"][:, 0][None]).permute(1, 0, 2)
        y0 = m.encoder.encoder.layers[0](x0)
        assert_close(host(y0.permute(1, 0, 2).float()), g_tf["enc_layer0"], tol_enc, "encoder layer 0 vs reference")


def test_whole_model_transformer_configuration_fp32(tf_config, sd_tf, g_tf):
    """Model.forward / recognize with ENCODER = DECODER = 'Transformer' from the crops: the reference's own Model.forward
    raises under torch 2.11 (see the fixture), its parts chained by the fixture give these ids."""
    m = _model(sd_tf, "fp32")
    B = g_tf["ids"].shape[0]
    x = synth.make_crops(123, B)
    ov, sc, io = synth.make_semantics(124, B)
    st = {}
    with torch.no_grad():
        out = m.recognize(dev(x), dev(ov), dev(sc), dev(io), st)
        logits = m(dev(x), dev(g_tf["text"]), dev(ov), dev(sc), dev(io), is_train=False)
    assert_close(host(st["visual"]), g_tf["visual"], 1e-4, "visual features")
    assert_close(host(st["enc1"]), g_tf["enc"], 2e-4, "encoder output from the crops")
    assert np.array_equal(host(out["ids"]), g_tf["ids"])
    assert_close(host(out["logits"]), g_tf["logits"], 5e-4, "logits from the crops")
    assert torch.equal(logits, out["logits"])


def test_linear_decoder_configuration(tf_config, g_tf):
    sd = synth.make_state_dict_transformer(999, decoder="Linear")
    for precision, tol in (("fp32", 1e-5), ("tf32", 1e-5), ("bf16", 1e-2)):
        m = _model(sd, precision, decoder="Linear")
        with torch.no_grad():
            out = m.decoder.decode(dev(g_tf["enc"]), None, False)
        assert_close(host(out["logits"]), g_tf["logits_linear"], tol, "Linear_Decoder (%s)" % precision)
        if precision != "bf16":
            assert np.array_equal(host(out["ids"]), g_tf["logits_linear"].argmax(2))


def test_bf16_mode_against_reference(tf_config, sd_tf, g_tf):
    """bf16 throughput mode (tcgen05 GEMMs, bf16 activations, fp32 LayerNorm / softmax arithmetic): teacher-forced
    logits (no feedback) within the north star's bf16 tolerance of the reference; greedy ids where the reference's
    top-1 / top-2 margin exceeds the logit error."""
    m = _model(sd_tf, "bf16")
    with torch.no_grad():
        enc = m.encoder.forward_features(dev(g_tf["visual"]))
        r_enc = assert_close(host(enc), g_tf["enc"], 6e-2, "TF_Encoder bf16 vs reference")
        tf = m.decoder(dev(g_tf["enc"]), dev(g_tf["text"]), None, True)
        r_tf = assert_close(host(tf), g_tf["logits_tf"], 2e-2, "teacher-forced bf16 logits vs reference")
        out = m.decoder.decode(dev(g_tf["enc"]), None, False)
    ids, ref = host(out["ids"]), g_tf["ids"]
    top2 = np.sort(g_tf["logits"], axis=2)[:, :, -2:]
    margin = top2[:, :, 1] - top2[:, :, 0]
    d = np.abs(host(out["logits"]) - g_tf["logits"]).max(axis=2)
    # compare up to the first step where the reference's margin is within twice the observed logit error
    agree = total = 0
    for b in range(ids.shape[0]):
        thin = np.nonzero(margin[b] <= 2 * np.maximum(d[b], 1e-3))[0]
        k = thin[0] if len(thin) else ids.shape[1]
        agree += int((ids[b, :k] == ref[b, :k]).sum()); total += int(k)
    print("\nbf16 transformer: encoder %.2e, teacher-forced logits %.2e, greedy ids equal on %.3f of all steps, %d/%d of the "
          "steps before the first thin margin" % (r_enc, r_tf, (ids == ref).mean(), agree, total))
    assert agree == total
    assert (ids == ref).mean() > 0.5


# ------------------------------------------------------------------------------------------------
# Oscar / BERT encoder
# ------------------------------------------------------------------------------------------------
def test_gelu_kernel():
    r = rng(6)
    x = (r.standard_normal((1000, 3072)) * 2).astype(np.float32)
    ref = tfo.gelu(x.astype(np.float64))
    assert_close(host(ops.gelu_(dev(x))), ref, 2e-6, "gelu f32")
    xb = orc.round_bf16(x)
    assert_close(host(ops.gelu_(dev(xb, BF))), tfo.gelu(xb.astype(np.float64)), 5e-3, "gelu bf16")


def test_oscar_bert_matches_reference_golden(tf_config, golden_dir):
    """Oscar_Bert (hid_to_bert -> transformers.BertModel through inputs_embeds -> bert_to_hid) rebuilt from the two
    seeds of the fixture, against the output of the reference's own class."""
    from multimodal_scene_text_recognition_b200.modules.encoders import Oscar_Bert
    g = np.load(os.path.join(golden_dir, "cfg_oscar.npz"))
    torch.manual_seed(999)
    m = Oscar_Bert().eval()
    synth.perturb_parameters_(m, 1234)
    m = m.to(DEV)
    for precision, tol in (("fp32", 1e-4), ("tf32", 1e-4), ("bf16", 6e-2)):
        config.PRECISION = precision
        with torch.no_grad():
            y = m.forward_features(dev(g["x"]))
            y2 = m(dev(g["x"]), None, False)
        r = assert_close(host(y), g["y"], tol, "Oscar_Bert (%s)" % precision)
        assert y2.dtype == torch.float32 and torch.equal(y2, y.float())
        print("\nOscar_Bert %s: output vs reference %.2e" % (precision, r))


def test_transformer_configuration_edge_cases(tf_config, sd_tf):
    """Batch sizes around the GEMM tile (1, 129), an empty batch, shard invariance, DataParallel-style prefix loading
    and the CUDA-graph replay of the whole greedy forward."""
    from multimodal_scene_text_recognition_b200.pipeline import GraphedRecognizer
    m = _model(sd_tf, "bf16")
    x = synth.make_crops(321, 129)
    ov, sc, io = synth.make_semantics(322, 129)
    with torch.no_grad():
        full = m.recognize(dev(x), dev(ov), dev(sc), dev(io))
        one = m.recognize(dev(x[:1]), dev(ov[:1]), dev(sc[:1]), dev(io[:1]))
        part = m.recognize(dev(x[64:]), dev(ov[64:]), dev(sc[64:]), dev(io[64:]))
        empty = m.recognize(dev(x[:0]), dev(ov[:0]), dev(sc[:0]), dev(io[:0]))
        graphed = GraphedRecognizer(m)
        g1 = {k: v.clone() for k, v in graphed(dev(x), dev(ov), dev(sc), dev(io)).items()}
        g2 = graphed(dev(x), dev(ov), dev(sc), dev(io))
    torch.cuda.synchronize()
    assert tuple(full["logits"].shape) == (129, 25, 97) and tuple(full["ids"].shape) == (129, 25)
    assert np.isfinite(host(full["logits"])).all()
    # a crop's result does not depend on the batch it arrived in (every kernel of the path is row-independent)
    assert torch.equal(one["ids"], full["ids"][:1]) and torch.equal(one["logits"], full["logits"][:1])
    assert torch.equal(part["ids"], full["ids"][64:]) and torch.equal(part["logits"], full["logits"][64:])
    assert tuple(empty["logits"].shape) == (0, 25, 97) and tuple(empty["ids"].shape) == (0, 25)
    for g in (g1, g2):
        assert torch.equal(g["ids"], full["ids"]) and torch.equal(g["logits"], full["logits"])
    # state dict with the DataParallel prefix (get_model, model.py:115) loads into the wrapped model
    wrapped = torch.nn.DataParallel(M.Model().eval(), device_ids=[0])
    wrapped.load_state_dict({"module." + k: torch.from_numpy(v) for k, v in sd_tf.items()}, strict=True)


# ------------------------------------------------------------------------------------------------
# semantic fusion: PRE_ENCODER_MLP + PRE_DECODER_MLP
# ------------------------------------------------------------------------------------------------
def test_relevant_semantic_kernels():
    """pair_add_relu + relevance_reduce against the direct evaluation (53 objects: two score slots per lane)."""
    r = rng(8)
    B, T, O, C, E = 3, 26, 53, 512, 256
    a = r.standard_normal((B, T, C)).astype(np.float32)
    s_ = r.standard_normal((B, O, C)).astype(np.float32)
    h = ops.pair_add_relu(dev(a), dev(s_))
    ref_h = np.maximum(a[:, :, None] + s_[:, None], 0)
    assert np.array_equal(host(h), ref_h)
    w3 = (r.standard_normal(C) / np.sqrt(C)).astype(np.float32)
    sem = r.standard_normal((B, O, E)).astype(np.float32)
    out = ops.relevance_reduce(h, dev(w3), 0.3, dev(sem))
    sc = ref_h.astype(np.float64) @ w3.astype(np.float64) + 0.3
    p = tfo.softmax(sc, axis=2)
    ref = (p[..., None] * sem[:, None].astype(np.float64)).sum(axis=2)
    assert_close(host(out), ref, 5e-6, "relevance_reduce f32")
    hb = ops.pair_add_relu(dev(a, BF), dev(s_, BF))
    outb = ops.relevance_reduce(hb, dev(w3), 0.3, dev(sem))
    assert_close(host(outb), ref, 3e-2, "relevance_reduce bf16")


def test_semantic_fusion_matches_reference_golden(tf_config, golden_dir):
    """Model with PRE_ENCODER_MLP + PRE_DECODER_MLP: the object-tag vectors reach the prediction.  Fused features, encoder
    output, greedy ids / logits against the reference's own modules; the whole model from the crops; and the ids must
    change when the semantic vectors do."""
    g = np.load(os.path.join(golden_dir, "cfg_tf_fusion.npz"))
    sd = synth.make_state_dict_transformer(999, pre_encoder_mlp=True, pre_decoder_mlp=True)
    try:
        config.PRE_ENCODER_MLP = config.PRE_DECODER_MLP = True
        for precision, tol in (("fp32", 3e-4), ("tf32", 3e-4)):
            m = _model(sd, precision)
            sem = dev(g["semantics"])
            with torch.no_grad():
                rel = m.encoder.get_relevant_semantic(dev(g["visual"]), sem)
                assert_close(host(rel), g["relevant"], 1e-4, "relevant semantics (%s)" % precision)
                enc = m.encoder.forward_features(dev(g["visual"]), sem)
                assert_close(host(enc), g["enc"], 2e-4, "encoder output with fusion (%s)" % precision)
                out = m.decoder.decode(dev(g["enc"]), None, False, semantics=sem)
                assert np.array_equal(host(out["ids"]), g["ids"])
                assert_close(host(out["logits"]), g["logits"], tol, "logits with fusion (%s)" % precision)
        B = g["ids"].shape[0]
        x = synth.make_crops(123, B)
        ov, sc, io = synth.make_semantics(124, B)
        with torch.no_grad():
            full = m.recognize(dev(x), dev(ov), dev(sc), dev(io))
            other = m.recognize(dev(x), dev(np.roll(ov, 1, axis=0)), dev(sc), dev(io))
        assert np.array_equal(host(full["ids"]), g["ids"]), "whole model from the crops"
        assert (host(other["ids"]) != g["ids"]).mean() > 0.2, "the prediction must depend on the object tags"
        mb = _model(sd, "bf16")
        with torch.no_grad():
            encb = mb.encoder.forward_features(dev(g["visual"]), dev(g["semantics"]))
        assert_close(host(encb), g["enc"], 8e-2, "encoder output with fusion (bf16)")
    finally:
        config.PRE_ENCODER_MLP = config.PRE_DECODER_MLP = False


def test_cls_and_post_decoder_switches_match_reference_golden(tf_config, golden_dir, g_tf):
    """CLS_DECODER_INIT (the first target embedding replaced by the reference's "semantic cls" vector, which is ones) and
    POST_DECODER_MLP (the logits refined with the semantic vectors mapped to class space; 97-wide layers on 128-wide
    GEMM tiles) against the reference's unmodified TF_Decoder.forward, greedy and teacher forced."""
    g = np.load(os.path.join(golden_dir, "cfg_tf_fusion2.npz"))
    sd = synth.make_state_dict_transformer(999, cls_decoder_init=True, post_decoder_mlp=True)
    try:
        config.CLS_DECODER_INIT = config.POST_DECODER_MLP = True
        for precision, tol in (("fp32", 3e-4), ("tf32", 3e-4)):
            m = _model(sd, precision)
            sem = dev(g["semantics"])
            with torch.no_grad():
                out = m.decoder.decode(dev(g_tf["enc"]), None, False, semantics=sem)
                tf = m.decoder(dev(g_tf["enc"]), dev(g_tf["text"]), sem, True)
            assert np.array_equal(host(out["ids"]), g["ids"]), precision
            assert_close(host(out["logits"]), g["logits"], tol, "logits with CLS + POST (%s)" % precision)
            assert_close(host(tf), g["logits_tf"], tol, "teacher-forced logits with CLS + POST (%s)" % precision)
        mb = _model(sd, "bf16")
        with torch.no_grad():
            tfb = mb.decoder(dev(g_tf["enc"]), dev(g_tf["text"]), dev(g["semantics"]), True)
        assert_close(host(tfb), g["logits_tf"], 3e-2, "teacher-forced logits with CLS + POST (bf16)")
    finally:
        config.CLS_DECODER_INIT = config.POST_DECODER_MLP = False
