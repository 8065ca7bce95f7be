"""Whole-path parity: Model.forward on the B200 (through the nn.Module mirrors -> C-ABI kernels)
against the numpy oracle, the committed golden fixtures (outputs of the reference's own modules)
and size-independent properties at the benchmark batch size.  GPU only."""
import os

import numpy as np
import pytest
import torch

from multimodal_scene_text_recognition_b200 import config, synth
from oracle import recognition_oracle as orc
from gpu_util import DEV, assert_close, dev, err, host, load_model, nchw

pytestmark = pytest.mark.gpu


def _inputs(B, geom=synth.CFG_A, seed=123):
    x = synth.make_crops(seed, B, geom)
    ov, sc, io = synth.make_semantics(seed + 1, B)
    return x, ov, sc, io


def _run(model, x, ov, sc, io, stages=None):
    with torch.no_grad():
        out = model.recognize(dev(x), dev(ov), dev(sc), dev(io), stages)
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------------------------------------
# fp32 parity mode
# ------------------------------------------------------------------------------------------------
def test_fp32_matches_reference_golden(sd_cfg_a, golden_dir):
    """tests/golden/cfg_a.npz holds what the reference's own Model.forward produced for these
    weights and crops (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "cfg_a.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    x, ov, sc, io = _inputs(B, seed=dseed)
    m = load_model(sd_cfg_a, precision="fp32", TPS_GRID="reference", DECODER_EARLY_EXIT=False,
                   SEMANTIC_VECTOR="overlap")
    st = {}
    out = _run(m, x, ov, sc, io, st)
    d, _ = err(host(st["rectified"]), g["rectified"])
    assert d < 2e-4, "rectified image: max abs err %.3e" % d
    assert_close(nchw(host(st["features"])), g["features"], 1e-4, "ResNet features vs reference")
    assert_close(host(st["enc1"]), g["enc1"], 1e-4, "encoder output vs reference")
    assert np.array_equal(host(st["semantics"])[:, :, :16], g["semantics"])
    assert np.array_equal(host(out["ids"]), g["ids"]), "decoded ids differ from the reference"
    # TF32-mode tolerance of the north star is 1e-4 relative; two fp32 implementations of this
    # network differ by up to ~1.4e-4 (fp32 TPS-grid cancellation, see tests/test_oracle_golden.py)
    assert_close(host(out["logits"]), g["logits"], 3e-4, "logits vs reference")
    # the nn.Module forward returns the same logits
    with torch.no_grad():
        logits = m(dev(x), torch.zeros(B, 26, dtype=torch.long, device=DEV), dev(ov), dev(sc), dev(io), is_train=False)
    assert torch.equal(logits, out["logits"])
    # teacher-forced branch
    with torch.no_grad():
        tf = m(dev(x), dev(g["tf_text"]), dev(ov), dev(sc), dev(io), is_train=True)
    assert_close(host(tf), g["logits_tf"], 3e-4, "teacher-forced logits vs reference")


@pytest.mark.parametrize("precision,tol,tol_grid", [("fp32", 1e-6, 2e-5), ("tf32", 2e-6, 2e-5), ("bf16", 2e-3, 2e-3)])
def test_fiducial_points_match_reference_golden(sd_cfg_a, golden_dir, precision, tol, tol_grid):
    """Direct gate on LocalizationNetwork (transformation.py:74-82): the fiducial points C' [B, F, 2] the reference
    produced for these crops (non-zero fc2 weight in the recipe, so C' depends on the image), and the sampling grid
    built from them.  The head (pool + fc1 + fc2) is fp32 in every mode; bf16 enters through the three convolutions.
    Measured: C' 5.7e-8 (fp32), 2.0e-7 (3xTF32 convolutions), 2.7e-4 (bf16); the fp32 grid carries ~1e-5 of cancellation
    error in both implementations (tests/test_oracle_golden.py)."""
    g = np.load(os.path.join(golden_dir, "cfg_a.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    x, _, _, _ = _inputs(B, seed=dseed)
    m = load_model(sd_cfg_a, precision=precision, TPS_GRID="reference")
    tps = m.Transformation
    with torch.no_grad():
        c = tps.LocalizationNetwork(dev(x))
        grid = tps.GridGenerator.build_P_prime(c)
    torch.cuda.synchronize()
    assert c.dtype == torch.float32 and tuple(c.shape) == tuple(g["c_prime"].shape)
    r = assert_close(host(c), g["c_prime"], tol, "fiducial points vs reference (%s)" % precision)
    rg = assert_close(host(grid).reshape(B, -1, 2)[:, ::7], g["grid"], tol_grid, "sampling grid vs reference (%s)" % precision)
    print("\nC' vs reference golden (%s): %.2e, grid %.2e" % (precision, r, rg))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tf32", 1e-4), ("bf16", 2e-2)])
def test_reference_default_init_golden(golden_dir, precision, tol):
    """The reference's own random init (torch.manual_seed(config.RANDOM_SEED); Model()) rebuilt from the seed
    by the mirror's constructor, against what the reference's Model.forward produced for it
    (tests/golden/default_init.npz).  North-star tolerances: ids identical, logits 1e-4 (tf32) / 2e-2 (bf16)
    relative."""
    from multimodal_scene_text_recognition_b200 import model as M
    g = np.load(os.path.join(golden_dir, "default_init.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    config.PRECISION = precision
    config.MAX_TEXT_LENGTH = 25
    config.TPS_GRID, config.DECODER_EARLY_EXIT, config.SEMANTIC_VECTOR = "reference", False, "overlap"
    torch.manual_seed(wseed)
    m = M.Model().eval().to(DEV)
    x, ov, sc, io = _inputs(B, seed=dseed)
    st = {}
    out = _run(m, x, ov, sc, io, st)
    assert_close(host(st["enc1"]), g["enc1"], tol, "encoder output vs reference (default init)")
    assert_close(host(out["logits"]), g["logits"], tol, "logits vs reference (default init)")
    assert np.array_equal(host(out["ids"]), g["ids"]), "decoded ids differ from the reference (default init)"


def test_empty_batch_returns_empty_result(sd_cfg_a):
    m = load_model(sd_cfg_a, precision="bf16")
    x = torch.zeros((0, 1, 32, 100), device=DEV)
    z = torch.zeros((0, 15), device=DEV), torch.zeros((0, 52), device=DEV), torch.zeros((0, 52), device=DEV)
    with torch.no_grad():
        out = m.recognize(x, *z)
        logits = m(x, torch.zeros((0, 26), dtype=torch.long, device=DEV), *z, is_train=False)
    assert out["logits"].shape == (0, 26, 97) and out["ids"].shape == (0, 26) and out["lengths"].shape == (0,)
    assert logits.shape == (0, 26, 97)


def test_tf32_tensor_core_parity_mode(sd_cfg_a, golden_dir):
    """config.PRECISION = 'tf32': convs / linears on tcgen05.mma.kind::tf32 with the 3xTF32 split.  North-star
    tolerance for the TF32 mode: ids bit-identical, logits within 1e-4 relative (the fp32 noise floor of this
    network is ~2e-4, see test_fp32_matches_reference_golden)."""
    g = np.load(os.path.join(golden_dir, "cfg_a.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    x, ov, sc, io = _inputs(B, seed=dseed)
    m = load_model(sd_cfg_a, precision="tf32", TPS_GRID="reference", DECODER_EARLY_EXIT=False)
    st = {}
    out = _run(m, x, ov, sc, io, st)
    assert_close(nchw(host(st["features"])), g["features"], 1e-4, "tf32 ResNet features vs reference")
    assert_close(host(st["enc1"]), g["enc1"], 1e-4, "tf32 encoder output vs reference")
    assert np.array_equal(host(out["ids"]), g["ids"]), "decoded ids differ from the reference"
    assert_close(host(out["logits"]), g["logits"], 3e-4, "tf32 logits vs reference")
    # a larger batch against the oracle
    B2 = 24
    x, ov, sc, io = _inputs(B2, seed=321)
    ref = orc.forward(sd_cfg_a, x, ov, sc, io, stages=True)
    out = _run(m, x, ov, sc, io)
    assert np.array_equal(host(out["ids"]), ref["ids"])
    assert_close(host(out["logits"]), ref["logits"], 3e-4, "tf32 logits vs oracle")


def test_fp32_matches_oracle_batch(sd_cfg_a):
    B = 24
    x, ov, sc, io = _inputs(B, seed=321)
    ref = orc.forward(sd_cfg_a, x, ov, sc, io, stages=True)
    m = load_model(sd_cfg_a, precision="fp32", TPS_GRID="reference", DECODER_EARLY_EXIT=False)
    st = {}
    out = _run(m, x, ov, sc, io, st)
    assert_close(nchw(host(st["features"])), ref["features"], 1e-4, "features")
    assert_close(host(st["enc1"]), ref["enc1"], 1e-4, "enc1")
    ids = host(out["ids"])
    same = (ids == ref["ids"]).all(axis=1)
    assert same.all(), "ids differ on %d of %d crops" % ((~same).sum(), B)
    assert_close(host(out["logits"]), ref["logits"], 3e-4, "logits")
    assert len({tuple(r) for r in ids}) > B // 3, "degenerate decode: parity would be vacuous"
    # folded TPS constant (default) is more accurate than the reference order; same ids
    m2 = load_model(sd_cfg_a, precision="fp32", TPS_GRID="folded", DECODER_EARLY_EXIT=True)
    out2 = _run(m2, x, ov, sc, io)
    L = host(out2["lengths"])
    ids2 = host(out2["ids"])
    for b in range(B):
        k = min(int(L[b]) + 1, 26)
        assert np.array_equal(ids2[b, :k], ref["ids"][b, :k])


def test_scene_semantics_mode(sd_cfg_a):
    B = 6
    x, ov, sc, io = _inputs(B)
    m = load_model(sd_cfg_a, precision="fp32", SEMANTIC_VECTOR="scene", DECODER_EARLY_EXIT=False)
    st = {}
    out = _run(m, x, ov, sc, io, st)
    ref = orc.linear_embedding(ov, sc, io, sd_cfg_a["get_semantic_vectors.embed.weight"], "scene")
    assert_close(host(st["semantics"]), ref, 1e-6, "scene semantics")
    config.SEMANTIC_VECTOR = "overlap"
    out2 = _run(m, x, ov, sc, io)
    assert torch.equal(out["logits"], out2["logits"]), "semantics must not change the logits (decoders.py:79-101)"


def test_fp32_stress_config(sd_cfg_s, golden_dir):
    """3x64x256 crops, MAX_TEXT_LENGTH=50: T=65 encoder columns, 51 decode steps."""
    g = np.load(os.path.join(golden_dir, "cfg_s.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    geom = synth.CFG_S
    x = synth.make_crops(dseed, B, geom)
    ov, sc, io = synth.make_semantics(dseed + 1, B)
    m = load_model(sd_cfg_s, geom, precision="fp32", TPS_GRID="reference", DECODER_EARLY_EXIT=False)
    st = {}
    try:
        out = _run(m, x, ov, sc, io, st)
    finally:
        config.MAX_TEXT_LENGTH = 25
    assert tuple(st["features"].shape) == (B, 3, 65, 512)
    assert_close(nchw(host(st["features"])), g["features"], 3e-4, "features (stress)")
    assert_close(host(st["enc1"]), g["enc1"], 3e-4, "enc1 (stress)")
    assert host(out["logits"]).shape == (B, 51, 97)
    assert np.array_equal(host(out["ids"]), g["ids"])
    assert_close(host(out["logits"]), g["logits"], 1e-3, "logits (stress)")


# ------------------------------------------------------------------------------------------------
# bf16 throughput mode
# ------------------------------------------------------------------------------------------------
def test_bf16_matches_quantisation_matched_oracle(sd_cfg_a):
    """bf16 tensor-core mode, whole model free running from the crops, vs (a) the oracle with bf16 rounding inserted at
    the same points and (b) the plain fp32 oracle: a RECORD of how far two bf16 evaluations of this untrained network
    drift apart, with sanity gates only.  The tight bf16 gates live where rounding history cannot blur them:
    tests/test_gpu_bf16_layers.py - every conv / linear launch against the rounding-matched function of ITS OWN inputs
    (100 % of the elements within one bf16 ulp, >= 99 % bit-equal), the sequence model from the GPU's own features
    against the matched oracle (teacher-forced logits <= 3e-2, greedy ids >= 95 % of the steps) - and
    tests/test_gpu_reference.py (at least as close to the fp32 reference as the reference's own bf16 autocast run)."""
    B = 32
    x, ov, sc, io = _inputs(B, seed=555)
    ref = orc.forward(sd_cfg_a, x, ov, sc, io, quant="bf16", stages=True)
    ref32 = orc.forward(sd_cfg_a, x, ov, sc, io, stages=True)
    m = load_model(sd_cfg_a, precision="bf16", TPS_GRID="reference", DECODER_EARLY_EXIT=False)
    st = {}
    out = _run(m, x, ov, sc, io, st)
    feats = nchw(host(st["features"]))
    r_feat = assert_close(feats, ref["features"], 4e-2, "bf16 features vs matched oracle")
    r_feat32 = assert_close(feats, ref32["features"], 4e-2, "bf16 features vs fp32 oracle")
    r_enc = assert_close(host(st["enc1"]), ref["enc1"], 1e-1, "bf16 enc1 vs matched oracle")
    # the decoder alone, from the oracle's encoder output: isolates the sequence kernels from conv rounding
    ids = host(out["ids"])
    same = (ids == ref["ids"]).all(axis=1)
    same32 = (ids == ref32["ids"]).all(axis=1)
    r_log = err(host(out["logits"]), ref["logits"])[1]
    r_log32 = err(host(out["logits"]), ref32["logits"])[1]
    # teacher-forced (no greedy feedback): error of the logits themselves
    text = ref32["ids"].astype(np.int64)
    text = np.concatenate([np.zeros((B, 1), np.int64), text[:, :-1]], axis=1)
    with torch.no_grad():
        tf = m(dev(x), dev(text), dev(ov), dev(sc), dev(io), is_train=True)
    ref_tf = orc.forward(sd_cfg_a, x, ov, sc, io, text=text, is_train=True)
    r_tf = err(host(tf), ref_tf)[1]
    print("\nbf16 mode, B=%d: features %.2e (matched) %.2e (fp32) | enc1 %.2e | greedy logits %.2e (matched) "
          "%.2e (fp32) | teacher-forced logits vs fp32 %.2e | ids equal %d/%d (matched) %d/%d (fp32)" %
          (B, r_feat, r_feat32, r_enc, r_log, r_log32, r_tf, same.sum(), B, same32.sum(), B))
    steps32 = float((ids == ref32["ids"]).mean())
    assert steps32 >= 0.8, "greedy ids equal to the fp32 oracle on only %.3f of the steps" % steps32


def test_bf16_stress_config_runs_on_tensor_cores(sd_cfg_s, golden_dir):
    """3x64x256 crops / max length 50 in the bf16 mode (T=65 columns, 51 steps): every tensor-core kernel
    accepts the geometry; features stay within bf16 storage accuracy of the reference's."""
    g = np.load(os.path.join(golden_dir, "cfg_s.npz"))
    wseed, dseed, B = [int(v) for v in g["meta"]]
    geom = synth.CFG_S
    x = synth.make_crops(dseed, B, geom)
    ov, sc, io = synth.make_semantics(dseed + 1, B)
    m = load_model(sd_cfg_s, geom, precision="bf16", DECODER_EARLY_EXIT=False)
    st = {}
    try:
        out = _run(m, x, ov, sc, io, st)
    finally:
        config.MAX_TEXT_LENGTH = 25
    assert tuple(st["features"].shape) == (B, 3, 65, 512)
    assert_close(nchw(host(st["features"])), g["features"], 5e-2, "bf16 features (stress)")
    logits = host(out["logits"])
    assert logits.shape == (B, 51, 97) and np.isfinite(logits).all()
    assert (host(out["ids"]) == g["ids"]).mean() > 0.5


def test_bf16_properties_at_benchmark_batch(sd_cfg_a):
    """B=4096 (BASELINE.json config 3): results do not depend on how the batch is split or ordered,
    repeated runs are bit-identical, and the decode is not degenerate."""
    B = 4096
    x, ov, sc, io = _inputs(B, seed=2024)
    m = load_model(sd_cfg_a, precision="bf16", DECODER_EARLY_EXIT=False)
    full = _run(m, x, ov, sc, io)
    again = _run(m, x, ov, sc, io)
    assert torch.equal(full["logits"], again["logits"]) and torch.equal(full["ids"], again["ids"])
    # shards of 1/8 (what each rank sees at 8 GPUs) give the same rows
    lo, hi = 3 * 512, 4 * 512
    part = _run(m, x[lo:hi], ov[lo:hi], sc[lo:hi], io[lo:hi])
    assert torch.equal(part["ids"], full["ids"][lo:hi])
    assert torch.equal(part["logits"], full["logits"][lo:hi])
    # a permuted batch gives permuted results
    perm = np.random.Generator(np.random.PCG64(1)).permutation(256)
    sub = _run(m, x[perm], ov[perm], sc[perm], io[perm])
    assert torch.equal(sub["ids"], full["ids"][torch.from_numpy(perm).to(DEV)])
    ids = host(full["ids"])
    assert np.isfinite(host(full["logits"])).all()
    assert len({tuple(r) for r in ids}) > B // 8
    # early exit: identical up to and including [s], zero afterwards
    config.DECODER_EARLY_EXIT = True
    early = _run(m, x[:512], ov[:512], sc[:512], io[:512])
    config.DECODER_EARLY_EXIT = False
    L = host(early["lengths"])
    assert np.array_equal(L, host(full["lengths"])[:512])
    ei, fi = host(early["ids"]), ids[:512]
    for b in range(512):
        k = min(int(L[b]) + 1, 26)
        assert np.array_equal(ei[b, :k], fi[b, :k]) and not ei[b, k:].any()
    assert (L < 26).mean() > 0.02, "the synthetic weights should make some crops emit [s]"


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_ragged_batch_sizes_give_the_same_rows(sd_cfg_a, precision):
    """Batch sizes on both sides of every tile / wave boundary of the kernels (1 crop; 127 / 129 crops: the 128-row MMA
    tiles of the recurrences; 593: first size past the one-crop-per-four-warps attention; 801: past the BiLSTM cluster
    kernel's capacity; 2369: more crops than decoder warps): every crop's ids and logits are the ones it gets inside
    one 2369-crop batch, whatever batch it arrives in."""
    sizes = [1, 3, 127, 129, 593, 801]
    B = 2369
    x, ov, sc, io = _inputs(B, seed=77)
    m = load_model(sd_cfg_a, precision=precision, DECODER_EARLY_EXIT=False)
    full = _run(m, x, ov, sc, io)
    assert np.isfinite(host(full["logits"])).all()
    lo = 0
    for n in sizes:
        part = _run(m, x[lo:lo + n], ov[lo:lo + n], sc[lo:lo + n], io[lo:lo + n])
        assert torch.equal(part["ids"], full["ids"][lo:lo + n]), "ids of a %d-crop batch (%s)" % (n, precision)
        assert torch.equal(part["logits"], full["logits"][lo:lo + n]), "logits of a %d-crop batch (%s)" % (n, precision)
        assert torch.equal(part["lengths"], full["lengths"][lo:lo + n])
        lo += n
