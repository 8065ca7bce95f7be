"""Helpers shared by the GPU parity tests (tests/test_gpu_*.py)."""
import numpy as np
import torch

DEV = "cuda:0"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def host(t):
    return t.detach().float().cpu().numpy() if t.dtype == torch.bfloat16 else t.detach().cpu().numpy()


def nhwc(a):   # NCHW numpy -> NHWC numpy
    return np.ascontiguousarray(a.transpose(0, 2, 3, 1))


def nchw(a):   # NHWC numpy -> NCHW numpy
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2))


def err(a, b):
    """(max abs error, max abs error / max |b|)"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = float(np.abs(a - b).max()) if a.size else 0.0
    return d, d / max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)


def assert_close(a, b, rel, what):
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    assert np.isfinite(np.asarray(a, np.float64)).all(), "%s: non-finite values in the CUDA result" % what
    d, r = err(a, b)
    assert r <= rel, "%s: max abs err %.3e, rel-to-max %.3e > %.1e" % (what, d, r, rel)
    return r


def load_model(sd, geom=None, precision="fp32", **cfg):
    """Model on the GPU with the numpy state dict ``sd`` and config overrides applied."""
    from multimodal_scene_text_recognition_b200 import config, model as M, synth
    geom = geom or synth.CFG_A
    config.PRECISION = precision
    config.MAX_TEXT_LENGTH = geom.max_len
    for k, v in cfg.items():
        setattr(config, k, v)
    m = M.Model(imgH=geom.H, imgW=geom.W, input_channel=geom.in_ch).eval()
    M.load_numpy_state_dict(m, sd)
    return m.to(DEV)
