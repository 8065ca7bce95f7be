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


class OpRecorder:
    """Record every call of the listed ``ops`` front ends (positional / keyword arguments and result) while a
    model runs: the per-layer parity tests re-evaluate each recorded launch with the oracle ON THE GPU PATH'S OWN
    INPUTS, so a layer is checked in isolation from the rounding history upstream of it."""
    NAMES = ("conv3x3_first", "conv3x3_flat", "conv3x3_flat_pool", "conv_gemm", "conv_gemm_dual", "maxpool_nhwc", "maxpool2x2_ring",
             "bilstm_recurrent_bf16", "attn_decode_bf16", "mean_h")

    def __init__(self, names=None):
        self.names = tuple(names or self.NAMES)
        self.calls = []

    def __enter__(self):
        from multimodal_scene_text_recognition_b200 import ops
        self._ops, self._saved = ops, {}
        for name in self.names:
            fn = getattr(ops, name)
            self._saved[name] = fn

            def wrapped(*a, _fn=fn, _name=name, **kw):
                out = _fn(*a, **kw)
                self.calls.append((_name, a, kw, out))
                return out
            setattr(ops, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(self._ops, name, fn)
        return False


def bf16_ulp(ref):
    """Size of one bf16 unit in the last place at |ref| (float64 array)."""
    a = np.abs(np.asarray(ref, np.float64))
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -126)))
    return 2.0 ** (e - 7)


def bf16_agreement(got, ref, floor_rel=2e-5):
    """Compare a bf16 result (as float) with an exactly evaluated reference (float64, NOT yet rounded).

    Returns (fraction bit-equal to round_bf16(ref), fraction within one bf16 ulp of ref, max |err| / max |ref|).
    "Within one ulp" allows ulp(|ref|) plus ``floor_rel * max|ref|``: the fp32 accumulation noise of a long dot
    product is absolute, so for results near zero (cancellation, ReLU boundary) it exceeds their tiny ulp."""
    from oracle import recognition_oracle as orc
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    exact = float((got == orc.round_bf16(ref.astype(np.float32)).astype(np.float64)).mean())
    m = max(float(np.abs(ref).max()), 1e-30)
    within = float((np.abs(got - ref) <= bf16_ulp(ref) + floor_rel * m).mean())
    return exact, within, float(np.abs(got - ref).max()) / m
