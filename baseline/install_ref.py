"""Place the UNMODIFIED reference under ``baseline/_ref/`` (git-ignored, NOT gpurun-ignored) so that it
travels to the GPU box, where ``/root/reference`` does not exist.

    python baseline/install_ref.py

The reference is not a package: ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse
--target baseline/_ref /root/reference`` fails with "Neither 'setup.py' nor 'pyproject.toml' found" (outcome
recorded in DESIGN.md), so the files of the recognition path are copied byte for byte instead:
``model.py``, ``config.py``, ``utils.py``, ``modules/*.py`` and the caller ``training_functions.py`` (its
``validate`` loop drives the path in tests/ref_dropin_check.py).  Nothing is edited; a manifest with the
sha256 of every file is written next to them so that tests can check that what they import is what the
reference ships.  ``baseline/_ref`` never enters the git history (.gitignore) and nothing in the product
package imports it: it is used by ``bench.py --impl reference`` / ``cpu_baseline`` / ``gpu_eager_baseline``
and by the parity tests.

How the users of ``baseline/_ref`` import it (``load()`` below): the reference reads ``config`` at import and
at call time and hard-codes ``ENCODER = DECODER = 'Transformer'`` and ``DEVICE_IDS = [1]`` in config.py:4-28,
so the attributes of the imported ``config`` MODULE OBJECT are set before ``model`` is imported - the file
stays as shipped (the same recipe as SURVEY.md App. C).
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
FILES = ["model.py", "config.py", "utils.py", "training_functions.py", "modules/decoders.py", "modules/encoders.py",
         "modules/feature_extraction.py", "modules/semantic_vectors.py", "modules/transformation.py"]
MODULE_NAMES = ["config", "utils", "model", "modules", "modules.decoders", "modules.encoders",
                "modules.feature_extraction", "modules.semantic_vectors", "modules.transformation"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def install(src: str = REF_SRC, dst: str = REF_DST) -> bool:
    """Copy the reference's recognition-path files. Returns False (and leaves ``dst`` alone) when the
    reference tree is not present (the GPU box uses what the build container placed)."""
    if not os.path.isdir(src):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and _sha(d) == _sha(s)):
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
        manifest[rel] = _sha(d)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    return True


def available(dst: str = REF_DST) -> bool:
    return all(os.path.exists(os.path.join(dst, rel)) for rel in FILES)


def verify(dst: str = REF_DST) -> None:
    """Raise if a file under ``dst`` differs from the manifest written at install time."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        want = json.load(f)["sha256"]
    for rel, h in want.items():
        if _sha(os.path.join(dst, rel)) != h:
            raise RuntimeError("baseline/_ref/%s differs from the installed reference file" % rel)


def load(device: str = "cpu", device_ids=(0,), dst: str = REF_DST):
    """Import the unmodified reference ``model`` module with the LSTM/LSTM path selected.
    Returns (model_module, config_module).  ``config.PRIMARY_DEVICE`` is read at call time by
    decoders.py:36-38,54, so callers may change it between a CPU and a GPU run."""
    if not available(dst):
        raise FileNotFoundError("baseline/_ref is not installed: run `python baseline/install_ref.py` in the "
                                "build container (needs /root/reference)")
    verify(dst)
    for name in MODULE_NAMES:
        mod = sys.modules.get(name)
        if mod is None:
            continue
        # `modules` has no __init__.py in the reference: a namespace package, located by __path__ instead of __file__
        where = getattr(mod, "__file__", None) or next(iter(getattr(mod, "__path__", [])), "")
        if not os.path.abspath(where).startswith(dst):
            raise RuntimeError("module %r is already imported from %s" % (name, where))
    if dst not in sys.path:
        sys.path.insert(0, dst)
    cfg = importlib.import_module("config")
    cfg.ENCODER, cfg.DECODER = "LSTM", "LSTM"
    cfg.DEVICE_IDS = list(device_ids)
    cfg.PRIMARY_DEVICE = device
    model = importlib.import_module("model")
    return model, cfg


if __name__ == "__main__":
    ok = install()
    print("baseline/_ref: %s" % ("installed %d files from %s" % (len(FILES), REF_SRC) if ok else
                                 "reference tree %s not present" % REF_SRC))
    sys.exit(0 if ok or available() else 1)
