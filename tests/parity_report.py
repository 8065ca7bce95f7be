"""Per-stage parity of the CUDA path against the oracle (fp32, and bf16 with / without the
quantisation-matched oracle) -> gpurun_out/parity_report.json.  Diagnostic, run on the GPU box."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # gpu_util
from multimodal_scene_text_recognition_b200 import config, synth  # noqa: E402
from oracle import recognition_oracle as orc  # noqa: E402
from gpu_util import dev, err, host, load_model, nchw  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = synth.make_state_dict(999)
x = synth.make_crops(4242, B)
ov, sc, io = synth.make_semantics(4243, B)
ref32 = orc.forward(sd, x, ov, sc, io, stages=True)
refq = orc.forward(sd, x, ov, sc, io, quant="bf16", stages=True)
ref64 = orc.forward(sd, x[:16], ov[:16], sc[:16], io[:16], dtype=np.float64, stages=True)
report = {"B": B}
print("fp32 oracle vs fp64 oracle (first 16): logits %.2e, ids equal %d/16" % (
    err(ref32["logits"][:16], ref64["logits"])[1], (ref32["ids"][:16] == ref64["ids"]).all(1).sum()))
for precision, grid in (("fp32", "reference"), ("tf32", "reference"), ("fp32", "folded"), ("bf16", "reference")):
    m = load_model(sd, precision=precision, TPS_GRID=grid, DECODER_EARLY_EXIT=False)
    st = {}
    with torch.no_grad():
        out = m.recognize(dev(x), dev(ov), dev(sc), dev(io), st)
    torch.cuda.synchronize()
    row = {}
    for name, ref in (("fp32_oracle", ref32), ("bf16_matched_oracle", refq)):
        ids = host(out["ids"])
        row[name] = {
            "rectified": err(host(st["rectified"]), ref["rectified"])[0],
            "features": err(nchw(host(st["features"])), ref["features"])[1],
            "enc1": err(host(st["enc1"]), ref["enc1"])[1],
            "logits": err(host(out["logits"]), ref["logits"])[1],
            "ids_equal_crops": int((ids == ref["ids"]).all(1).sum()),
            "ids_equal_steps": float((ids == ref["ids"]).mean()),
        }
    report["%s/%s" % (precision, grid)] = row
    print(precision, grid, json.dumps(row))

# The reference's OWN random init (torch.manual_seed(config.RANDOM_SEED); Model()): the mirror consumes the RNG
# identically, so the same seed rebuilds the same 253 tensors; oracle = the torch CPU port on those weights.
from multimodal_scene_text_recognition_b200 import model as M  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
config.MAX_TEXT_LENGTH = 25
torch.manual_seed(999)
cpu_model = M.Model().eval()
sd_def = {k: v.clone() for k, v in cpu_model.state_dict().items()}
ref_def = tp.forward(sd_def, torch.from_numpy(x), torch.from_numpy(ov), torch.from_numpy(sc), torch.from_numpy(io))
ref_logits = ref_def["logits"].numpy() if isinstance(ref_def, dict) else ref_def.numpy()
ref_ids = ref_logits.argmax(2)
row = {}
for precision in ("fp32", "tf32", "bf16"):
    config.PRECISION = precision
    config.TPS_GRID, config.DECODER_EARLY_EXIT, config.SEMANTIC_VECTOR = "reference", False, "overlap"
    torch.manual_seed(999)
    m = M.Model().eval().to("cuda:0")
    with torch.no_grad():
        out = m.recognize(dev(x), dev(ov), dev(sc), dev(io))
    torch.cuda.synchronize()
    ids = host(out["ids"])
    row[precision] = {"logits": err(host(out["logits"]), ref_logits)[1],
                      "ids_equal_crops": int((ids == ref_ids).all(1).sum()), "crops": B}
report["reference_default_init_seed999"] = row
print("default init", json.dumps(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
    json.dump(report, f, indent=1)
