"""Print the parity report (CUDA path vs CPU oracle) for a list of configurations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.parity import run_forward_parity

cfgs = [("default", 64, 96, 1, 1), ("calibrated", 64, 64, 2, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "full":
    cfgs += [("default", 224, 384, 1, 1), ("calibrated", 224, 384, 1, 0)]
for init, h, w, b, seed in cfgs:
    r = run_forward_parity(h, w, b, init=init, seed=seed, verbose=False)
    ref, out = r.pop("ref_out"), r.pop("out")
    rng = (ref.max() - ref.min()).item()
    print(json.dumps({"init": init, "hw": [h, w], "b": b, "logit_range": rng, "logit_err_over_range": r["logit_maxabs"] / rng,
                      **{k: v for k, v in r.items() if k != "taps"}, "taps_max": max(r["taps"].values())}))
