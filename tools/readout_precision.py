import sys, os
sys.path.insert(0, "/root/repo")
from tests.parity import run_forward_parity
for args in ((64, 96, 1, "default", 1), (224, 384, 1, "default", 1), (64, 64, 2, "calibrated", 0)):
    r = run_forward_parity(args[0], args[1], args[2], init=args[3], seed=args[4])
    rng = (r["ref_out"].max() - r["ref_out"].min()).item()
    print(os.environ.get("MSPI_READOUT1_BF16", "0"), args, "map", round(r["map_maxabs_minmax"], 5), "logit/rng", round(r["logit_maxabs"] / rng, 5), flush=True)
