"""Gradient parity report of one training step (CUDA path vs oracle).  python tools/train_report.py [h w init] [--whole]"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.parity import run_train_parity, run_train_segment_parity  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
h = int(args[0]) if len(args) > 0 else 64
w = int(args[1]) if len(args) > 1 else 64
init = args[2] if len(args) > 2 else "calibrated"
try:
    if "--whole" in sys.argv:
        r = run_train_parity(height=h, width=w, batch=2, init=init, seed=3, verbose=True)
        print({k: v for k, v in r.items() if k != "grad_errs"})
    else:
        r = run_train_segment_parity(height=h, width=w, batch=2, init=init, seed=3, verbose=True)
        print({k: v for k, v in r.items() if k not in ("decoder_errs",)})
except Exception:
    traceback.print_exc()
    sys.exit(1)
