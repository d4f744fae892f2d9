"""Per-parameter gradient parity report of one training step (CUDA path vs oracle).  python tools/train_report.py [h w init]"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.parity import run_train_parity  # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = int(sys.argv[2]) if len(sys.argv) > 2 else 64
init = sys.argv[3] if len(sys.argv) > 3 else "calibrated"
try:
    r = run_train_parity(height=h, width=w, batch=2, init=init, seed=3, verbose=True)
    print({k: v for k, v in r.items() if k != "grad_errs"})
except Exception:
    traceback.print_exc()
    sys.exit(1)
