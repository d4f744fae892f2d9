#!/usr/bin/env python
"""K optimisation steps on one fixed batch: CUDA training step (tf32 tensor cores, bf16 frozen encoders) against the fp32
oracle (train_grads + adamw_step on the CPU).  Prints both loss curves.

  python tools/train_trajectory.py [--h 128 --w 128 --batch 2 --steps 20 --lr 1e-4 --init calibrated --seed 3]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(h, w, batch, steps, lr, init, seed, verbose=True):
    from oracle import mspi_oracle as orc
    from tests.parity import build_product_model
    torch.set_num_threads(os.cpu_count() or 1)
    sd = orc.make_state_dict(seed, init)
    clips, aud = orc.make_inputs(batch, h, w, 2023)
    gt, _ = orc.make_gt(orc.forward(sd, clips, aud)[0])
    model = build_product_model(sd)
    model.train()
    cu = []
    cc, ac, gc = clips.cuda(), aud.cuda(), gt.cuda()
    for _ in range(steps):
        cu.append(model.train_step(cc, ac, gc, lr=lr).tolist())
    ref = []
    work = {k: v.clone() for k, v in sd.items()}
    keys = orc.trainable_keys(work)
    m = {k: torch.zeros_like(work[k]) for k in keys}
    v = {k: torch.zeros_like(work[k]) for k in keys}
    t0 = time.time()
    for step in range(1, steps + 1):
        r = orc.train_grads(work, clips, aud, gt)
        ref.append([float(r["loss"]), float(r["kl"]), float(r["cc"]), float(r["loss_va"])])
        for k in keys:
            work[k], m[k], v[k] = orc.adamw_step(work[k], r["grads"][k], m[k], v[k], step, lr)
        work.update(r["stats"])
    if verbose:
        print(f"oracle: {(time.time() - t0) / steps:.2f} s/step")
        for i, (a, b) in enumerate(zip(cu, ref)):
            print(f"step {i + 1:2d}  cuda loss {a[0]:+.5f} kld {a[1]:.5f} cc {a[2]:.5f} va {a[3]:+.5f} | oracle loss {b[0]:+.5f} kld {b[1]:.5f} "
                  f"cc {b[2]:.5f} va {b[3]:+.5f} | d {a[0] - b[0]:+.2e}")
    return cu, ref


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=128)
    ap.add_argument("--w", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--init", default="calibrated")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cu, ref = run(a.h, a.w, a.batch, a.steps, a.lr, a.init, a.seed)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"args": vars(a), "cuda": cu, "oracle": ref}, f)
