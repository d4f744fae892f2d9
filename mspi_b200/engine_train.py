"""train_one_epoch — the reference's engine_train.py:11-81 on the B200 training plan.

Same signature and return value (`{meter: global average}`), same per-iteration semantics: `model.train();
model.frozen_encoder()`, `output, loss_va = model(imgs, audio)`, `loss = criterion(output, label) + gamma*loss_va`, NaN check
(`Exception("Loss is NaN.")`), `optimizer.zero_grad(); loss.backward(); optimizer.step()` — except that forward, loss,
backward and the AdamW update are ONE call into the CUDA plan (model.train_step), so `optimizer` is only read for its
learning rate (`param_groups[*]['lr']`, which the reference's epoch schedule rewrites, train.py:160-166) and `criterion`
only receives the step's KLD / CC for its `.log` meters.  Under torch.distributed the flat gradient buffer is all-reduced
(sum) and scaled by 1/world_size inside the optimiser kernel, which is what DistributedDataParallel's gradient averaging does.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch


def train_one_epoch(model, criterion, data_loader: Iterable, optimizer, device, epoch: int, cfg, start_steps=None,
                    update_freq=1, gamma=1.0):
    model.train()
    model.frozen_encoder()
    world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
    from .distributed import allreduce_gradients
    allreduce = allreduce_gradients if world > 1 else None
    sums, n = {"loss": 0.0, "kld": 0.0, "cc": 0.0, "loss_va": 0.0}, 0
    lr = cfg.SOLVER.LR
    for n_iter, batch_data in enumerate(data_loader):
        if optimizer is not None:
            lr = max(g["lr"] for g in optimizer.param_groups)
        if cfg.DATA.USE_SOUND:                       # engine_train.py:30-38
            imgs, audio, label = batch_data
            audio = audio.to(device, non_blocking=True)
        else:                                        # engine_train.py:39-47 (VisualSaliencyModel)
            (imgs, label), audio = batch_data, None
        imgs = imgs.to(device, non_blocking=True)
        label = label.to(device, non_blocking=True)
        res = model.train_step(imgs, audio, label, lr=lr, gamma=gamma, allreduce=allreduce, world_size=world)
        loss_value, kld, cc, loss_va = res.tolist()          # the step's only device->host copy (4 floats)
        if math.isnan(loss_value):
            raise Exception("Loss is NaN.")
        if criterion is not None and hasattr(criterion, "log"):
            criterion.log["kl"].update(kld)
            criterion.log["cc"].update(cc)
            criterion.log["loss"].update(kld - cc)
        for k, v in zip(sums, (loss_value, kld, cc, loss_va)):
            sums[k] += v
        n += 1
    model.sync_from_training()
    out = {k: v / max(n, 1) for k, v in sums.items()}
    out["lr"] = lr
    return out
