"""train_one_epoch / validation_one_epoch — the reference's engine_train.py:11-125 on the B200 plans.

Same signatures and return values (`{meter: global average}` with the reference's meter names: loss, kld, cc, sim, nss, lr,
min_lr, weight_decay, grad_norm), same per-iteration semantics: `model.train(); model.frozen_encoder()`,
`output, loss_va = model(imgs, audio)`, `loss = criterion(output, label) + gamma*loss_va`, NaN check
(`Exception("Loss is NaN.")`), `optimizer.zero_grad(); loss.backward(); optimizer.step()` — except that forward, loss,
backward and the AdamW update are ONE call into the CUDA plan (model.train_step).

What the torch optimizer passed in is used for:
  * its hyper-parameters — `param_groups[0]` lr (rewritten per epoch by the reference's schedule, train.py:160-166), betas,
    eps, weight_decay — are read every iteration and handed to the plan's AdamW kernel;
  * its STATE: moments found in it when the epoch starts (a resumed run: `optimizer.load_state_dict(ckpt['optimizer'])`,
    utils/optim.py:52-62) are loaded into the plan, and after the epoch the plan's moments and step count are written back
    with `optimizer.load_state_dict(model.optimizer_state_dict(...))`, so `optimizer.state_dict()` checkpoints the real AdamW
    state exactly as the reference's save_model does.
`criterion` (SalLoss) receives the step's KLD / CC / SIM for its `.log` meters; it is a metrics object here (its forward runs
under no_grad: there is no autograd graph to back-propagate through — the gradient comes from the plan's backward kernels).

Under torch.distributed (world_size > 1): parameters and BatchNorm buffers are broadcast from rank 0 before the first step
(what DistributedDataParallel does at construction), the flat gradient buffer is all-reduced (sum) each step and scaled by
1/world_size inside the optimiser kernel, and the returned meters are averaged over ranks
(MetricLogger.synchronize_between_processes, utils/log.py:36-47).
"""
from __future__ import annotations

import math
from typing import Iterable

import torch


def _world():
    d = torch.distributed
    return d.get_world_size() if d.is_available() and d.is_initialized() else 1


def _sync_meters(sums: dict, n: int, device) -> tuple:
    """Sum the meter totals and the sample count over ranks (utils/log.py:36-47)."""
    if _world() == 1:
        return sums, n
    keys = sorted(sums)
    t = torch.tensor([sums[k] for k in keys] + [float(n)], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t)
    vals = t.tolist()
    return dict(zip(keys, vals[:-1])), int(vals[-1])


def _adamw_hyper(optimizer, cfg):
    if optimizer is None or not optimizer.param_groups:
        return cfg.SOLVER.LR, (0.9, 0.999), 1e-8, 0.0, cfg.SOLVER.LR
    groups = optimizer.param_groups
    g0 = groups[0]
    for g in groups[1:]:   # the flat AdamW kernel applies one set of hyper-parameters to all trainable tensors
        same = all(g.get(k) == g0.get(k) for k in ("betas", "eps", "weight_decay", "lr"))
        if not same:
            raise ValueError("mspi_b200's flat AdamW step needs identical hyper-parameters in every param group "
                             "(the reference builds one group, train.py:157-158)")
    return (max(g["lr"] for g in groups), tuple(g0.get("betas", (0.9, 0.999))), g0.get("eps", 1e-8),
            g0.get("weight_decay", 0.0), min(g["lr"] for g in groups))


def train_one_epoch(model, criterion, data_loader: Iterable, optimizer, device, epoch: int, cfg, start_steps=None,
                    update_freq=1, gamma=1.0):
    from .distributed import allreduce_gradients, broadcast_training_state
    from .utils.compute_saliency_metrics import saliency_metrics
    model.train()
    model.frozen_encoder()
    world = _world()
    allreduce = allreduce_gradients if world > 1 else None
    sums = {k: 0.0 for k in ("loss", "kld", "cc", "sim", "nss", "grad_norm")}
    n = 0
    lr, betas, eps, wd, min_lr = _adamw_hyper(optimizer, cfg)
    state = model.training_state(torch.device(device))
    if world > 1 and not getattr(state, "_broadcast_done", False):
        broadcast_training_state(state)          # every rank starts from rank 0's parameters and buffers
        state._broadcast_done = True
    if optimizer is not None and state.step_count == 0 and len(getattr(optimizer, "state", {})) > 0:
        state.load_optimizer_state_dict(optimizer.state_dict())   # resumed run: continue the moments / bias correction
    for n_iter, batch_data in enumerate(data_loader):
        lr, betas, eps, wd, min_lr = _adamw_hyper(optimizer, cfg)
        if cfg.DATA.USE_SOUND:                       # engine_train.py:30-38
            imgs, audio, label = batch_data
            audio = audio.to(device, non_blocking=True)
        else:                                        # engine_train.py:39-47 (VisualSaliencyModel)
            (imgs, label), audio = batch_data, None
        imgs = imgs.to(device, non_blocking=True)
        label = label.to(device, non_blocking=True)
        res = model.train_step(imgs, audio, label, lr=lr, gamma=gamma, allreduce=allreduce, world_size=world, betas=betas,
                               eps=eps, weight_decay=wd)
        plan = model._last_train_plan
        extra = saliency_metrics(plan.out, plan.gt, None, pred_is_log=True)     # SIM of the train-mode map (criterion.log['sim'])
        gnorm = torch.linalg.vector_norm(plan.flat_g) / world                     # this step's gradient norm (a logged meter)
        loss_value, kld, cc, loss_va = res.tolist()          # the step's device->host copies: 4 + 5 + 1 floats
        sim = float(extra[2])
        if math.isnan(loss_value):
            raise Exception("Loss is NaN.")
        if criterion is not None and hasattr(criterion, "log"):
            criterion.log["kl"].update(kld)
            criterion.log["cc"].update(cc)
            criterion.log["sim"].update(sim)
            criterion.log["loss"].update(kld - cc)
        for k, v in zip(("loss", "kld", "cc", "sim", "nss", "grad_norm"), (loss_value, kld, cc, sim, 0.0, float(gnorm))):
            sums[k] += v      # nss: the trainer never passes fixations, the reference logs its initial 0 (engine_train.py:56)
        n += 1
    model.sync_from_training()
    if optimizer is not None and hasattr(optimizer, "load_state_dict"):
        osd = model.optimizer_state_dict(lr, betas, eps, wd)
        if osd["param_groups"] and sum(len(g["params"]) for g in optimizer.param_groups) == len(osd["param_groups"][0]["params"]):
            cur = optimizer.state_dict()
            osd["param_groups"] = cur["param_groups"]        # keep the caller's groups (schedulers edit them), replace the state
            optimizer.load_state_dict(osd)
        elif osd["param_groups"]:
            print("mspi_b200.train_one_epoch: the optimizer does not hold exactly the model's trainable tensors "
                  f"({sum(len(g['params']) for g in optimizer.param_groups)} vs {len(osd['param_groups'][0]['params'])}); its state was "
                  "not updated — checkpoint model.optimizer_state_dict() instead")
    sums, n = _sync_meters(sums, n, device)
    out = {k: v / max(n, 1) for k, v in sums.items()}
    out.update({"lr": lr, "min_lr": min_lr, "weight_decay": wd if wd > 0 else None})
    return out


@torch.no_grad()
def validation_one_epoch(model, data_loader: Iterable, device, cfg):
    """engine_train.py:84-125: eval-mode forward + SalLoss meters; returns {loss, kld, cc, sim} global averages."""
    from .utils.loss import SalLoss
    criterion = SalLoss()
    model.eval()
    sums, n = {"loss": 0.0, "kld": 0.0, "cc": 0.0, "sim": 0.0}, 0
    for batch_data in data_loader:
        if cfg.DATA.USE_SOUND:
            imgs, audio, label = batch_data
            output, _ = model(imgs.to(device, non_blocking=True), audio.to(device, non_blocking=True))
        else:
            imgs, label = batch_data
            output, _ = model(imgs.to(device, non_blocking=True))
        loss = criterion(output, label.to(device, non_blocking=True))
        sums["loss"] += float(loss)
        sums["kld"] += criterion.log["kl"].val
        sums["cc"] += criterion.log["cc"].val
        sums["sim"] += criterion.log["sim"].val
        n += 1
    sums, n = _sync_meters(sums, n, device)
    out = {k: v / max(n, 1) for k, v in sums.items()}
    print('* Kldiv {:.3f} CC {:.3f} SIM {:.3f} loss {:.3f}'.format(out["kld"], out["cc"], out["sim"], out["loss"]))
    return out
