"""SalLoss — same interface as utils/loss.py:6-49 of the reference: `SalLoss()(log_map, gt, fixations=None)`
returns the scalar loss (KLD - CC, minus 0.1*NSS when fixations are given) and updates `.log[...]` meters.
All reductions run in one CUDA kernel on the log map (exp fused); the meters read the five results with a
single device->host copy instead of the reference's 4-5 `.item()` syncs.

SalLoss here is a METRICS object: its forward runs under `torch.no_grad()` and the returned scalar carries no autograd graph,
so the reference's `criterion(output, label).backward()` has no counterpart — the loss and its gradient are computed inside
the training plan (`model.train_step`, mspi_salloss_bwd), which feeds this object's `.log` meters (engine_train.py)."""
import torch
import torch.nn as nn

from .compute_saliency_metrics import saliency_metrics


class AverageMeter:
    """timm.utils.AverageMeter (the reference's meter type)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class SalLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.reset_records()

    def reset_records(self):
        self.log = {k: AverageMeter() for k in ('kl', 'cc', 'sim', 'nss', 'loss')}

    @torch.no_grad()
    def forward(self, inputs, targets, fixations=None, targets2=None):
        res = saliency_metrics(inputs, targets, fixations, pred_is_log=True)
        kl, cc_, sim, nss_, loss = res.tolist()
        self.log['kl'].update(kl)
        self.log['cc'].update(cc_)
        self.log['sim'].update(sim)
        if fixations is not None:
            self.log['nss'].update(nss_)
        self.log['loss'].update(loss)
        return res[4]
