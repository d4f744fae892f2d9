import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import mspi_oracle as orc
from mspi_b200.train_engine import TrainPlan
sd = orc.make_state_dict(3, "calibrated")
clips, aud = orc.make_inputs(2, 64, 64, 2023)
gt, _ = orc.make_gt(orc.forward(sd, clips, aud)[0])
plan = TrainPlan(sd, 2, 16, 64, 64)
plan.forward_backward(clips.cuda(), aud.cuda(), gt.cuda())
torch.cuda.synchronize()
ht = {k: v.cpu() for k, v in plan.head_tensors.items()}
ha = {k: v.cpu() for k, v in plan.head_acts.items()}
gr = lambda t: plan._gbufs[t.data_ptr()].cpu().view(t.shape)
for side, p, z, hp in (("vis", "pv", "za", "mlp_vis"), ("aud", "pa", "zv", "mlp_aud")):
    pt = ht[p].clone().requires_grad_(True)
    (-0.5 * F.cosine_similarity(pt, ht[z], dim=-1).mean()).backward()
    dp_gpu = gr(plan.head_tensors[p])
    print(side, "dp err", float((dp_gpu - pt.grad).norm() / pt.grad.norm()))
    W3 = sd[hp + ".3.weight"]
    h = ha[hp + ".1"]                      # relu(LN(lin0)) = input of linear 3
    dh_ref = pt.grad @ W3
    dh_gpu = gr(plan.head_acts[hp + ".1"])
    print(side, "dh err", float((dh_gpu - dh_ref).norm() / dh_ref.norm()), "|dh|", float(dh_ref.norm()))
    # LN+ReLU backward reference
    x0 = ha[hp + ".0"].clone().requires_grad_(True)
    y = F.relu(F.layer_norm(x0, (x0.shape[1],), sd[hp + ".1.weight"], sd[hp + ".1.bias"], 1e-5))
    print(side, "h fwd err", float((y.detach() - h).norm() / h.norm()), "frac>0", float((h > 0).float().mean()))
    y.backward(dh_gpu)
    dx_gpu = gr(plan.head_acts[hp + ".0"])
    print(side, "ln dx err", float((dx_gpu - x0.grad).norm() / x0.grad.norm()))
