import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load_case, rel_err
from tests.test_parity_gpu import build_model
from idee_b200.models.losses import train_step_loss
from oracle import idee_oracle as O
cfg, sd, ins, train, ev, grads = load_case(sys.argv[1] if len(sys.argv) > 1 else "lfq_4_codes")
model = build_model(cfg, sd)
for which in ("loss_bce", "loss_var", "loss_anomaly", "TOTAL"):
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t, o = O.train_step_loss(sdg, ins["x"], ins["mask_extreme"], ins["mask_extreme_loss"], cfg)
    o["z_q"].retain_grad()
    (t if which == "TOTAL" else o[which]).sum().backward()
    model.zero_grad(set_to_none=True)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    out["z_q"].retain_grad()
    (total if which == "TOTAL" else out[which]).sum().backward()
    a, b = out["z_q"].grad.cpu(), o["z_q"].grad          # [N,V,C,T,H,W]
    per_v = [rel_err(a[:, v], b[:, v]) for v in range(a.shape[1])]
    per_c = [round(rel_err(a[:, :, c], b[:, :, c]), 3) for c in range(16)]
    per_t = [round(rel_err(a[:, :, :, t], b[:, :, :, t]), 3) for t in range(a.shape[3])]
    print(which, "z_q grad rel err per variable", per_v, "\n    per channel", per_c, "\n    per t", per_t)
