import os, sys, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load_case, rel_err
from tests.test_parity_gpu import build_model
from idee_b200.models.losses import train_step_loss
from oracle import idee_oracle as O
cfg, sd, ins, train, ev, grads = load_case("lfq_16_codes")
ocfg = dataclasses.replace(cfg, exact_ste=True)
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
wt, wo = O.train_step_loss(sdg, ins["x"], ins["mask_extreme"], ins["mask_extreme_loss"], ocfg)
wo["z_q"].retain_grad(); wo["z_enc"].retain_grad()
wt.backward()
model = build_model(cfg, sd)
total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
out["z_q"].retain_grad()
total.backward()
named = dict(model.named_parameters())
errs = sorted(((rel_err(named[k].grad, sdg[k].grad), k) for k in sd), reverse=True)
for e, k in errs[:10]: print(f"{e:.3e} {k}  |oracle| {float(sdg[k].grad.abs().max()):.3e}  |golden(ref)| {float(grads[k].abs().max()):.3e} ref-vs-oracle {rel_err(grads[k], sdg[k].grad):.3e}")
print("z_q grad err", rel_err(out["z_q"].grad, wo["z_q"].grad), "idx equal", bool((out["anomaly"].cpu() == wo["anomaly"]).all()))
d = (out["z_q"].grad.cpu() - wo["z_q"].grad).abs()
print("z_q grad: #elements off", int((d > 1e-6 * wo["z_q"].grad.abs().max()).sum()), "of", d.numel())
