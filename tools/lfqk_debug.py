import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load_case, rel_err
from tests.test_parity_gpu import build_model
from idee_b200.models.losses import train_step_loss
from oracle import idee_oracle as O
cfg, sd, ins, train, ev, grads = load_case("lfq_4_codes")
model = build_model(cfg, sd)
total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
total.backward()
named = dict(model.named_parameters())
errs = sorted(((rel_err(named[k].grad, g), k) for k, g in grads.items()), reverse=True)
for e, k in errs[:8]: print(f"{e:.3e} {k}")
print("...")
for e, k in errs:
    if k.startswith("vq.") or k.startswith("cls."): print(f"{e:.3e} {k}")
print("loss parts", {k: (float(out[k]), float(train[k])) for k in ("loss_bce", "loss_anomaly", "loss_var", "loss_z_q")})
# gradient w.r.t. z_q from each loss separately
import torch
for which in ("loss_bce", "loss_anomaly", "loss_var"):
    model.zero_grad(set_to_none=True)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    out[which].backward()
    print(which, "d project_out.bias", named["vq.project_out.bias"].grad[:4].tolist())


from idee_b200 import ops
orig = ops.LFQGeneralFn.backward
def dbg(ctx, gzq, gidx, gaux):
    print("   LFQ bwd: gzq", None if gzq is None else (tuple(gzq.shape), gzq.stride(), float(gzq.abs().sum()), float(gzq.sum())), "gaux", None if gaux is None else float(gaux))
    r = orig(ctx, gzq, gidx, gaux)
    print("   -> g_b_out sum", float(r[4].sum()), "gz abs sum", float(r[0].abs().sum()))
    return r
ops.LFQGeneralFn.backward = staticmethod(dbg)
for combo in (("loss_anomaly",), ("loss_anomaly", "loss_z_q"), ("loss_bce", "loss_anomaly"), ("TOTAL",)):
    model.zero_grad(set_to_none=True)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    print(combo)
    (total if combo == ("TOTAL",) else sum(out[c].sum() for c in combo)).backward()
    print("   project_out.bias grad", named["vq.project_out.bias"].grad[:4].tolist(), "golden total", grads["vq.project_out.bias"][:4].tolist())
