import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.use_deterministic_algorithms(True, warn_only=True)
torch.utils.deterministic.fill_uninitialized_memory = True      # torch.empty() -> NaN: an uninitialised read shows up as NaN
from tests.golden_util import load_case, rel_err
from tests.test_parity_gpu import build_model
from idee_b200.models.losses import train_step_loss
from idee_b200 import _lib
case = sys.argv[1] if len(sys.argv) > 1 else "lfq_4_codes"
_lib.set_precision(sys.argv[2] if len(sys.argv) > 2 else "fp32")
cfg, sd, ins, train, ev, grads = load_case(case)
model = build_model(cfg, sd)
for it in range(2):
    model.zero_grad(set_to_none=True)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    total.backward()
    named = dict(model.named_parameters())
    bad = [k for k in grads if not torch.isfinite(named[k].grad).all()]
    errs = sorted(((rel_err(torch.nan_to_num(named[k].grad), g), k) for k, g in grads.items()), reverse=True)
    print(f"iteration {it}: total {float(total):.6f} (golden {float(train['total']):.6f}); non-finite grads: {len(bad)} {bad[:6]}; worst {errs[0]}")
    for k in ("pred", "z_q"):
        print("   ", k, "finite", bool(torch.isfinite(out[k]).all()))
if len(sys.argv) > 3:
    from idee_b200 import ops
    orig = ops.LFQGeneralFn.backward
    mode = sys.argv[3]
    def dbg(ctx, gzq, gidx, gaux):
        if mode == "pre": torch.cuda.synchronize()
        if mode == "clone": gzq = gzq.clone()
        if mode == "print": print("     gzq contiguous", gzq.is_contiguous(), gzq.stride(), gzq.shape, "storage_offset", gzq.storage_offset(), "gaux", gaux.shape if gaux is not None else None, gaux.stride() if gaux is not None else None)
        r = orig(ctx, gzq, gidx, gaux)
        if mode == "post": torch.cuda.synchronize()
        return r
    ops.LFQGeneralFn.backward = staticmethod(dbg)
    model.zero_grad(set_to_none=True)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    total.backward()
    errs = sorted(((rel_err(torch.nan_to_num(named[k].grad), g), k) for k, g in grads.items()), reverse=True)
    print(f"mode {mode}: worst {errs[0]}")
print("---- per-parameter errors (last run)")
for e, k in errs:
    if k.startswith("vq.") or "proj_var" in k or "blocks.0.mlp.fc2" in k: print(f"{e:.3e} {k}   |golden| {float(grads[k].abs().max()):.3e} |got| {float(named[k].grad.abs().max()):.3e}")
