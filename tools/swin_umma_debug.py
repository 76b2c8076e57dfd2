import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D

window, shift, dims = (2, 4, 4), (0, 0, 0), (2, 3, 8, 16, 24)
if len(sys.argv) > 1 and sys.argv[1] == "s2":
    window, shift, dims = (8, 1, 1), (0, 0, 0), (2, 3, 8, 12, 20)
N, V, T, H, W = dims
if len(sys.argv) > 2:
    V = int(sys.argv[2])
torch.manual_seed(5)
blocks = [SwinTransformerBlock3D(16, 2, window, shift, 4., True).cuda() for _ in range(V)]
with torch.no_grad():
    for b in blocks:
        for p in b.parameters():
            p.normal_(0.0, 0.25)
pack = ops.ParamPack([b.packed_parameters() for b in blocks])
ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)
gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)
res = {}
for name, prec, umma in (("fp32", "fp32", False), ("umma", "bf16", True)):
    _lib.set_precision(prec); _lib.set_swin_umma(umma)
    for p in pack.params():
        p.grad = None
    xin = (x if umma else x.float()).clone().requires_grad_(True)
    y = ops.swin_block(xin, pack, idx, ws, ss, rows, scale, heads, hidden)
    y.backward(gy if umma else gy.float())
    torch.cuda.synchronize()
    res[name] = (xin.grad.float().clone(), [p.grad.clone() for p in pack.params()])
names = ["rpb", "qkv.w", "qkv.b", "proj.w", "proj.b", "fc1.w", "fc1.b", "fc2.w", "fc2.b"]
def rl2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
gu, gf = res["umma"][0], res["fp32"][0]
print("gx finite frac", float(torch.isfinite(gu).float().mean()), "rel_l2 (finite part)", rl2(torch.nan_to_num(gu), gf))
bad = (~torch.isfinite(gu)).nonzero()
print("non-finite count", bad.shape[0], "first:", bad[:5].tolist())
err = (torch.nan_to_num(gu) - gf).abs().amax(-1)        # per token
print("per-token max err: mean", float(err.mean()), "max", float(err.max()), "ref max", float(gf.abs().max()))
for v in range(V):
    if os.environ.get("QUIET"):
        a = torch.cat([t.reshape(-1) for t in res["umma"][1][v * 9:(v + 1) * 9]])
        print(f"v{v} params finite {float(torch.isfinite(a).float().mean()):.3f}")
        continue
    for i, n in enumerate(names):
        a, b = res["umma"][1][v * 9 + i], res["fp32"][1][v * 9 + i]
        print(f"v{v} {n:7s} finite {float(torch.isfinite(a).float().mean()):.3f} rel_l2 {rl2(torch.nan_to_num(a), b):.3e} |ref| {float(b.abs().max()):.3e}")
badtok = (~torch.isfinite(gu)).any(-1).nonzero()
wins = {}
for n, v, t, h, w in badtok.tolist():
    win = ((n * (T // window[0]) + t // window[0]) * (H // window[1]) + h // window[1]) * (W // window[2]) + w // window[2]
    wins.setdefault(v, set()).add(win)
for v in sorted(wins):
    print("v", v, "bad windows", sorted(wins[v]))
nb = torch.isnan(gu).sum().item(); ni = torch.isinf(gu).sum().item()
print("nan", nb, "inf", ni)
