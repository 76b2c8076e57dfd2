"""tcgen05 16->16 proj conv vs the mma.sync kernel: forward (bf16 / fp32 output) and data gradient, small + benchmark shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
_lib.set_precision("bf16")
torch.manual_seed(0)


def run(shape, umma, out_bf16, reps=1):
    _lib.set_umma16(umma)
    N, V, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).requires_grad_(True)
    x16 = x.detach().to(torch.bfloat16)
    w = (torch.randn(V, 16, 16, 3, 3, 3, device="cuda", generator=g) * 0.08).requires_grad_(True)
    b = (torch.randn(V, 16, device="cuda", generator=g) * 0.1).requires_grad_(True)
    gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g)
    if out_bf16:
        gy = gy.to(torch.bfloat16)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        x.grad = None
        y = ops.conv3d_cl(x, w, b, True, True, consumer_masks=True, x16=x16, out_bf16=out_bf16)
        y.backward(gy)
    torch.cuda.synchronize()
    return y.detach().float(), x.grad.clone(), w.grad.clone(), (time.time() - t0) / reps


for shape in [(1, 1, 3, 16, 8), (1, 2, 4, 21, 37), (2, 3, 8, 40, 48), (8, 6, 8, 200, 200)]:
    for ob in (True, False):
        y0, gx0, gw0, t0 = run(shape, False, ob)
        y1, gx1, gw1, t1 = run(shape, True, ob)
        ey = float((y0 - y1).abs().max()), float(y0.abs().max())
        eg = float((gx0 - gx1).abs().max()), float(gx0.abs().max())
        print(f"shape {shape} out_bf16={ob}: y maxdiff {ey[0]:.3e} (max {ey[1]:.2f})  gx maxdiff {eg[0]:.3e} (max {eg[1]:.2f})  "
              f"gw maxdiff {float((gw0 - gw1).abs().max()):.3e}", flush=True)
y0, gx0, gw0, t0 = run((8, 6, 8, 200, 200), False, True, reps=5)
y1, gx1, gw1, t1 = run((8, 6, 8, 200, 200), True, True, reps=5)
print(f"fwd+bwd wall per rep: mma.sync {t0*1e3:.2f} ms, tcgen05 {t1*1e3:.2f} ms")
