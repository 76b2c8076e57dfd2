"""Warp-specialised tcgen05 96->96 classifier conv vs the mma.sync kernel: forward + data gradient, small + benchmark shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
_lib.set_precision("bf16")


def run(shape, umma, reps=1):
    _lib.set_umma96(umma)
    N, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, 6, T, H, W, 16, device="cuda", generator=g).relu_().requires_grad_(True)   # joint head: 6 planes of 16 = 96 ch
    w = (torch.randn(1, 96, 96, 2, 3, 3, device="cuda", generator=g) * 0.03).requires_grad_(True)
    b = (torch.randn(1, 96, device="cuda", generator=g) * 0.1).requires_grad_(True)
    To = (T - 2) // 2 + 1
    gy = torch.randn(N, 1, To, H, W, 96, device="cuda", generator=g)
    xin = x.permute(0, 2, 3, 4, 1, 5).reshape(N, 1, T, H, W, 96)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        x.grad = None
        y = ops.conv3d_cl(xin, w, b, False, True, input_is_relu=True)
        y.backward(gy)
    ev[1].record()
    torch.cuda.synchronize()
    return y.detach(), x.grad.clone(), ev[0].elapsed_time(ev[1]) / reps


for shape in [(1, 2, 16, 8), (1, 4, 21, 37), (2, 8, 40, 48), (8, 4, 200, 200)]:
    y0, gx0, t0 = run(shape, False)
    y1, gx1, t1 = run(shape, True)
    print(f"shape {shape}: y maxdiff {float((y0 - y1).abs().max()):.3e} (max {float(y0.abs().max()):.2f})  "
          f"gx maxdiff {float((gx0 - gx1).abs().max()):.3e} (max {float(gx0.abs().max()):.2f})", flush=True)
_, _, t0 = run((8, 4, 200, 200), False, reps=5)
_, _, t1 = run((8, 4, 200, 200), True, reps=5)
print(f"fwd+bwd per rep: mma.sync {t0:.2f} ms, tcgen05 {t1:.2f} ms")
