"""Times the Swin block entry points at the benchmark shape (B=8, V=6, T=8, 200x200), tcgen05 vs mma.sync kernels.
usage: python tools/swin_bench.py [fwd|bwd|both] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N, V, T, H, W = 8, 6, 8, 200, 200
_lib.set_precision("bf16")
torch.manual_seed(0)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for window, shift in (((2, 4, 4), (0, 0, 0)), ((2, 4, 4), (1, 2, 2)), ((8, 1, 1), (0, 0, 0))):
    blocks = [SwinTransformerBlock3D(16, 2, window, shift, 4., True).cuda() for _ in range(V)]
    pack = ops.ParamPack([b.packed_parameters() for b in blocks])
    ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
    x32 = torch.randn(N, V, T, H, W, 16, device="cuda")
    x16 = x32.to(torch.bfloat16)
    for umma in (False, True):
        _lib.set_swin_umma(umma)
        x = x16 if umma else x32
        if what in ("fwd", "both"):
            with torch.no_grad():
                ms = timeit(lambda: ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden))
            print(f"fwd (no ymid) window {window} shift {shift} {'tcgen05' if umma else 'mma.sync'}: {ms:.3f} ms")
        if what in ("bwd", "both"):
            xr = x.clone().requires_grad_(True)
            y = ops.swin_block(xr, pack, idx, ws, ss, rows, scale, heads, hidden)
            gy = torch.randn_like(y)
            ms = timeit(lambda: torch.autograd.grad(y, xr, gy, retain_graph=True))
            print(f"bwd window {window} shift {shift} {'tcgen05' if umma else 'mma.sync'}: {ms:.3f} ms")
    del x32, x16
