"""Tiny driver: proj_var conv (16->16, 3x3x3 replicate) forward + backward at the benchmark shape, bf16 path (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
_lib.set_precision("bf16")
torch.manual_seed(0)
x = torch.randn(8, 6, 8, 200, 200, 16, device="cuda", requires_grad=True)
w = (torch.randn(6, 16, 16, 3, 3, 3, device="cuda") * 0.05).requires_grad_(True)
b = torch.zeros(6, 16, device="cuda", requires_grad=True)
for _ in range(3):
    y = ops.conv3d_cl(x, w, b, True, True)
    y.sum().backward()
torch.cuda.synchronize()
print("ok", float(y.sum()))
