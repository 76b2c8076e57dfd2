"""Tiny driver: the 96->96 classifier conv (forward + backward) at the benchmark shape on the tcgen05 path (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
_lib.set_precision("bf16"); _lib.set_umma(os.environ.get("IDEE_B200_UMMA", "1") == "1")
torch.manual_seed(0)
x = torch.randn(8, 1, 4, 200, 200, 96, device="cuda", requires_grad=True)
w = (torch.randn(1, 96, 96, 2, 3, 3, device="cuda") * 0.05).requires_grad_(True)
b = torch.zeros(1, 96, device="cuda", requires_grad=True)
for _ in range(3):
    y = ops.conv3d_cl(x, w, b, False, True)
    y.sum().backward()
torch.cuda.synchronize()
print("ok", float(y.sum()))
