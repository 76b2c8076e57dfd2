"""Tiny driver: proj_var conv -> ReLU -> conv (16->16, 3x3x3 replicate) forward + backward at the benchmark shape, bf16 path
with bf16 activation storage as the encoder uses it (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
_lib.set_precision("bf16")
torch.manual_seed(0)
x = torch.randn(8, 6, 8, 200, 200, 16, device="cuda", requires_grad=True)
x16 = x.detach().to(torch.bfloat16)
w0 = (torch.randn(6, 16, 16, 3, 3, 3, device="cuda") * 0.05).requires_grad_(True)
w1 = (torch.randn(6, 16, 16, 3, 3, 3, device="cuda") * 0.05).requires_grad_(True)
b0 = torch.zeros(6, 16, device="cuda", requires_grad=True)
b1 = torch.zeros(6, 16, device="cuda", requires_grad=True)
for _ in range(3):
    h = ops.conv3d_cl(x, w0, b0, True, True, consumer_masks=True, x16=x16, out_bf16=True)
    z = ops.conv3d_cl(h, w1, b1, True, False, input_is_relu=True)
    z.sum().backward()
torch.cuda.synchronize()
print("ok", float(z.sum()))
