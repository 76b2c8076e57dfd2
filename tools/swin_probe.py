"""Tiny driver: one Swin block (window (2,4,4), shifted) forward + backward at the benchmark shape, bf16 path (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib, ops
from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D
_lib.set_precision("bf16")
torch.manual_seed(0)
blocks = [SwinTransformerBlock3D(16, 2, (2, 4, 4), (1, 2, 2), 4., True).cuda() for _ in range(6)]
pack = ops.ParamPack([b.packed_parameters() for b in blocks])
ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(8, 200, 200)
x = torch.randn(8, 6, 8, 200, 200, 16, device="cuda", requires_grad=True)
for _ in range(3):
    y = ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden)
    y.sum().backward()
torch.cuda.synchronize()
print("ok", float(y.sum()))
