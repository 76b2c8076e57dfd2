"""B200-native 3-D CNN extremes classifier: drop-in for the reference ``models/classifier/CNN_3D.py``.

Same constructor, module tree and parameter names (``conv1..3`` joint head, ``layers[v].conv1..3`` per-variable heads) and
the same ``forward(x[N,V,C,T,H,W]) -> (z[N,1,H,W], [y_v[N,1,H,W]]*V)`` contract (classifier/CNN_3D.py:63-139).  The six
per-variable heads run as ONE batched launch per layer (weight set = variable), the joint head reads the V channel-last
planes of z_q directly as a 16*V-channel image (no reshape copy), ReLU is fused into the conv epilogue.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib, ops


def _conv():
    return dict(kernel_size=(2, 3, 3), stride=(2, 1, 1), padding=(0, 1, 1), bias=True)


class Layer_v(nn.Module):
    """Per-variable head (classifier/CNN_3D.py:17-58)."""

    def __init__(self, embed_dim: int = 16, dim: int = 16, n_classes: int = 1, drop_rate: int = 0.1):
        super().__init__()
        self.embed_dim, self.dim, self.n_classes, self.drop_rate = embed_dim, dim, n_classes, drop_rate
        self.conv1 = nn.Conv3d(embed_dim, dim, **_conv())
        self.conv2 = nn.Conv3d(dim, dim, **_conv())
        self.conv3 = nn.Conv3d(dim, n_classes, **_conv())
        self.act = nn.ReLU()
        self.drop = nn.Dropout(drop_rate, inplace=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,C,T,H,W] -> [N,n_classes,H,W]"""
        tok = _channel_last(x.unsqueeze(1))
        out = _head(tok, [(c.weight.unsqueeze(0), c.bias.unsqueeze(0)) for c in (self.conv1, self.conv2, self.conv3)],
                    self.drop, groups=1, fuse1=self.drop_rate == 0 or not self.training)
        return out[:, 0].permute(0, 4, 1, 2, 3).squeeze(2)


class HeadLogits(list):
    """The reference's list of per-variable logit maps [y_v [N,1,H,W]] * V (classifier/CNN_3D.py:139); ``stacked`` [N,V,H,W] is
    the one tensor they are views of, so a loss over all heads can run as a single batched launch."""
    stacked = None


def _channel_last(x: torch.Tensor) -> torch.Tensor:
    """logical [N,V,C,T,H,W] -> token view [N,V,T,H,W,C] with C contiguous (copy only if the storage is not channel-last)."""
    tok = x.permute(0, 1, 3, 4, 5, 2)
    return tok if tok.stride(5) == 1 and ops._dense(tok) else tok.contiguous()


def _head(tok, wb, drop, groups, first=None, fuse1=True, tok16=None):
    """conv1 -> ReLU -> Dropout -> conv2 -> ReLU -> conv3.  Each ReLU output has exactly one consumer, so the ReLU backward
    masks are applied in the consumers' data-gradient epilogues (fuse1 is False when an active Dropout sits in between)."""
    (w1, b1), (w2, b2), (w3, b3) = wb
    # bf16 mode, 16 -> 16 heads: the first ReLU activation (and its gradient) is only ever read by tensor-core kernels that round
    # it to bf16 as they load it, so it is stored as bf16 (bit-identical, see Conv3dCL)
    y1_bf16 = first is None and groups == 1 and fuse1 and _lib.PRECISION == "bf16" and w1.shape[1] == 16 and w1.shape[2] == 16
    h = first if first is not None else ops.conv3d_cl(tok, w1, b1, proj=False, relu=True, groups=groups, consumer_masks=fuse1,
                                                      x16=tok16 if y1_bf16 else None, out_bf16=y1_bf16)
    h = drop(h)
    h = ops.conv3d_cl(h, w2, b2, proj=False, relu=True, input_is_relu=fuse1, consumer_masks=True)
    return ops.conv3d_cl(h, w3, b3, proj=False, relu=False, input_is_relu=True)


class CNN_3D(nn.Module):
    def __init__(self, in_var: int = 6, embed_dim: int = 16, dim: int = 16, n_classes: int = 1, drop_rate: int = 0.2):
        super().__init__()
        if embed_dim != 16 or dim % 16 != 0 or n_classes != 1:
            raise NotImplementedError("idee_b200: classifier kernels are built for embed_dim 16, dim multiple of 16, 1 class")
        self.in_var = in_var
        self.var_embed_dim, self.var_dim = embed_dim, dim
        self.embed_dim = embed_dim * in_var
        self.dim = dim * in_var
        self.n_classes, self.drop_rate = n_classes, drop_rate
        self.conv1 = nn.Conv3d(self.embed_dim, self.dim, **_conv())
        self.conv2 = nn.Conv3d(self.dim, self.dim, **_conv())
        self.conv3 = nn.Conv3d(self.dim, self.n_classes, **_conv())
        self.layers = nn.ModuleList([Layer_v(embed_dim=embed_dim, dim=dim, n_classes=1, drop_rate=drop_rate) for _ in range(in_var)])
        self.act = nn.ReLU()
        self.drop = nn.Dropout(drop_rate, inplace=False)
        self._packs = None

    def init_weights(self):
        """trunc_normal(.04) (classifier/CNN_3D.py:95-110; never called by the reference)."""
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv3d)):
                nn.init.trunc_normal_(m.weight, std=.04)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def packs(self):
        """ParamPacks of the per-variable heads (the joint head's parameters are used in place)."""
        if self._packs is None:
            V = self.in_var
            self._packs = [(ops.ParamPack([[getattr(self.layers[v], n).weight] for v in range(V)]),
                            ops.ParamPack([[getattr(self.layers[v], n).bias] for v in range(V)])) for n in ("conv1", "conv2", "conv3")]
        return [p for pair in self._packs for p in pair]

    def _head_params(self):
        self.packs()
        V, e, d = self.in_var, self.var_embed_dim, self.var_dim
        shapes = [(V, d, e, 2, 3, 3), (V, d, d, 2, 3, 3), (V, 1, d, 2, 3, 3)]
        return [(ops.packed(pw, s), ops.packed(pb, (V, s[1]))) for (pw, pb), s in zip(self._packs, shapes)]

    def _joint_conv1_rank1(self, xq, w_out, b_out):
        """First conv of the joint head on the rank-1 form of z_q.

        z_q[n,v,c,t,h,w] = x[n,v,t,h,w] * w_out[c] + b_out[c] (LFQ.py:284), so Conv3d(V*C -> dim) over z_q equals a conv over
        V+1 planes -- the V scalar planes x_v with weights sum_c W[o, v*C+c] w_out[c] and a plane of ones with weights
        sum_{v,c} W[o, v*C+c] b_out[c] (the ones plane reproduces the zero-padding border exactly).  Same result (fp32
        rounding aside), 1/6 of the multiply-adds, and +-1 / 1 / 0 inputs are exact in bf16.  The folded weights are built with
        differentiable torch ops, so gradients reach conv1.weight, project_out.weight and project_out.bias through autograd."""
        N, V, T, H, W = xq.shape
        Co, C = self.dim, self.var_embed_dim
        planes = ops.Rank1Planes.apply(xq)                    # [N,T,H,W,16]: V scalar planes | ones | zeros
        W5 = self.conv1.weight.view(Co, V, C, 2, 3, 3)
        Wq = torch.einsum('ovcthw,c->ovthw', W5, w_out.reshape(-1))
        Wb = torch.einsum('ovcthw,c->othw', W5, b_out.reshape(-1))
        W16 = torch.cat([Wq, Wb.unsqueeze(1), Wq.new_zeros(Co, 15 - V, 2, 3, 3)], dim=1)
        # only the V scalar planes need a gradient (ones / zero planes are constants): cin_real lets the data gradient skip the rest
        return ops.conv3d_cl(planes.unsqueeze(1), W16.unsqueeze(0), self.conv1.bias.unsqueeze(0), proj=False, relu=True,
                             consumer_masks=self._fuse1(), cin_real=V + 1)

    def _fuse1(self) -> bool:
        """conv2 may apply conv1's ReLU mask only when the Dropout between them is the identity."""
        return self.drop_rate == 0 or not self.training

    def forward(self, x, rank1=None):
        """x [N,V,C,T,H,W] -> (z [N,n_classes,H,W], [y_v [N,1,H,W]] * V).
        rank1 (optional): (xq [N,V,T,H,W], w_out [C,1], b_out [C]) with x == xq * w_out + b_out, as produced by LFQ."""
        N, V, C, T, H, W = x.shape
        tok = _channel_last(x)
        # multi-head classifier: all V heads per launch
        tok16 = getattr(x, "_idee_bf16", None)                # bf16 copy of z_q written by the quantiser kernel (VQ_model)
        if tok16 is not None and (tok16.shape != tok.shape or tok16.stride() != tok.stride()):
            tok16 = None
        yh = _head(tok, self._head_params(), self.drop, groups=1, fuse1=self._fuse1(), tok16=tok16)   # [N,V,T',H,W,1]
        y = HeadLogits(yh[:, i].permute(0, 4, 1, 2, 3).squeeze(2) for i in range(self.in_var))
        if yh.shape[2] == 1:
            y.stacked = yh[:, :, 0, :, :, 0]                                              # [N,V,H,W] view
        # joint head over the V*C channels
        first = None
        if rank1 is not None and V + 1 <= 16 and C == self.var_embed_dim:
            first = self._joint_conv1_rank1(*rank1)
        zj = _head(tok, [(c.weight.unsqueeze(0), c.bias.unsqueeze(0)) for c in (self.conv1, self.conv2, self.conv3)],
                   self.drop, groups=V, first=first, fuse1=self._fuse1())         # [N,1,T',H,W,1]
        z = zj[:, 0].permute(0, 4, 1, 2, 3).squeeze(2)
        return z, y
