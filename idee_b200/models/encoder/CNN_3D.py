"""B200-native 3D-CNN encoder: drop-in for the reference ``models/encoder/CNN_3D.py`` (BASELINE.json configs[3]).

Same constructor arguments, module tree and parameter names (``state_dict`` compatible, bit-identical initialisation under the
same seed) and the same ``forward(x[N,V,C,D,H,W]) -> [N,V,out_channels[-1],D,H,W]`` contract (CNN_3D.py:149-237).  Compute: the V
per-variable encoders run as one batched sequence of CUDA kernels on channel-last tokens

    embed (1x1x1 conv, no bias) + LN  ->  [conv3^3(replicate) -> LN(affine) -> ReLU -> +shortcut] x 2 per block  ->  proj_var

The 3x3x3 convs are the encoder's proj-conv kernels (conv_tc.cu; tcgen05 / TMEM conv16_umma.cu on bf16 storage in bf16 mode), the
LayerNorm + ReLU + residual tail is one fused streaming kernel each way (cnn_enc.cu).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib, ops


class PatchEmbed3D(nn.Module):
    """Conv3d(in, embed, k = stride = patch, bias=False) + LayerNorm(no affine) (CNN_3D.py:17-71); built for patch (1,1,1)."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=16, embed_dim=64, norm_layer=None):
        super().__init__()
        self.patch_size, self.in_chans, self.embed_dim = tuple(patch_size), in_chans, embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.norm = norm_layer(embed_dim, elementwise_affine=False) if norm_layer is not None else None
        if self.patch_size != (1, 1, 1) or self.norm is None:
            raise NotImplementedError("idee_b200: the CNN_3D encoder embeds with patch (1,1,1) + LayerNorm (CNN_3D.py:113-117)")


class conv_block(nn.Module):
    """Parameter holder of one residual block (CNN_3D.py:74-147); the arithmetic runs variable-batched in CNN_3D.forward_tokens."""

    def __init__(self, in_channels=96, out_channels=96, kernel_size=(3, 3, 3), drop_rate=0., drop_path=0.):
        super().__init__()
        if drop_path > 0. or tuple(kernel_size) != (3, 3, 3):
            raise NotImplementedError("idee_b200: conv_block is built for 3x3x3 kernels without stochastic depth (reference defaults)")
        self.in_channels, self.out_channels, self.drop_rate = in_channels, out_channels, drop_rate
        self.drop_path = nn.Identity()
        self.conv1 = nn.Conv3d(out_channels, out_channels, kernel_size=kernel_size, stride=(1, 1, 1), padding=(1, 1, 1),
                               padding_mode='replicate', bias=False)
        self.norm1 = nn.LayerNorm(out_channels)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size=kernel_size, stride=(1, 1, 1), padding=(1, 1, 1),
                               padding_mode='replicate', bias=False)
        self.norm2 = nn.LayerNorm(out_channels)
        self.act = nn.ReLU(inplace=True)
        self.downsample = PatchEmbed3D(patch_size=(1, 1, 1), in_chans=in_channels, embed_dim=out_channels,
                                       norm_layer=nn.LayerNorm) if in_channels != out_channels else None

    def forward(self, x):
        raise NotImplementedError("conv_block is evaluated variable-batched by CNN_3D.forward in idee_b200")


class CNN_3D(nn.Module):
    def __init__(self, in_vars: int = 6, in_channels: int = 1, out_channels: list = None, drop_path_rate: float = 0.,
                 drop_rate: float = 0.):
        super().__init__()
        self.in_vars = in_vars
        self.out_channels = list(out_channels) if out_channels is not None else [16, 16]
        self.n_layers = len(self.out_channels)
        self.in_channels = [in_channels] + self.out_channels[:-1]
        self.drop_path_rate, self.drop_rate = drop_path_rate, drop_rate
        if any(c != 16 for c in self.out_channels) or in_channels == 16:
            raise NotImplementedError("idee_b200: the CNN_3D encoder is built for out_channels 16 and in_channels != 16 (config.py:51)")
        self.layers_var, self.proj_var = nn.ModuleList(), nn.ModuleList()
        E = self.out_channels[-1]
        for _ in range(in_vars):
            self.layers_var.append(nn.ModuleList([conv_block(self.in_channels[l], self.out_channels[l], (3, 3, 3), drop_rate, drop_path_rate)
                                                  for l in range(self.n_layers)]))
            self.proj_var.append(nn.Sequential(
                nn.Conv3d(E, E, kernel_size=3, stride=1, padding=1, padding_mode='replicate', bias=True), nn.ReLU(),
                nn.Conv3d(E, E, kernel_size=3, stride=1, padding=1, padding_mode='replicate', bias=True)))
        self.init_weights()
        self.register_buffer("_zero_bias", torch.zeros(in_vars, 16), persistent=False)
        self._packs = None

    def init_weights(self):
        """trunc_normal(.02) on every Linear/Conv weight, zero biases, LayerNorm (1, 0) (CNN_3D.py:196-212)."""
        def _init(m):
            if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv3d)):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm) and m.elementwise_affine:
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        self.apply(_init)

    def _build_packs(self):
        V = self.in_vars
        pk = {"embed_w": ops.ParamPack([[self.layers_var[v][0].downsample.proj.weight] for v in range(V)]), "blocks": []}
        for l in range(self.n_layers):
            if l > 0 and self.layers_var[0][l].downsample is not None:
                raise NotImplementedError("idee_b200: a channel-changing block after the first is not built")
            blk = {}
            for name in ("conv1", "conv2"):
                blk[name] = ops.ParamPack([[getattr(self.layers_var[v][l], name).weight] for v in range(V)])
            for name in ("norm1", "norm2"):
                blk[name + "_w"] = ops.ParamPack([[getattr(self.layers_var[v][l], name).weight] for v in range(V)])
                blk[name + "_b"] = ops.ParamPack([[getattr(self.layers_var[v][l], name).bias] for v in range(V)])
            pk["blocks"].append(blk)
        for i in (0, 2):
            pk[f"proj{i}_w"] = ops.ParamPack([[self.proj_var[v][i].weight] for v in range(V)])
            pk[f"proj{i}_b"] = ops.ParamPack([[self.proj_var[v][i].bias] for v in range(V)])
        self._zero_pack = ops.ParamPack([[nn.Parameter(self._zero_bias[v], requires_grad=False)] for v in range(V)])
        self._packs = pk

    def packs(self):
        if self._packs is None:
            self._build_packs()
        pk = self._packs
        out = [pk["embed_w"]]
        for blk in pk["blocks"]:
            out += [blk["conv1"], blk["norm1_w"], blk["norm1_b"], blk["conv2"], blk["norm2_w"], blk["norm2_b"]]
        return out + [pk["proj0_w"], pk["proj0_b"], pk["proj2_w"], pk["proj2_b"]]

    def forward_tokens(self, x: torch.Tensor, fold_last=None) -> torch.Tensor:
        """x [N,V,C,D,H,W] -> channel-last encoder output [N,V,D,H,W,16]."""
        if x.dim() == 5 and self.in_channels[0] == 1:
            x = x.unsqueeze(2)
        if self._packs is None:
            self._build_packs()
        N, V, Cin, D, H, W = x.shape
        assert V == self.in_vars and Cin == self.in_channels[0], "input must be [N, in_vars, in_channels, D, H, W]"
        pk = self._packs
        E = 16
        bf16 = _lib.PRECISION == "bf16"
        zb = self._zero_bias if self._zero_bias.device == x.device else self._zero_bias.to(x.device)
        tok = ops.embed_ln(x, pk["embed_w"], self._zero_pack)                       # bias=False: a zero bias pack (no gradient)
        tok16 = tok.detach().to(torch.bfloat16) if bf16 else None
        for blk in pk["blocks"]:
            for conv, norm in (("conv1", "norm1"), ("conv2", "norm2")):
                w = ops.packed(blk[conv], (V, E, E, 3, 3, 3))
                y = ops.conv3d_cl(tok, w, zb, proj=True, relu=False, x16=tok16)     # conv3^3 replicate, no bias
                res = ops.ln_act_res(y, tok, blk[norm + "_w"], blk[norm + "_b"], want_bf16=bf16)
                tok, tok16 = res if bf16 else (res, None)
        w0, b0 = ops.packed(pk["proj0_w"], (V, E, E, 3, 3, 3)), ops.packed(pk["proj0_b"], (V, E))
        w2, b2 = ops.packed(pk["proj2_w"], (V, E, E, 3, 3, 3)), ops.packed(pk["proj2_b"], (V, E))
        tok = ops.conv3d_cl(tok, w0, b0, proj=True, relu=True, consumer_masks=True, x16=tok16, out_bf16=bf16)
        if fold_last is not None and bf16:
            w_in, b_in = fold_last
            wf = torch.einsum('vocthw,o->vcthw', w2, w_in.reshape(-1)).unsqueeze(1)
            bf = (b2 @ w_in.reshape(-1) + b_in.reshape(())).unsqueeze(1)
            return ops.conv3d_cl(tok, wf, bf, proj=True, relu=False, input_is_relu=True).squeeze(-1)
        return ops.conv3d_cl(tok, w2, b2, proj=True, relu=False, input_is_relu=True)

    def forward(self, x):
        """x [N,V,C,D,H,W] -> [N,V,16,D,H,W] (a permuted view of the channel-last result)."""
        return self.forward_tokens(x).permute(0, 1, 5, 2, 3, 4)
